"""Drop-in for `model.roi_crop.functions.crop_resize` (reference lib/model/roi_crop/functions/crop_resize.py:8-37).

The reference file is a second `RoICropFunction`: same CUDA launchers as functions/roi_crop.py
(`BilinearSamplerBHWD_updateOutput_cuda` / `_updateGradInput_cuda`), plus two things of its own, mirrored here:
  * it records the device it ran on (`self.device`: `torch.cuda.current_device()` for CUDA inputs, -1 otherwise, :15-19)
    and allocates the backward's outputs there (:33-35);
  * it has a CPU branch (`BilinearSamplerBHWD_updateOutput`, roi_crop.c:246-497).  That branch is outside SURVEY.md 8:
    CPU tensors raise NotImplementedError, as everywhere in this package.
"""
import torch

from .roi_crop import RoICropFunction as _RoICropFunction


class RoICropFunction(_RoICropFunction):
    def __init__(self):
        super().__init__()
        self.device = -1

    def _note_device(self, input1):
        self.device = input1.device.index if input1.is_cuda else -1
        if self.device is None:
            self.device = torch.cuda.current_device()

    def __call__(self, input1, input2):
        self._note_device(input1)
        return super().__call__(input1, input2)

    def forward(self, input1, input2):
        self._note_device(input1)
        return super().forward(input1, input2)

    def backward(self, grad_output):
        with torch.cuda.device(self.device if self.device >= 0 else grad_output.device):
            return super().backward(grad_output)
