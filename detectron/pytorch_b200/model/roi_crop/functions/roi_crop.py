"""Drop-in for `model.roi_crop.functions.roi_crop` (reference lib/model/roi_crop/functions/roi_crop.py:7-21).

`RoICropFunction()(input1, input2)`: input1 = image (N,C,H,W), input2 = sampling grid (R,h,w,2) in
(y, x) order.  Gradient w.r.t. the grid is all zeros, as with the reference's CUDA kernel (which
accumulates the dot products but never stores them, roi_crop_cuda_kernel.cu:111-194).  The
reference's two defensive `.clone()`s of the inputs (:9-10) are not needed: nothing is mutated."""
from detectron.pytorch_b200 import ops as _ops


class RoICropFunction(object):
    def __init__(self):
        self.input1 = None
        self.input2 = None

    def __call__(self, input1, input2):
        self.input1, self.input2 = input1, input2
        if not input1.is_cuda:
            raise NotImplementedError("RoICropFunction: CUDA tensors only")
        return _ops._RoICrop.apply(input1, input2)

    def forward(self, input1, input2):
        self.input1, self.input2 = input1, input2
        if not input1.is_cuda:
            raise NotImplementedError("RoICropFunction: CUDA tensors only")
        return _ops.roi_crop_forward(input1.detach(), input2.detach())

    def backward(self, grad_output):
        assert grad_output.is_cuda
        return _ops.roi_crop_backward(grad_output, self.input2.detach(), tuple(self.input1.size()))
