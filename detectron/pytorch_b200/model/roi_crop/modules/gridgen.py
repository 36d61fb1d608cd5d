"""nn.Module wrapper of the affine grid generator; mirrors lib/model/roi_crop/modules/gridgen.py:12-25 (reference)."""
from torch.nn.modules.module import Module

from ..functions.gridgen import AffineGridGenFunction


class _AffineGridGen(Module):
    def __init__(self, height, width, lr=1, aux_loss=False):
        super(_AffineGridGen, self).__init__()
        self.height, self.width, self.lr, self.aux_loss = height, width, lr, aux_loss
        self.f = AffineGridGenFunction(height, width, lr=lr)

    def forward(self, input):
        return self.f(input)
