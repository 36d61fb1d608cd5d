"""Affine sampling-grid generator used in front of RoICrop.

Mirrors lib/model/roi_crop/functions/gridgen.py:7-46 (reference): `AffineGridGenFunction(height, width)(theta)`
maps a batch of 2x3 affine matrices to a `(B, height, width, 2)` grid in the reference's (y, x) order, over the
base lattice `y_i = -1 + 2 i / height`, `x_j = -1 + 2 j / width` (note: NOT align-corners; the last lattice point
is 1 - 2/size, exactly as the reference builds it with np.arange(-1, 1, 2/size)).

The reference keeps this in Python (two bmm calls); so do we -- it is a (B*H*W x 3) @ (3 x 2) product with no
device-specific work.  Plain differentiable torch ops, so the backward (reference :38-46: dTheta = dGrid^T @ base)
comes from autograd.
"""
import torch


def _base_lattice(height, width, like):
    ys = -1.0 + 2.0 * torch.arange(height, dtype=torch.float32, device=like.device) / float(height)
    xs = -1.0 + 2.0 * torch.arange(width, dtype=torch.float32, device=like.device) / float(width)
    base = torch.empty((height, width, 3), dtype=torch.float32, device=like.device)
    base[:, :, 0] = ys[:, None]
    base[:, :, 1] = xs[None, :]
    base[:, :, 2] = 1.0
    return base.to(like.dtype)


class AffineGridGenFunction(object):
    """Legacy-style callable: construct with the grid size, then call with theta `(B, 2, 3)`."""

    def __init__(self, height, width, lr=1):
        self.height, self.width, self.lr = int(height), int(width), lr

    def __call__(self, theta):
        if theta.dim() != 3 or theta.size(1) != 2 or theta.size(2) != 3:
            raise ValueError("theta must be (B, 2, 3), got %s" % (tuple(theta.shape),))
        base = _base_lattice(self.height, self.width, theta).view(1, self.height * self.width, 3)
        out = torch.bmm(base.expand(theta.size(0), -1, -1), theta.transpose(1, 2))
        return out.view(-1, self.height, self.width, 2)

    forward = __call__
