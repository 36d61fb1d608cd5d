"""the pure-torch grid helpers in front of RoICrop (CPU tests; no CUDA involved)."""
import numpy as np
import torch

from detectron.pytorch_b200.model.roi_crop.functions.crop_resize import RoICropFunction as CropResizeFn
from detectron.pytorch_b200.model.roi_crop.functions.gridgen import AffineGridGenFunction
from detectron.pytorch_b200.model.roi_crop.functions.roi_crop import RoICropFunction
from detectron.pytorch_b200.model.roi_crop.modules.gridgen import _AffineGridGen
from detectron.pytorch_b200.utils.net import affine_grid_gen


def test_affine_grid_gen_function_matches_the_reference_lattice():
    rng = np.random.RandomState(0)
    theta = rng.standard_normal((3, 2, 3)).astype(np.float32)
    H, W = 5, 7
    out = AffineGridGenFunction(H, W)(torch.from_numpy(theta)).numpy()
    assert out.shape == (3, H, W, 2)
    ys = np.arange(-1, 1, 2.0 / H); xs = np.arange(-1, 1, 2.0 / W)      # the reference's lattice (gridgen.py:13-14)
    for b in range(3):
        for i in range(H):
            for j in range(W):
                v = np.array([ys[i], xs[j], 1.0], dtype=np.float32)
                np.testing.assert_allclose(out[b, i, j], theta[b] @ v, rtol=1e-6, atol=1e-6)


def test_affine_grid_gen_function_gradient_is_grid_transposed_times_lattice():
    theta = torch.randn(2, 2, 3, requires_grad=True)
    m = _AffineGridGen(4, 6)
    g = torch.randn(2, 4, 6, 2)
    m(theta).backward(g)
    ys = -1 + 2 * torch.arange(4.) / 4; xs = -1 + 2 * torch.arange(6.) / 6
    base = torch.stack([ys[:, None].expand(4, 6), xs[None, :].expand(4, 6), torch.ones(4, 6)], dim=2).view(1, 24, 3)
    expect = torch.bmm(g.view(2, 24, 2).transpose(1, 2), base.expand(2, -1, -1))       # reference gridgen.py:45
    torch.testing.assert_close(theta.grad, expect, rtol=1e-5, atol=1e-5)


def test_affine_grid_gen_helper_identity_roi_covers_the_map():
    H, W = 20, 30
    rois = torch.tensor([[0, 0.0, 0.0, (W - 1) * 16.0, (H - 1) * 16.0]])
    grid = affine_grid_gen(rois, (H, W), 7)
    assert grid.shape == (1, 7, 7, 2)
    torch.testing.assert_close(grid[0, 0, 0], torch.tensor([-1.0, -1.0]))
    torch.testing.assert_close(grid[0, -1, -1], torch.tensor([1.0, 1.0]))
    half = torch.tensor([[0, 0.0, 0.0, (W - 1) * 8.0, (H - 1) * 8.0]])            # top-left quarter
    g2 = affine_grid_gen(half, (H, W), 3)
    torch.testing.assert_close(g2[0, -1, -1], torch.tensor([0.0, 0.0]), atol=1e-6, rtol=0)


def test_crop_resize_is_the_second_front_end_of_the_same_op():
    """lib/model/roi_crop/functions/crop_resize.py: a second RoICropFunction over the same launchers that records its device."""
    assert issubclass(CropResizeFn, RoICropFunction) and CropResizeFn().device == -1
