"""CPU check of the DECOMPOSITION the row-stationary backward kernel uses (csrc/roi_align_bwd_rows.cu), restated in
numpy: every valid y-sample contributes a unit to the row of its low cell (weight hy) and one to the row of its high
cell (weight ly); every valid x-sample of a unit contributes the taps (x_low, wy * hx) and (x_high, wy * lx), the
latter dropped when the sample is clamped to the last column (its weight is exactly 0); the addend is
(dY / count) * w.  Summed per (row, cell) this must equal the reference's scatter (oracle, fp64 accumulation).
Guards the algorithm against refactors that cannot be run without a GPU."""
import numpy as np
import pytest

from detectron.pytorch_b200 import synthetic as S
from oracle import cpu as O

f32 = np.float32


def _axis(start, binsz, P, g, size):
    """Per-sample (low, high, l, h, valid) with the kernel's rounding (common.cuh: xfrom_coord / xfrom_axis)."""
    out = []
    for s in range(P * g):
        p, i = s // g, s % g
        base = f32(np.float64(f32(p)) * np.float64(binsz) + np.float64(start))                  # FFMA(p, bin, start)
        off = f32(f32(f32(f32(i) + f32(.5)) * binsz) / f32(g))
        v = f32(base + off)
        valid = not (v < -1.0 or v > size)
        v = max(v, f32(0))
        low = int(v)
        if low >= size - 1:
            low = high = size - 1; v = f32(size - 1)
        else:
            high = low + 1
        l = f32(v - f32(low)); h = f32(f32(1) - l)
        out.append((low, high, l, h, valid))
    return out


def rows_backward_numpy(dy, rois, shape, P, sr, scale):
    N, C, H, W = shape
    dx = np.zeros(shape, dtype=np.float64)
    scale = f32(scale)
    count = f32(sr * sr)
    for r, roi in enumerate(rois.astype(np.float32)):
        b = int(roi[0])
        if b < 0 or b >= N:
            continue
        sw, sh = f32(roi[1] * scale), f32(roi[2] * scale)
        rw = max(f32(np.float64(roi[3]) * np.float64(scale) - np.float64(sw)), f32(1))          # FFMA(x2, s, -x1*s)
        rh = max(f32(np.float64(roi[4]) * np.float64(scale) - np.float64(sh)), f32(1))
        bh, bw = f32(rh / f32(P)), f32(rw / f32(P))
        ys, xs = _axis(sh, bh, P, sr, H), _axis(sw, bw, P, sr, W)
        g = (dy[r].astype(np.float32) / count).astype(np.float32)                                # (C, P, P): dY / count
        for i, (ylow, yhigh, ly, hy, yv) in enumerate(ys):
            if not yv:
                continue
            for row, wy in ((ylow, hy), (yhigh, ly)):                                             # the two units of this y-sample
                for j, (xlow, xhigh, lx, hx, xv) in enumerate(xs):
                    if not xv:
                        continue
                    gv = g[:, i // sr, j // sr].astype(np.float64)
                    dx[b, :, row, xlow] += gv * np.float64(f32(wy * hx))
                    if xhigh != xlow:                                                             # clamped: lx == 0 exactly
                        dx[b, :, row, xhigh] += gv * np.float64(f32(wy * lx))
                    else:
                        assert lx == 0
    return dx


@pytest.mark.parametrize("shape,scale,P,sr,n", [((2, 4, 25, 45), 1.0 / 16, 7, 2, 30), ((1, 3, 30, 33), 1.0 / 8, 7, 1, 20),
                                                ((1, 2, 12, 9), 1.0 / 32, 14, 2, 8)])
def test_unit_and_tap_decomposition_equals_the_reference_scatter(shape, scale, P, sr, n):
    rois = np.concatenate([S.make_rois(n, shape, scale, seed=3), S.make_edge_rois(shape, scale)]).astype(np.float32)
    dy = np.random.RandomState(5).standard_normal((rois.shape[0], shape[1], P, P)).astype(np.float32)
    ours = rows_backward_numpy(dy, rois, shape, P, sr, scale)
    ref = O.roi_align_backward(dy, rois, shape, P, P, scale, sr, acc64=True)
    np.testing.assert_allclose(ours, ref, rtol=1e-5, atol=1e-5)
    assert np.array_equal(ours == 0, ref == 0)                                                    # same support
