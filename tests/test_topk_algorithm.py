"""CPU restatement of b200_topk_batched's selection logic (detectron/pytorch_b200/csrc/topk.cu): order-preserving key of a float,
three most-significant-first histogram passes (11 + 11 + 10 bits) with the "largest bin whose upper count stays below k" rule,
threshold key T and the number of ties to admit -- against numpy's sort.  No GPU: this pins the algorithm."""
import numpy as np
import pytest


def ordered_key(x):
    b = np.asarray(x, np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def key_to_float(k):
    k = np.asarray(k, np.uint32)
    return np.where(k & np.uint32(0x80000000), k & np.uint32(0x7fffffff), ~k).astype(np.uint32).view(np.float32)


def select_bin(hist, need):
    """largest bin b with count(bins > b) < need <= count(bins >= b); returns (b, need - count(bins > b))"""
    above = 0
    for b in range(len(hist) - 1, -1, -1):
        if need <= above + hist[b]:
            return b, need - above
        above += hist[b]
    raise AssertionError("fewer than `need` elements")


def radix_select(keys, k):
    d0 = keys >> 21; d1 = (keys >> 10) & 2047; d2 = keys & 1023
    b0, need = select_bin(np.bincount(d0, minlength=2048), k)
    m = d0 == b0
    b1, need = select_bin(np.bincount(d1[m], minlength=2048), need)
    m &= d1 == b1
    b2, need = select_bin(np.bincount(d2[m], minlength=2048), need)
    return (np.uint32(b0) << 21) | (np.uint32(b1) << 10) | np.uint32(b2), need


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("kind", ["uniform", "probabilities", "ties", "signed", "tiny"])
def test_radix_select_matches_a_sort(seed, kind):
    rng = np.random.RandomState(seed)
    n = 50000
    if kind == "uniform":
        x = rng.uniform(0, 1, n)
    elif kind == "probabilities":
        x = 1 / (1 + np.exp(-rng.standard_normal(n) * 4))          # sigmoid outputs: clustered near 0 and 1
    elif kind == "ties":
        x = rng.randint(0, 9, n) / 9.0
    elif kind == "signed":
        x = rng.standard_normal(n) * 1e3
    else:
        x = rng.standard_normal(n) * 1e-30
    x = x.astype(np.float32)
    keys = ordered_key(x)
    assert np.array_equal(key_to_float(keys), x)                    # the key is invertible ...
    order = np.argsort(x, kind="stable")
    assert np.all(np.diff(keys[order].astype(np.int64)) >= 0)        # ... and order preserving (negative zero sorts below zero)
    for k in (1, 7, 2000, 12000, n):
        T, ties = radix_select(keys, k)
        gt = int((keys > T).sum()); eq = int((keys == T).sum())
        assert gt < k <= gt + eq and ties == k - gt
        kth = np.sort(keys)[::-1][k - 1]
        assert T == kth
        # the selected multiset is the top-k multiset
        sel = np.concatenate([keys[keys > T], np.full(ties, T, np.uint32)])
        assert np.array_equal(np.sort(sel)[::-1], np.sort(keys)[::-1][:k])
