"""Shared test helpers: label-permutation-invariant comparison of region-id maps."""
import numpy as np


def partition_equal(a: np.ndarray, b: np.ndarray) -> bool:
    """True iff the two label arrays induce the same partition."""
    a = a.ravel().astype(np.int64)
    b = b.ravel().astype(np.int64)
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    if ua.size != ub.size:
        return False
    pair = ia * ub.size + ib
    return np.unique(pair).size == ua.size


def overseg_iou(ref: np.ndarray, test: np.ndarray) -> float:
    """Over-segmentation IoU up to label permutation: every reference region is matched
    with the test region that overlaps it most, IoU = |R & T| / |R | T|, and the result
    is the area-weighted mean over reference regions (symmetrised by taking the minimum
    of both directions)."""
    def one(a, b):
        a = a.ravel().astype(np.int64)
        b = b.ravel().astype(np.int64)
        ua, ia = np.unique(a, return_inverse=True)
        ub, ib = np.unique(b, return_inverse=True)
        inter = np.zeros((ua.size, ub.size), np.int64) if ua.size * ub.size < 5e7 else None
        if inter is not None:
            np.add.at(inter, (ia, ib), 1)
            sa = inter.sum(1)
            sb = inter.sum(0)
            j = inter.argmax(1)
            best = inter[np.arange(ua.size), j]
            iou = best / (sa + sb[j] - best)
            return float((iou * sa).sum() / sa.sum())
        pair = ia * ub.size + ib
        up, cnt = np.unique(pair, return_counts=True)
        pa, pb = up // ub.size, up % ub.size
        sa = np.bincount(ia, minlength=ua.size)
        sb = np.bincount(ib, minlength=ub.size)
        order = np.lexsort((-cnt, pa))
        first = np.ones(order.size, bool)
        first[1:] = pa[order][1:] != pa[order][:-1]
        sel = order[first]
        best = cnt[sel]
        iou = best / (sa[pa[sel]] + sb[pb[sel]] - best)
        return float((iou * sa[pa[sel]]).sum() / sa.sum())
    return min(one(ref, test), one(test, ref))
