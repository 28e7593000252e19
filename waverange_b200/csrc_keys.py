"""numpy twin of dkey()/dunkey() in csrc/wr_common.cuh (order-preserving u64 keys of doubles);
used by the host-side slab logic and its tests."""
import numpy as np


def dkey_np(x):
    b = np.ascontiguousarray(x, dtype=np.float64).view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    return np.where(neg, ~b, b | np.uint64(1 << 63)).astype(np.uint64)


def dunkey_np(k):
    k = np.ascontiguousarray(k, dtype=np.uint64)
    pos = (k >> np.uint64(63)).astype(bool)
    return np.where(pos, k & np.uint64((1 << 63) - 1), ~k).astype(np.uint64).view(np.float64)
