"""Helpers shared by the parity tests."""
import hashlib

import numpy as np


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def sym_hash(n):
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(2654435761)) >> np.uint64(24)).astype(np.uint8)


def sym_lcg(n):
    st, out = 1, np.empty(n, dtype=np.uint8)
    for k in range(n):
        st = (st * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        v = (st >> 33) & 0xFF
        out[k] = (v * v) >> 8
    return out


def sym_peaked(n, seed):
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(128 + 1.5 * rng.standard_normal(n)), 0, 255).astype(np.uint8)


SYM_GENS = {"hash": sym_hash, "lcg": sym_lcg, "peaked": lambda n: sym_peaked(n, 7),
            "zeros": lambda n: np.zeros(n, np.uint8)}


def bits_equal(a, b):
    """bitwise equality of float arrays (distinguishes -0.0 / +0.0, treats equal NaN payloads as equal)"""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def max_ulp(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).view(np.int64).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float64).view(np.int64).astype(np.int64)
    a = np.where(a < 0, np.int64(-2**63) - a, a)
    b = np.where(b < 0, np.int64(-2**63) - b, b)
    return int(np.abs(a - b).max()) if a.size else 0


def header_tuple(h, nlay=None):
    """comparable view of either oracle.binding.Header or waverange_b200.api.Header"""
    n = h.nlay if nlay is None else nlay
    deps = list(getattr(h, "deps", None) or getattr(h, "deps_vec"))[:n]
    minv = list(getattr(h, "minval", None) or getattr(h, "minval_vec"))[:n]
    half = h.halfspan if hasattr(h, "halfspan") else h.halfspanval
    return (np.array([h.tolabs, h.midval, half]).tobytes(), int(h.wlev), int(h.nlay),
            np.array(deps).tobytes(), np.array(minv).tobytes())
