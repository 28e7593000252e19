#!/usr/bin/env python3
"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists and oracle/build_oracle.py has compiled
oracle/_ref/libwaverange_ref_strict.so from it):

    python tests/golden/make_golden.py

Every expected value in the fixture is an output of the reference's own code
(waveletcdf97_3d, range_encode/range_decode via oracle/ref_shim.cpp, encoding_wrap, decoding_wrap)
built with -O2 -ffp-contract=off.  Inputs are stored next to the outputs so the tests never
depend on a random generator's stream.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import Reference, Restatement  # noqa: E402


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


def main():
    ref = Reference("strict")
    gen = Restatement()          # only for the deterministic probe-field generator
    out = {}
    meta = []

    # ---- range coder known answers ---------------------------------------------------------
    cases = [("hash", n) for n in (1, 5, 255, 59999, 60000, 60001, 119999, 120000)]
    cases += [("lcg", 150001), ("peaked", 70000), ("zeros", 10), ("zeros", 60000)]
    for name, n in cases:
        sym = {"hash": sym_hash, "lcg": sym_lcg, "peaked": lambda m: sym_peaked(m, 7),
               "zeros": lambda m: np.zeros(m, np.uint8)}[name](n)
        s = ref.range_encode(sym)
        key = "rc_%s_%d" % (name, n)
        out[key + "_sym"] = sym if n <= 255 else np.zeros(0, np.uint8)
        out[key + "_len"] = np.array([len(s)], dtype=np.int64)
        out[key + "_sha"] = np.frombuffer(hashlib.sha256(s.tobytes()).digest(), dtype=np.uint8)
        out[key + "_head"] = s[:16].copy()
        out[key + "_tail"] = s[-16:].copy()
        if n <= 255:
            out[key + "_stream"] = s
        assert np.array_equal(ref.range_decode(s, n), sym)
        meta.append(key)

    # ---- wavelet known answers -------------------------------------------------------------
    rng = np.random.default_rng(20261018)
    shapes = [(1, 1, 8), (1, 1, 5), (1, 1, 16), (1, 1, 2), (1, 1, 3), (2, 2, 2), (3, 4, 5), (7, 1, 9), (1, 6, 1),
              (17, 9, 33), (16, 16, 16), (5, 18, 31)]
    for shp in shapes:
        x = rng.standard_normal(shp)
        key = "wv_%dx%dx%d" % shp
        out[key + "_in"] = x
        for lvl in (1, 4):
            w = ref.wavelet3d(x, lvl)
            out[key + "_fwd%d" % lvl] = w
            out[key + "_inv%d" % lvl] = ref.wavelet3d(w, -lvl)
        meta.append(key)
    x = np.arange(1, 9, dtype=np.float64).reshape(1, 1, 8)
    out["wv_ramp8_in"] = x
    out["wv_ramp8_fwd1"] = ref.wavelet3d(x, 1)

    # ---- encoding_wrap / decoding_wrap known answers ----------------------------------------
    e2e = [("e2e_a", (20, 24, 32), 1e-6, 1, False), ("e2e_b", (40, 40, 40), 1e-3, 1, True),
           ("e2e_c", (9, 30, 17), 1e-10, 1, False), ("e2e_d", (16, 16, 16), 1e-4, 0, False)]
    for key, shp, tol, wt, f32 in e2e:
        f = gen.probe_field(shp, seed=4242 + len(key), nm=12, round_f32=f32)
        enc = ref.encode(f, tol, wtflag=wt)
        h = enc["header"]
        rec = ref.decode(shp, h, enc["data"])
        out[key + "_in"] = f
        out[key + "_tol"] = np.array([tol, wt], dtype=np.float64)
        out[key + "_scal"] = np.array([h.tolabs, h.midval, h.halfspan], dtype=np.float64)
        out[key + "_int"] = np.array([h.wlev, h.nlay, h.ntot_enc], dtype=np.int64)
        out[key + "_deps"] = np.array(list(h.deps)[:h.nlay])
        out[key + "_minval"] = np.array(list(h.minval)[:h.nlay])
        out[key + "_len"] = np.array(list(h.len)[:h.nlay], dtype=np.int64)
        out[key + "_data"] = enc["data"]
        out[key + "_rec_sha"] = np.frombuffer(hashlib.sha256(rec.tobytes()).digest(), dtype=np.uint8)
        out[key + "_residual_sha"] = np.frombuffer(hashlib.sha256(enc["residual"].tobytes()).digest(), dtype=np.uint8)
        meta.append(key)
    # constant field: trivial exit (wrappers.cpp:257-266)
    f = np.full((4, 5, 6), 3.25)
    enc = ref.encode(f, 1e-6)
    h = enc["header"]
    out["e2e_const_int"] = np.array([h.wlev, h.nlay, h.ntot_enc], dtype=np.int64)
    out["e2e_const_scal"] = np.array([h.tolabs, h.midval, h.halfspan])

    # ---- ind_p2w_3d ---------------------------------------------------------------------------
    pts = []
    for n in [(8, 8, 8), (5, 7, 9), (16, 1, 3)]:
        for i3 in range(n[2]):
            for i2 in range(n[1]):
                for i1 in range(n[0]):
                    pts.append(list(n) + [i1, i2, i3] + list(ref.ind_p2w(4, n, (i1, i2, i3))))
    out["p2w"] = np.array(pts, dtype=np.int32)

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
