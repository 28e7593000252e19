"""CPU tests: the oracle restatement against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py), and against oracle/_ref directly where it is present."""
import numpy as np
import pytest

from util import SYM_GENS, bits_equal, sha

RC_CASES = [("hash", n) for n in (1, 5, 255, 59999, 60000, 60001, 119999, 120000)] + \
           [("lcg", 150001), ("peaked", 70000), ("zeros", 10), ("zeros", 60000)]
WV_SHAPES = [(1, 1, 8), (1, 1, 5), (1, 1, 16), (1, 1, 2), (1, 1, 3), (2, 2, 2), (3, 4, 5), (7, 1, 9), (1, 6, 1),
             (17, 9, 33), (16, 16, 16), (5, 18, 31)]
E2E = ["e2e_a", "e2e_b", "e2e_c", "e2e_d"]


@pytest.mark.parametrize("name,n", RC_CASES)
def test_range_encode_golden(oracle, golden, name, n):
    sym = SYM_GENS[name](n)
    s = oracle.range_encode(sym)
    key = "rc_%s_%d" % (name, n)
    assert len(s) == int(golden[key + "_len"][0])
    assert np.array_equal(sha(s), golden[key + "_sha"])
    assert np.array_equal(s[:16], golden[key + "_head"]) and np.array_equal(s[-16:], golden[key + "_tail"])
    dec, cnt = oracle.range_decode(s, n)
    assert cnt == n and np.array_equal(dec, sym)


def test_range_encode_empty_trailing_block(oracle):
    # n % 60000 == 0 appends an empty block: 513 extra bytes (hard part 2)
    a = len(oracle.range_encode(SYM_GENS["hash"](59999)))
    b = len(oracle.range_encode(SYM_GENS["hash"](60000)))
    assert b - a == 513


@pytest.mark.parametrize("shape", WV_SHAPES)
@pytest.mark.parametrize("lvl", [1, 4])
def test_wavelet_golden(oracle, golden, shape, lvl):
    key = "wv_%dx%dx%d" % shape
    x = golden[key + "_in"]
    w = oracle.wavelet3d(x, lvl)
    assert bits_equal(w, golden[key + "_fwd%d" % lvl])
    assert bits_equal(oracle.wavelet3d(w, -lvl), golden[key + "_inv%d" % lvl])


def test_wavelet_survey_kat(oracle, golden):
    # waveletcdf97_3d(8,1,1,1,{1..8})
    w = oracle.wavelet3d(golden["wv_ramp8_in"], 1).ravel()
    assert bits_equal(w, golden["wv_ramp8_fwd1"].ravel())
    assert abs(w[0] - 1.8860525098649028) < 1e-15 and abs(w[7] - 0.61170892111963648) < 1e-15


@pytest.mark.parametrize("key", E2E)
def test_encode_decode_golden(oracle, golden, key):
    f = golden[key + "_in"]
    tol, wt = golden[key + "_tol"]
    e = oracle.encode(f, float(tol), wtflag=int(wt))
    h = e["header"]
    assert [h.wlev, h.nlay, h.ntot_enc] == list(golden[key + "_int"])
    assert bits_equal(np.array([h.tolabs, h.midval, h.halfspan]), golden[key + "_scal"])
    assert bits_equal(np.array(list(h.deps)[:h.nlay]), golden[key + "_deps"])
    assert bits_equal(np.array(list(h.minval)[:h.nlay]), golden[key + "_minval"])
    assert list(h.len)[:h.nlay] == list(golden[key + "_len"])
    assert np.array_equal(e["data"], golden[key + "_data"])
    assert np.array_equal(sha(e["residual"]), golden[key + "_residual_sha"])
    rec = oracle.decode(f.shape, h, e["data"])
    assert np.array_equal(sha(rec), golden[key + "_rec_sha"])
    # the reference's accuracy contract (examples/fortran/example_fort.f90:43-45): Linf error ~ tol*max|f|
    assert np.abs(rec - f).max() <= float(tol) * np.abs(f).max()


def test_constant_field_golden(oracle, golden):
    e = oracle.encode(np.full((4, 5, 6), 3.25), 1e-6)
    h = e["header"]
    assert [h.wlev, h.nlay, h.ntot_enc] == list(golden["e2e_const_int"])
    assert bits_equal(np.array([h.tolabs, h.midval, h.halfspan]), golden["e2e_const_scal"])
    assert np.all(oracle.decode((4, 5, 6), h, e["data"]) == 3.25)


def test_ind_p2w_golden(oracle, golden):
    for row in golden["p2w"][::7]:
        n, i, want = row[:3], row[3:6], row[6:]
        assert oracle.ind_p2w(4, tuple(n), tuple(i)) == tuple(want)


def test_chunked_streams_are_reference_streams(oracle, golden):
    """A chunked encode is the per-chunk range_encode of the same symbols; decode agrees bit for bit."""
    f = golden["e2e_b_in"]
    L = 59999
    whole = oracle.encode(f, 1e-3, want_symbols=True)
    ch = oracle.encode(f, 1e-3, chunk_len=L, want_symbols=True)
    assert np.array_equal(whole["symbols"], ch["symbols"])
    off = 0
    for l in range(ch["header"].nlay):
        for c, n in enumerate(ch["chunk_lens"][l]):
            want = oracle.range_encode(ch["symbols"][l][c * L:(c + 1) * L])
            assert np.array_equal(ch["data"][off:off + n], want)
            off += int(n)
    rec = oracle.decode(f.shape, ch["header"], ch["data"], L, ch["chunk_lens"])
    assert bits_equal(rec, oracle.decode(f.shape, whole["header"], whole["data"]))
    # container overhead of chunking stays far below the 1 % ratio budget
    assert ch["header"].ntot_enc <= 1.01 * whole["header"].ntot_enc


# ---- direct comparison with the compiled reference (only where oracle/_ref is present) --------
def test_restatement_vs_reference_random(oracle, ref):
    rng = np.random.default_rng(5)
    for shp in [(6, 7, 8), (1, 1, 37), (13, 1, 4), (32, 32, 32), (11, 21, 31)]:
        x = rng.standard_normal(shp) * 10
        for lvl in (1, 2, 3, 4):
            w = oracle.wavelet3d(x, lvl)
            assert bits_equal(w, ref.wavelet3d(x, lvl))
            assert bits_equal(oracle.wavelet3d(w, -lvl), ref.wavelet3d(w, -lvl))
    for n in (1, 2, 1000, 60000, 123457):
        sym = rng.integers(0, 256, n, dtype=np.uint8)
        s = oracle.range_encode(sym)
        assert np.array_equal(s, ref.range_encode(sym))
        assert np.array_equal(ref.range_decode(s, n), sym)
    f = oracle.probe_field((24, 28, 36))
    for tol in (1e-2, 1e-6, 1e-12, 1e-16):
        a, b = oracle.encode(f, tol), ref.encode(f, tol)
        assert np.array_equal(a["data"], b["data"])
        assert bytes(a["header"]) == bytes(b["header"])
        assert bits_equal(oracle.decode(f.shape, a["header"], a["data"]), ref.decode(f.shape, b["header"], b["data"]))


@pytest.mark.parametrize("wt", [0, 1])
@pytest.mark.parametrize("grid", [(2, 2, 2), (3, 1, 2), (1, 1, 4)])
def test_local_cutoff_restatement_vs_reference(oracle, ref, wt, grid):
    """encoding_wrap with mx*my*mz > 1 (wrappers.cpp:343-379): the restatement against the compiled reference.
    With the transform on, ind_p2w_3d() reports level 4 for every point, so the result is that of the uniform
    minimum cutoff; with the transform off every point is coded to its block's precision."""
    rng = np.random.default_rng(grid[0] * 7 + grid[2] + wt)
    f = oracle.probe_field((12, 10, 14), seed=5 + wt, nm=10)
    vals = 10.0 ** rng.uniform(-6, -1, size=grid[0] * grid[1] * grid[2])
    a = oracle.encode(f, 0.0, wtflag=wt, cutoff=(*grid, vals))
    b = ref.encode(f, 0.0, wtflag=wt, cutoff=(*grid, vals))
    ha, hb = a["header"], b["header"]
    assert (ha.nlay, ha.ntot_enc, ha.tolabs, list(ha.deps), list(ha.minval)) == (hb.nlay, hb.ntot_enc, hb.tolabs, list(hb.deps), list(hb.minval))
    assert a["data"].tobytes() == b["data"].tobytes()
    u = oracle.encode(f, float(vals.min()), wtflag=wt)
    if wt:
        assert u["data"].tobytes() == a["data"].tobytes()
    else:
        assert a["header"].ntot_enc < u["header"].ntot_enc


DIGEST_CASES = [((40, 33, 50), 1e-6, 1, 59999), ((24, 20, 28), 1e-9, 1, 999), ((1, 7, 300), 1e-3, 1, 59999),
                ((30, 30, 30), 1e-4, 0, 4999), ((65, 64, 63), 1e-12, 1, 59999)]


@pytest.mark.parametrize("shape,tol,wt,cl", DIGEST_CASES)
def test_digest_encode_equals_plain_encode(oracle, shape, tol, wt, cl):
    """The multi-threaded digest encode used by the full-size GPU parity tests (wro_encode_digest: same steps, work
    shared out over threads, only hashes kept) against the plain single-threaded restatement -- which the tests above
    pin to the reference: header bytes, chunk lengths, every stream hash and every chunk's symbol hash."""
    f = oracle.probe_field(shape, seed=11, nm=10)
    w = oracle.encode(f, tol, wtflag=wt, chunk_len=cl, want_symbols=True)
    d = oracle.encode_digest(f, tol, cl, wtflag=wt, sample_chunk=1)
    hw, hd = w["header"], d["header"]
    assert bytes(hw) == bytes(hd)
    assert np.array_equal(w["chunk_lens"], d["chunk_lens"])
    ntot = f.size
    off = 0
    for l in range(hw.nlay):
        for c, n in enumerate(w["chunk_lens"][l]):
            n = int(n)
            assert oracle.fnv1a(w["data"][off:off + n]) == int(d["stream_hash"][l, c])
            s0, s1 = c * cl, min(ntot, (c + 1) * cl)
            assert oracle.fnv1a(w["symbols"][l][s0:s1]) == int(d["symbol_hash"][l, c])
            off += n
        if ntot > cl:
            s1 = min(ntot, 2 * cl)
            assert np.array_equal(d["sample"][l][:s1 - cl], w["symbols"][l][cl:s1])
    # hashes of byte ranges pulled from a container, and the inverse from symbols
    offs = np.concatenate([[0], np.cumsum(w["chunk_lens"].ravel().astype(np.uint64))[:-1]]).astype(np.uint64)
    assert np.array_equal(oracle.fnv1a_many(w["data"], offs, w["chunk_lens"].ravel().astype(np.uint64)), d["stream_hash"].ravel())
    whole = oracle.encode(f, tol, wtflag=wt)
    assert bits_equal(oracle.decode_symbols(shape, hw, w["symbols"]), oracle.decode(shape, whole["header"], whole["data"]))


def test_threaded_transform_equals_plain(oracle, ref):
    """wro_wavelet3d_mt (lines shared out over threads) against the compiled reference, odd sizes included"""
    import ctypes as C
    rng = np.random.default_rng(5)
    for shape in [(33, 18, 47), (16, 16, 16), (1, 9, 64)]:
        x = rng.standard_normal(shape)
        for lvl in (1, 4):
            a = x.copy()
            nz, ny, nx = shape
            oracle.lib.wro_wavelet3d_mt(nx, ny, nz, lvl, a.ctypes.data_as(C.POINTER(C.c_double)))
            assert bits_equal(a, ref.wavelet3d(x, lvl))
            oracle.lib.wro_wavelet3d_mt(nx, ny, nz, -lvl, a.ctypes.data_as(C.POINTER(C.c_double)))
            assert bits_equal(a, ref.wavelet3d(ref.wavelet3d(x, lvl), -lvl))
