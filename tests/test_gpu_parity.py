"""GPU parity tests: every stage of the CUDA path, called through the C ABI, against the oracle
(oracle/wr_oracle.c, pinned to the reference) and the golden vectors.  Bit-exact everywhere:
wavelet coefficients (0 ULP), header doubles, symbols, chunk bytes, reconstructed field."""
import os

import numpy as np
import pytest

from util import SYM_GENS, bits_equal, header_tuple, sha

pytestmark = pytest.mark.gpu

F64, F32 = 0, 1
L1 = 59999


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


WV_SHAPES = [(1, 1, 8), (1, 1, 5), (1, 1, 16), (1, 1, 2), (1, 1, 3), (2, 2, 2), (3, 4, 5), (7, 1, 9), (1, 6, 1),
             (17, 9, 33), (16, 16, 16), (5, 18, 31)]


@pytest.mark.parametrize("shape", WV_SHAPES)
@pytest.mark.parametrize("lvl", [1, 4])
def test_wavelet_golden(codec, torch_cuda, golden, shape, lvl):
    key = "wv_%dx%dx%d" % shape
    x = dev(torch_cuda, golden[key + "_in"])
    nz, ny, nx = shape
    codec.wavelet3d_device(x.data_ptr(), nx, ny, nz, lvl)
    assert bits_equal(x.cpu().numpy(), golden[key + "_fwd%d" % lvl])
    codec.wavelet3d_device(x.data_ptr(), nx, ny, nz, -lvl)
    assert bits_equal(x.cpu().numpy(), golden[key + "_inv%d" % lvl])


@pytest.mark.parametrize("shape", [(64, 64, 64), (33, 47, 129), (1, 200, 300), (128, 96, 80), (70, 1, 513), (9, 9, 1025)])
def test_wavelet_vs_oracle(codec, torch_cuda, oracle, shape):
    rng = np.random.default_rng(sum(shape))
    a = rng.standard_normal(shape) * 3
    nz, ny, nx = shape
    for lvl in (1, 3, 4):
        x = dev(torch_cuda, a)
        codec.wavelet3d_device(x.data_ptr(), nx, ny, nz, lvl)
        w = oracle.wavelet3d(a, lvl)
        assert bits_equal(x.cpu().numpy(), w), "forward lvl %d" % lvl
        codec.wavelet3d_device(x.data_ptr(), nx, ny, nz, -lvl)
        assert bits_equal(x.cpu().numpy(), oracle.wavelet3d(w, -lvl)), "inverse lvl %d" % lvl


@pytest.mark.parametrize("name,n", [("hash", 1), ("hash", 5), ("hash", 255), ("hash", 59999), ("hash", 60000),
                                    ("hash", 60001), ("hash", 119999), ("hash", 120000), ("lcg", 150001),
                                    ("peaked", 70000), ("zeros", 10), ("zeros", 60000)])
def test_range_coder_single_stream_golden(codec, torch_cuda, golden, name, n):
    """chunk_len == 0: the whole array is one stream == the reference's range_encode()."""
    sym = SYM_GENS[name](n)
    d_sym = dev(torch_cuda, sym)
    out = torch_cuda.zeros(2 * n + 4096, dtype=torch_cuda.uint8, device="cuda")
    lens, total = codec.range_encode_device(d_sym.data_ptr(), n, 0, out.data_ptr(), out.numel())
    s = out[:total].cpu().numpy()
    key = "rc_%s_%d" % (name, n)
    assert lens == [total] and total == int(golden[key + "_len"][0])
    assert np.array_equal(sha(s), golden[key + "_sha"])
    back = torch_cuda.zeros(n, dtype=torch_cuda.uint8, device="cuda")
    codec.range_decode_device(out.data_ptr(), lens, n, 0, back.data_ptr())
    assert np.array_equal(back.cpu().numpy(), sym)


@pytest.mark.parametrize("n,chunk", [(300000, L1), (300000, 119999), (59999 * 40 + 17, L1), (1000, 300), (120000, 60000),
                                     (59999 * 33, L1)])
@pytest.mark.parametrize("kind", ["uniform", "peaked", "runs"])
def test_range_coder_chunks_vs_oracle(codec, torch_cuda, oracle, n, chunk, kind):
    rng = np.random.default_rng(n + chunk)
    if kind == "uniform":
        sym = rng.integers(0, 256, n, dtype=np.uint8)
    elif kind == "peaked":
        sym = np.clip(np.rint(127.5 + 0.6 * rng.standard_normal(n)), 0, 255).astype(np.uint8)
    else:
        sym = np.repeat(rng.integers(0, 256, n // 97 + 1, dtype=np.uint8), 97)[:n]
    d_sym = dev(torch_cuda, sym)
    out = torch_cuda.zeros(2 * n + 4096 * (n // chunk + 2), dtype=torch_cuda.uint8, device="cuda")
    lens, total = codec.range_encode_device(d_sym.data_ptr(), n, chunk, out.data_ptr(), out.numel())
    blob = out[:total].cpu().numpy()
    off = 0
    for c, ln in enumerate(lens):
        want = oracle.range_encode(sym[c * chunk:(c + 1) * chunk])
        assert ln == len(want), "chunk %d length" % c
        assert np.array_equal(blob[off:off + ln], want), "chunk %d bytes" % c
        off += ln
    assert off == total
    back = torch_cuda.zeros(n, dtype=torch_cuda.uint8, device="cuda")
    codec.range_decode_device(out.data_ptr(), lens, n, chunk, back.data_ptr())
    assert np.array_equal(back.cpu().numpy(), sym)


@pytest.mark.parametrize("pack", ["0", "1"])
@pytest.mark.parametrize("form", ["full", "compact"])
def test_encoder_table_forms_are_bit_identical(codec, torch_cuda, oracle, monkeypatch, form, pack):
    """the encoder's two table forms (32 KB full entries / 16.5 KB cumulative counts only, chosen by grid size) and its
    two ways of storing the raw entries (scattered 2-byte stores / packed groups of four): all must give the oracle's
    bytes; skewed, uniform and single-symbol data"""
    monkeypatch.setenv("WRB_ENC_TABLES", form)
    monkeypatch.setenv("WRB_ENC_PACK", pack)
    rng = np.random.default_rng(3)
    n = 59999 * 5 + 123
    for sym in (rng.integers(0, 256, n, dtype=np.uint8), np.clip(np.rint(128 + 0.7 * rng.standard_normal(n)), 0, 255).astype(np.uint8),
                np.full(n, 255, np.uint8), np.where(rng.random(n) < 0.001, 0, 255).astype(np.uint8)):
        d_sym = dev(torch_cuda, sym)
        out = torch_cuda.zeros(2 * n + 65536, dtype=torch_cuda.uint8, device="cuda")
        lens, total = codec.range_encode_device(d_sym.data_ptr(), n, L1, out.data_ptr(), out.numel())
        blob = out[:total].cpu().numpy()
        off = 0
        for c, ln in enumerate(lens):
            assert np.array_equal(blob[off:off + ln], oracle.range_encode(sym[c * L1:(c + 1) * L1])), "chunk %d" % c
            off += ln


def test_chunk_streams_accepted_by_reference_decoder(codec, torch_cuda, ref):
    """every chunk is a byte-valid stream for the reference's own range_decode()"""
    n = 200000
    sym = SYM_GENS["lcg"](n)
    d_sym = dev(torch_cuda, sym)
    out = torch_cuda.zeros(2 * n + 65536, dtype=torch_cuda.uint8, device="cuda")
    lens, total = codec.range_encode_device(d_sym.data_ptr(), n, L1, out.data_ptr(), out.numel())
    blob = out[:total].cpu().numpy()
    off = 0
    for c, ln in enumerate(lens):
        part = sym[c * L1:(c + 1) * L1]
        assert np.array_equal(ref.range_decode(blob[off:off + ln], part.size), part)
        assert np.array_equal(ref.range_encode(part), blob[off:off + ln])
        off += ln


def quantise(codec, torch, f, tol, wtflag=1, dtype=F64):
    nz, ny, nx = f.shape
    d_f = dev(torch, f.astype(np.float32) if dtype == F32 else f)
    coef = torch.zeros(f.size, dtype=torch.float64, device="cuda")
    sym = torch.zeros(8 * f.size, dtype=torch.uint8, device="cuda")
    h = codec.quantise_device(d_f.data_ptr(), dtype, nx, ny, nz, tol, wtflag, coef.data_ptr(), sym.data_ptr())
    return h, coef.cpu().numpy(), sym.cpu().numpy().reshape(8, f.size)[:h.nlay]


@pytest.mark.parametrize("shape,tol,wt", [((40, 40, 40), 1e-3, 1), ((20, 24, 32), 1e-6, 1), ((9, 30, 17), 1e-10, 1),
                                          ((16, 16, 16), 1e-4, 0), ((64, 64, 64), 1e-5, 1), ((48, 50, 77), 1e-16, 1),
                                          ((96, 96, 96), 1e-8, 1)])
def test_quantiser_vs_oracle(codec, torch_cuda, oracle, shape, tol, wt):
    f = oracle.probe_field(shape, seed=99 + shape[0], nm=16)
    want = oracle.encode(f, tol, wtflag=wt, want_symbols=True)
    hw = want["header"]
    h, coef, sym = quantise(codec, torch_cuda, f, tol, wt)
    assert bits_equal(coef.reshape(shape), oracle.wavelet3d(f, 4 if wt else 0))
    assert (h.wlev, h.nlay) == (hw.wlev, hw.nlay)
    assert bits_equal(np.array([h.tolabs, h.midval, h.halfspanval]), np.array([hw.tolabs, hw.midval, hw.halfspan]))
    assert bits_equal(np.array(list(h.deps_vec)[:h.nlay]), np.array(list(hw.deps)[:hw.nlay]))
    assert bits_equal(np.array(list(h.minval_vec)[:h.nlay]), np.array(list(hw.minval)[:hw.nlay]))
    assert np.array_equal(sym, want["symbols"])


def test_f32_input_is_widened_like_the_reference_front_end(codec, torch_cuda, oracle):
    # generic front-end widens element by element on the host (gen_aux.cpp:305-309)
    f32 = oracle.probe_field((48, 48, 48), seed=3).astype(np.float32)
    want = oracle.encode(f32.astype(np.float64), 1e-4, want_symbols=True)
    h, coef, sym = quantise(codec, torch_cuda, f32, 1e-4, 1, dtype=F32)
    assert h.nlay == want["header"].nlay and np.array_equal(sym, want["symbols"])
    assert bits_equal(np.array(list(h.deps_vec)[:h.nlay]), np.array(list(want["header"].deps)[:h.nlay]))


def encode_dev(codec, torch, f, tol, wt=1, dtype=F64):
    from waverange_b200 import api
    nz, ny, nx = f.shape
    d_f = dev(torch, f.astype(np.float32) if dtype == F32 else f)
    _, cap = api.setup_wr(nx, ny, nz)
    out = torch.zeros(cap + 64, dtype=torch.uint8, device="cuda")
    h = codec.encode_device(d_f.data_ptr(), dtype, nx, ny, nz, tol, out.data_ptr(), cap, wt)
    return h, out


@pytest.mark.parametrize("shape,tol", [((64, 64, 64), 1e-5), ((40, 50, 60), 1e-3), ((100, 100, 100), 1e-8),
                                       ((30, 1, 500), 1e-6), ((128, 128, 96), 1e-12)])
def test_encode_chunks_bit_exact_and_decode(codec, torch_cuda, oracle, shape, tol):
    from waverange_b200 import api
    f = oracle.probe_field(shape, seed=7 + shape[2], nm=20)
    h, out = encode_dev(codec, torch_cuda, f, tol)
    want = oracle.encode(f, tol, chunk_len=L1, want_symbols=True)
    hw = want["header"]
    assert h.nlay == hw.nlay
    blob = out[:h.ntot_enc].cpu().numpy()
    off, woff = 0, 0
    for l in range(h.nlay):
        layer = blob[off:off + h.len_enc_vec[l]]
        chunk_len, streams = api.parse_container(layer)
        assert chunk_len == min(L1, f.size) and len(streams) == want["chunk_lens"].shape[1]
        for c, s in enumerate(streams):
            n = int(want["chunk_lens"][l][c])
            assert s == want["data"][woff:woff + n].tobytes(), "layer %d chunk %d" % (l, c)
            woff += n
        off += h.len_enc_vec[l]
    # compression ratio within 1 % of the reference's whole-layer streams
    whole = oracle.encode(f, tol)
    assert h.ntot_enc <= 1.01 * whole["header"].ntot_enc
    # decode on the GPU: bit-identical to the reference decoder's reconstruction
    nz, ny, nx = shape
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
    codec.decode_device(rec.data_ptr(), F64, nx, ny, nz, h, out.data_ptr())
    want_rec = oracle.decode(shape, whole["header"], whole["data"])
    assert bits_equal(rec.cpu().numpy().reshape(shape), want_rec)
    assert np.abs(want_rec - f).max() <= tol * np.abs(f).max()


@pytest.mark.parametrize("tables", ["auto", "pair", "compact"])
@pytest.mark.parametrize("nseek", [-1, 0, 1, 3, 7])
def test_seek_points_change_neither_streams_nor_reconstruction(product_lib, torch_cuda, oracle, monkeypatch, nseek, tables):
    """Seek points (decoder entry points inside a chunk) only add table bytes: for every setting the chunk streams are the
    oracle's and the decoder -- running nseek+1 lanes per chunk -- returns the reference's reconstruction.  The field has a
    short last chunk (lanes without work, unused table entries).  -1 = the encoder's own choice, bounded to keep the
    container within 1 % of the reference's streams.  The decoder's two table forms (WRB_DEC_TABLES; by default chosen
    from the grid size) must both give the same symbols."""
    from waverange_b200 import api
    if tables != "auto":                                  # both table forms of the decoder at every lane count
        monkeypatch.setenv("WRB_DEC_TABLES", tables)
    shape, tol = (70, 64, 96), 1e-7                      # 430080 symbols: 7 full chunks + one of 10087
    f = oracle.probe_field(shape, seed=11, nm=20)
    c = api.Codec(device=0)
    c.set_seek_points(nseek)
    h, out = encode_dev(c, torch_cuda, f, tol)
    want = oracle.encode(f, tol, chunk_len=L1)
    whole = oracle.encode(f, tol)
    blob = out[:h.ntot_enc].cpu().numpy()
    off, woff, kept = 0, 0, set()
    for l in range(h.nlay):
        layer = blob[off:off + h.len_enc_vec[l]]
        assert bytes(layer[:4]) == b"WRCK" and int.from_bytes(bytes(layer[4:8]), "little") == 3
        kept.add(int.from_bytes(bytes(layer[28:32]), "little"))
        _, streams = api.parse_container(layer)
        for k, s in enumerate(streams):
            n = int(want["chunk_lens"][l][k])
            assert s == want["data"][woff:woff + n].tobytes(), "layer %d chunk %d" % (l, k)
            woff += n
        off += h.len_enc_vec[l]
    assert len(kept) == 1                                 # one geometry for all layers
    k = kept.pop()
    nchunks = want["chunk_lens"].shape[1]
    streams_total = int(want["chunk_lens"][:h.nlay].sum())
    assert h.ntot_enc == streams_total + h.nlay * (32 + nchunks * (4 + 10 * k))
    if nseek >= 0:
        assert k == nseek
    else:
        assert k in (0, 1, 3, 7) and h.ntot_enc <= 1.01 * whole["header"].ntot_enc
        assert h.nlay * nchunks * (4 + 10 * k) <= 0.0085 * streams_total
        if k < 7:
            assert h.nlay * nchunks * (4 + 10 * (2 * k + 1)) > 0.0085 * streams_total     # as many as the budget allows
    nz, ny, nx = shape
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
    c.decode_device(rec.data_ptr(), F64, nx, ny, nz, h, out.data_ptr())
    assert bits_equal(rec.cpu().numpy().reshape(shape), oracle.decode(shape, whole["header"], whole["data"]))
    c.close()


@pytest.mark.parametrize("key", ["e2e_a", "e2e_b", "e2e_c", "e2e_d"])
def test_single_stream_mode_is_byte_identical_to_reference(codec, torch_cuda, golden, key):
    """chunk_blocks = 0: data_enc, lengths and header doubles equal encoding_wrap() of the reference."""
    f = golden[key + "_in"]
    tol, wt = golden[key + "_tol"]
    codec.set_chunk_blocks(0)
    h, out = encode_dev(codec, torch_cuda, f, float(tol), int(wt))
    assert [h.wlev, h.nlay, h.ntot_enc] == list(golden[key + "_int"])
    assert bits_equal(np.array([h.tolabs, h.midval, h.halfspanval]), golden[key + "_scal"])
    assert bits_equal(np.array(list(h.deps_vec)[:h.nlay]), golden[key + "_deps"])
    assert bits_equal(np.array(list(h.minval_vec)[:h.nlay]), golden[key + "_minval"])
    assert list(h.len_enc_vec)[:h.nlay] == list(golden[key + "_len"])
    assert np.array_equal(out[:h.ntot_enc].cpu().numpy(), golden[key + "_data"])
    nz, ny, nx = f.shape
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
    codec.decode_device(rec.data_ptr(), F64, nx, ny, nz, h, out.data_ptr())
    assert np.array_equal(sha(rec.cpu().numpy()), golden[key + "_rec_sha"])


def test_constant_field_trivial_exit(codec, torch_cuda, golden):
    f = np.full((4, 5, 6), 3.25)
    h, out = encode_dev(codec, torch_cuda, f, 1e-6)
    assert [h.wlev, h.nlay, h.ntot_enc] == list(golden["e2e_const_int"])
    assert bits_equal(np.array([h.tolabs, h.midval, h.halfspanval]), golden["e2e_const_scal"])
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
    codec.decode_device(rec.data_ptr(), F64, 6, 5, 4, h, out.data_ptr())
    assert np.all(rec.cpu().numpy() == 3.25)


def test_reference_entry_points_on_host_buffers(product_lib, torch_cuda, oracle):
    """encoding_wrap / decoding_wrap with the reference's signatures (host arrays in, host arrays out)."""
    from waverange_b200 import api
    f = oracle.probe_field((50, 60, 70), seed=11)
    h, data = api.encoding_wrap(f, 1e-6)
    want = oracle.encode(f, 1e-6)
    hw = want["header"]
    assert (h.wlev, h.nlay) == (hw.wlev, hw.nlay)
    assert bits_equal(np.array(list(h.deps_vec)), np.array(list(hw.deps)))
    assert bits_equal(np.array(list(h.minval_vec)), np.array(list(hw.minval)))
    rec = api.decoding_wrap(f.shape, h, data)
    assert bits_equal(rec, oracle.decode(f.shape, hw, want["data"]))
    assert np.abs(rec - f).max() <= 1e-6 * np.abs(f).max()


def test_f32_roundtrip_host_api(codec, torch_cuda, oracle):
    f = oracle.probe_field((64, 64, 64), seed=21).astype(np.float32)
    h, data = codec.encode_host(f, 1e-4)
    rec = codec.decode_host(f.shape, h, data, dtype=np.float32)
    want = oracle.encode(f.astype(np.float64), 1e-4)
    want_rec = oracle.decode(f.shape, want["header"], want["data"]).astype(np.float32)
    assert bits_equal(rec, want_rec)
    assert np.abs(rec.astype(np.float64) - f).max() <= 1e-4 * np.abs(f).max()


def test_pageable_host_buffers_take_the_staged_path(product_lib, codec, torch_cuda, oracle):
    """ordinary (pageable) arrays, as a program written for the reference passes them, go through the pinned staging ring
    (three slabs here, so the ring wraps): same header, same bytes and same reconstruction as from pinned buffers"""
    from waverange_b200 import api
    shape, tol = (220, 200, 200), 1e-6                       # 70.4 MB as float64
    f = oracle.probe_field(shape, seed=3, nm=16)
    h1, d1 = api.encoding_wrap(f, tol)                        # reference entry point, pageable in and out
    pin = torch_cuda.empty(shape, dtype=torch_cuda.float64, pin_memory=True)
    pin.copy_(torch_cuda.from_numpy(f))
    pout = torch_cuda.empty(f.nbytes, dtype=torch_cuda.uint8, pin_memory=True)
    h2, d2 = codec.encode_host(pin.numpy(), tol, out=pout.numpy())
    assert header_tuple(h1) == header_tuple(h2) and h1.ntot_enc == h2.ntot_enc
    assert np.array_equal(np.asarray(d1), np.asarray(d2)[:h2.ntot_enc])
    r1 = api.decoding_wrap(shape, h1, d1)                     # pageable out (first touched by the copy threads)
    prec = torch_cuda.empty(shape, dtype=torch_cuda.float64, pin_memory=True)
    codec.decode_host(shape, h2, d2, out=prec.numpy())
    assert bits_equal(r1, prec.numpy())
    assert np.abs(r1 - f).max() <= tol * np.abs(f).max()
    f32 = f.astype(np.float32)                                # pageable float32 through the handle API
    h3, d3 = codec.encode_host(f32, tol)
    r3 = codec.decode_host(shape, h3, d3, dtype=np.float32)
    assert np.abs(r3.astype(np.float64) - f32).max() <= tol * np.abs(f32).max()


def test_overflow_is_reported(codec, torch_cuda, oracle):
    from waverange_b200 import api
    f = np.random.default_rng(1).standard_normal((32, 32, 32))
    d_f = dev(torch_cuda, f)
    out = torch_cuda.zeros(4096, dtype=torch_cuda.uint8, device="cuda")
    with pytest.raises(api.WaveRangeError):
        codec.encode_device(d_f.data_ptr(), F64, 32, 32, 32, 1e-12, out.data_ptr(), 1000)


def test_corrupt_container_is_rejected(codec, torch_cuda, oracle):
    from waverange_b200 import api
    f = oracle.probe_field((32, 32, 32), seed=2)
    h, out = encode_dev(codec, torch_cuda, f, 1e-4)
    out[0] = 0x55
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
    with pytest.raises(api.WaveRangeError):
        codec.decode_device(rec.data_ptr(), F64, 32, 32, 32, h, out.data_ptr())


def _put(out, off, value, nbytes):
    import torch
    out[off:off + nbytes] = torch.tensor(list(int(value).to_bytes(nbytes, "little")), dtype=torch.uint8, device=out.device)


@pytest.mark.parametrize("case", ["second_layer_magic", "nchunks_huge", "chunk_len_tiny", "len_table_garbage", "seek_pos_far",
                                  "truncated_layers", "stream_garbage", "len_sum_shifted"])
def test_malformed_containers_fail_cleanly(codec, torch_cuda, oracle, case):
    """A corrupt or truncated container must come back as an error (WRB_E_FORMAT) -- never as an out-of-bounds read: the
    blob sits at the very end of its allocation (no slack beyond the 64 bytes the interface asks for), and the handle
    must still work afterwards (a fault would leave a sticky context error)."""
    from waverange_b200 import api
    torch = torch_cuda
    f = oracle.probe_field((80, 96, 112), seed=2, nm=14)          # 860160 points: 15 chunks, seek points granted
    nz, ny, nx = f.shape
    h, out = encode_dev(codec, torch, f, 1e-5)
    assert h.nlay >= 2
    good = out[:h.ntot_enc + 64].clone()
    blob = good.clone()
    l1 = h.len_enc_vec[0]                                          # offset of the second layer
    nch = (f.size + L1 - 1) // L1
    nseek = int.from_bytes(blob[28:32].cpu().numpy().tobytes(), "little")
    h2 = api.Header.from_buffer_copy(bytes(h))
    if case == "second_layer_magic":
        blob[l1] = 0x55
    elif case == "nchunks_huge":
        _put(blob, 24, 0x7FFFFFF0, 4)
    elif case == "chunk_len_tiny":
        _put(blob, 8, 3, 8); _put(blob, 24, (f.size + 2) // 3, 4)
    elif case == "len_table_garbage":
        _put(blob, 32 + 4 * 3, 0xFFFFFFF0, 4)
    elif case == "seek_pos_far":
        assert nseek > 0
        for k in range(nseek):
            _put(blob, 32 + 4 * nch + 10 * (nseek * 2 + k) + 8, 0xFFFF, 2)
    elif case == "truncated_layers":
        h2.len_enc_vec[h.nlay - 1] = 40; h2.ntot_enc = sum(h2.len_enc_vec[:h.nlay])
        blob = blob[:h2.ntot_enc + 64].clone()
    elif case == "stream_garbage":
        a = 32 + 4 * nch + 10 * nseek * nch + 700
        blob[a:a + 4000] = torch.randint(0, 256, (4000,), dtype=torch.uint8, device="cuda")
    elif case == "len_sum_shifted":
        v = int.from_bytes(blob[32:36].cpu().numpy().tobytes(), "little")
        _put(blob, 32, v + 5, 4); w = int.from_bytes(blob[36:40].cpu().numpy().tobytes(), "little"); _put(blob, 36, w - 5, 4)
    rec = torch.zeros(f.size, dtype=torch.float64, device="cuda")
    if case in ("stream_garbage", "len_sum_shifted"):
        # random bytes inside a stream / a shifted stream start may still decode to *something*: what matters is that nothing faults
        try:
            codec.decode_device(rec.data_ptr(), F64, nx, ny, nz, h2, blob.data_ptr())
        except api.WaveRangeError:
            pass
    else:
        with pytest.raises(api.WaveRangeError):
            codec.decode_device(rec.data_ptr(), F64, nx, ny, nz, h2, blob.data_ptr())
    torch.cuda.synchronize()                                       # no sticky error
    codec.decode_device(rec.data_ptr(), F64, nx, ny, nz, h, good.data_ptr())
    whole = oracle.encode(f, 1e-5)
    assert bits_equal(rec.cpu().numpy().reshape(f.shape), oracle.decode(f.shape, whole["header"], whole["data"]))


def test_codec_handles_on_one_device_from_threads(torch_cuda, oracle):
    """several handles driven from several host threads: the one-time kernel attribute setup is per device and guarded"""
    import threading
    from waverange_b200 import api
    f = oracle.probe_field((40, 48, 56), seed=4, nm=10)
    want = oracle.encode(f, 1e-6, chunk_len=L1)
    errs = []

    def work():
        try:
            c = api.Codec(device=0)
            h, out = encode_dev(c, torch_cuda, f, 1e-6)
            assert h.ntot_enc > 0 and h.nlay == want["header"].nlay
            rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
            c.decode_device(rec.data_ptr(), F64, 56, 48, 40, h, out.data_ptr())
            c.close()
        except Exception as e:      # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=work) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_layer_count_guess_falls_back_to_all_layers(codec, torch_cuda, oracle, monkeypatch):
    """the encoder launches only as many layers as the tolerance can need (codec.cu layer_guess); when the guess is
    too small the `done` flag of the device state is clear and the call repeats with all eight: same bytes"""
    f = oracle.probe_field((48, 40, 56), seed=11, nm=12)
    tol = 1e-9
    h0, out0 = encode_dev(codec, torch_cuda, f, tol)
    assert h0.nlay >= 4
    assert codec.layer_guess_misses() == 0
    monkeypatch.setenv("WRB_LAYER_GUESS", "2")
    h1, out1 = encode_dev(codec, torch_cuda, f, tol)
    assert codec.layer_guess_misses() == 1
    assert h1.nlay == h0.nlay and h1.ntot_enc == h0.ntot_enc
    assert torch_cuda.equal(out0[:h0.ntot_enc], out1[:h1.ntot_enc])


@pytest.mark.parametrize("wt", [0, 1])
@pytest.mark.parametrize("shape,grid", [((12, 10, 14), (2, 2, 2)), ((40, 33, 50), (3, 2, 4)), ((70, 64, 64), (1, 1, 5))])
def test_local_cutoff_branch(product_lib, codec, torch_cuda, oracle, wt, shape, grid):
    """encoding_wrap(mx*my*mz > 1): the drop-in entry point, stock layout, against the oracle's bytes; and the chunked
    layout through the codec handle, decoded on the GPU, against the oracle's reconstruction"""
    from waverange_b200 import api
    rng = np.random.default_rng(shape[0] + grid[2] + wt)
    f = oracle.probe_field(shape, seed=31 + wt, nm=14)
    vals = 10.0 ** rng.uniform(-6, -2, size=grid[0] * grid[1] * grid[2])
    want = oracle.encode(f, 0.0, wtflag=wt, cutoff=(*grid, vals))
    hw = want["header"]
    nz, ny, nx = shape
    d_f = torch_cuda.from_numpy(f).cuda()
    _, cap = api.setup_wr(nx, ny, nz)
    blob = torch_cuda.zeros(cap + 64, dtype=torch_cuda.uint8, device="cuda")
    c0 = api.Codec(device=0, chunk_blocks=0)                      # stock layout: the reference's bytes
    c0.set_local_cutoff(*grid, vals)
    h0 = c0.encode_device(d_f.data_ptr(), F64, nx, ny, nz, 0.0, blob.data_ptr(), cap, wt)
    c0.close()
    assert (h0.nlay, h0.ntot_enc, list(h0.len_enc_vec)[:h0.nlay]) == (hw.nlay, hw.ntot_enc, list(hw.len)[:hw.nlay])
    assert blob[:h0.ntot_enc].cpu().numpy().tobytes() == want["data"].tobytes()
    h, data = api.encoding_wrap(f, 0.0, wtflag=wt, cutoff=(*grid, vals))          # the drop-in entry point
    assert (h.nlay, h.tolabs, list(h.deps_vec)[:h.nlay], list(h.minval_vec)[:h.nlay]) == \
           (hw.nlay, hw.tolabs, list(hw.deps)[:hw.nlay], list(hw.minval)[:hw.nlay])
    assert bits_equal(api.decoding_wrap(shape, h, data), oracle.decode(shape, hw, want["data"]))
    codec.set_local_cutoff(*grid, vals)
    h2 = codec.encode_device(d_f.data_ptr(), F64, nx, ny, nz, 123.0, blob.data_ptr(), cap, wt)     # tolrel is ignored
    codec.set_local_cutoff()
    assert (h2.nlay, h2.tolabs, list(h2.deps_vec)[:h2.nlay]) == (hw.nlay, hw.tolabs, list(hw.deps)[:hw.nlay])
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float64, device="cuda")
    codec.decode_device(rec.data_ptr(), F64, nx, ny, nz, h2, blob.data_ptr())
    assert bits_equal(rec.cpu().numpy().reshape(shape), oracle.decode(shape, hw, want["data"]))


def _random_cases(n, seed):
    """shapes that hit every dispatch of the transform: all levels fused (multiples of 16, >= 64 / 128), some levels
    fused and the coarser ones general (even but not multiples of 16), odd and tiny extents, 2-D and 1-D fields"""
    rng = np.random.default_rng(seed)
    pools = [[64, 80, 96, 128], [24, 40, 56, 72, 88, 100, 36], [9, 17, 31, 33, 50, 63, 7, 5], [1, 2, 3]]
    out = []
    for k in range(n):
        kind = k % 5
        if kind == 0:
            shape = tuple(int(rng.choice(pools[0])) for _ in range(3))
        elif kind == 1:
            shape = tuple(int(rng.choice(pools[1])) for _ in range(3))
        elif kind == 2:
            shape = tuple(int(rng.choice(pools[2])) for _ in range(3))
        elif kind == 3:
            shape = (int(rng.choice(pools[3])), int(rng.choice(pools[0] + pools[1])), int(rng.choice(pools[1] + pools[2])))
        else:
            shape = tuple(int(rng.choice(pools[int(rng.integers(0, 3))])) for _ in range(3))
        tol = float(10.0 ** rng.uniform(-12, -2))
        out.append((shape, tol, int(rng.integers(0, 2)), int(rng.integers(0, 5) > 0)))
    return out


@pytest.mark.parametrize("shape,tol,f32,wt", _random_cases(40, 2026))
def test_randomised_differential(codec, torch_cuda, oracle, shape, tol, f32, wt):
    """whole pipeline against the oracle on random shapes / tolerances / precisions: header doubles, every chunk
    stream, the stock-layout bytes and both decoders' reconstructions, bit for bit"""
    from waverange_b200 import api
    f = oracle.probe_field(shape, seed=sum(shape) + int(-np.log10(tol)), nm=12, round_f32=bool(f32))
    nz, ny, nx = shape
    dt = F32 if f32 else F64
    want = oracle.encode(f, tol, wtflag=wt, chunk_len=L1)
    whole = oracle.encode(f, tol, wtflag=wt)
    hw = want["header"]
    fin = f.astype(np.float32) if f32 else f
    h, out = encode_dev(codec, torch_cuda, fin, tol, wt, dtype=dt)
    assert (h.nlay, h.wlev) == (hw.nlay, hw.wlev)
    assert bits_equal(np.array([h.tolabs, h.midval, h.halfspanval]), np.array([hw.tolabs, hw.midval, hw.halfspan]))
    assert list(h.deps_vec)[:h.nlay] == list(hw.deps)[:h.nlay] and list(h.minval_vec)[:h.nlay] == list(hw.minval)[:h.nlay]
    blob = out[:h.ntot_enc].cpu().numpy()
    off, woff = 0, 0
    for l in range(h.nlay):
        _, streams = api.parse_container(blob[off:off + h.len_enc_vec[l]])
        n = int(want["chunk_lens"][l].sum())
        assert b"".join(streams) == want["data"][woff:woff + n].tobytes(), "layer %d" % l
        woff += n
        off += h.len_enc_vec[l]
    rec = torch_cuda.zeros(f.size, dtype=torch_cuda.float32 if f32 else torch_cuda.float64, device="cuda")
    codec.decode_device(rec.data_ptr(), dt, nx, ny, nz, h, out.data_ptr())
    want_rec = oracle.decode(shape, whole["header"], whole["data"])
    if f32:
        want_rec = want_rec.astype(np.float32)
    assert bits_equal(rec.cpu().numpy().reshape(shape), want_rec)
    # stock layout: the reference's own bytes, and its streams decode on the GPU
    c0 = api.Codec(device=0, chunk_blocks=0)
    h0, out0 = encode_dev(c0, torch_cuda, fin, tol, wt, dtype=dt)
    assert h0.ntot_enc == whole["header"].ntot_enc and out0[:h0.ntot_enc].cpu().numpy().tobytes() == whole["data"].tobytes()
    rec.zero_()
    c0.decode_device(rec.data_ptr(), dt, nx, ny, nz, h0, out0.data_ptr())
    c0.close()
    assert bits_equal(rec.cpu().numpy().reshape(shape), want_rec)
