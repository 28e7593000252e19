"""ctypes bindings for the parity oracle -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by waverange_b200/.

Two libraries:
  * Restatement  -- oracle/libwr_oracle.so   (oracle/wr_oracle.c, always buildable)
  * Reference    -- oracle/_ref/libwaverange_ref_{strict,fma}.so: the UNMODIFIED
                    reference compiled by oracle/build_oracle.py where /root/reference
                    exists; the binaries travel to the GPU box.
"""
import contextlib
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NLAYMAX = 8
BLOCK = 60000

u8p = C.POINTER(C.c_uint8)
f64p = C.POINTER(C.c_double)
u32p = C.POINTER(C.c_uint32)


class Header(C.Structure):
    """mirror of wro_header (oracle/wr_oracle.c)"""
    _fields_ = [("tolabs", C.c_double), ("midval", C.c_double), ("halfspan", C.c_double),
                ("wlev", C.c_uint8), ("nlay", C.c_uint8), ("ntot_enc", C.c_uint64),
                ("deps", C.c_double * NLAYMAX), ("minval", C.c_double * NLAYMAX),
                ("len", C.c_uint64 * NLAYMAX)]


def _p(a, t):
    return a.ctypes.data_as(t)


@contextlib.contextmanager
def quiet_stdout():
    """The reference prints progress from inside the library (wrappers.cpp:232,...)."""
    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(null)
        os.close(saved)


class Restatement:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "libwr_oracle.so")
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build_restatement()
        L = self.lib = C.CDLL(path)
        L.wro_wavelet3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f64p]
        L.wro_range_bound.restype = C.c_size_t
        L.wro_range_bound.argtypes = [C.c_size_t]
        L.wro_range_encode.restype = C.c_size_t
        L.wro_range_encode.argtypes = [u8p, C.c_size_t, u8p]
        L.wro_range_decode.restype = C.c_size_t
        L.wro_range_decode.argtypes = [u8p, C.c_size_t, u8p, C.c_size_t]
        L.wro_encode.restype = C.c_int
        L.wro_encode_cutoff.restype = C.c_int
        L.wro_encode_cutoff.argtypes = [C.c_int, C.c_int, C.c_int, f64p, C.c_int, C.c_int, C.c_int, C.c_int, f64p,
                                        C.c_uint64, C.POINTER(Header), u8p, C.c_uint64, u8p, u32p]
        L.wro_encode.argtypes = [C.c_int, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_uint64,
                                 C.POINTER(Header), u8p, C.c_uint64, u8p, u32p]
        L.wro_decode.restype = C.c_int
        L.wro_decode.argtypes = [C.c_int, C.c_int, C.c_int, f64p, C.POINTER(Header), u8p, C.c_uint64, u32p]
        L.wro_probe_field.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_double, C.c_int, f64p]
        L.wro_fnv1a.restype = C.c_uint64
        L.wro_fnv1a.argtypes = [u8p, C.c_size_t]
        L.wro_ind_p2w.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int)] * 4
        u64p = C.POINTER(C.c_uint64)
        L.wro_wavelet3d_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f64p]
        L.wro_encode_digest.restype = C.c_int
        L.wro_encode_digest.argtypes = [C.c_int, C.c_int, C.c_int, f64p, C.c_int, C.c_double, C.c_uint64,
                                        C.POINTER(Header), u32p, u64p, u64p, C.c_longlong, u8p]
        L.wro_fnv1a_many.argtypes = [u8p, u64p, u64p, C.c_size_t, u64p]
        L.wro_decode_symbols_mt.argtypes = [C.c_int, C.c_int, C.c_int, f64p, C.POINTER(Header), u8p]

    # -- wavelet ---------------------------------------------------------
    def wavelet3d(self, a, lvl):
        """a: float64 array shaped (nz, ny, nx) (x fastest); returns transformed copy."""
        a = np.ascontiguousarray(a, dtype=np.float64).copy()
        nz, ny, nx = a.shape
        self.lib.wro_wavelet3d(nx, ny, nz, lvl, _p(a, f64p))
        return a

    # -- range coder -----------------------------------------------------
    def range_encode(self, sym):
        sym = np.ascontiguousarray(sym, dtype=np.uint8)
        out = np.empty(self.lib.wro_range_bound(sym.size), dtype=np.uint8)
        n = self.lib.wro_range_encode(_p(sym, u8p), sym.size, _p(out, u8p))
        return out[:n].copy()

    def range_decode(self, stream, cap):
        buf = np.zeros(len(stream) + 16, dtype=np.uint8)
        buf[:len(stream)] = np.frombuffer(bytes(stream), dtype=np.uint8)
        sym = np.empty(cap, dtype=np.uint8)
        n = self.lib.wro_range_decode(_p(buf, u8p), len(stream), _p(sym, u8p), cap)
        return sym[:min(n, cap)].copy(), n

    # -- whole path ------------------------------------------------------
    def encode(self, fld, tol, wtflag=1, chunk_len=0, want_symbols=False, cutoff=None):
        """fld: float64 (nz,ny,nx).  Returns dict(header, data, symbols, chunk_lens, residual).
        cutoff = (mx, my, mz, values): the local-precision grid of encoding_wrap (tol is then ignored)."""
        a = np.ascontiguousarray(fld, dtype=np.float64).copy()
        nz, ny, nx = a.shape
        ntot = a.size
        cap = 8 * max(1024, ntot) + 2 * 1024 * 1024
        data = np.zeros(cap + 16, dtype=np.uint8)
        hdr = Header()
        sym = np.empty(NLAYMAX * ntot, dtype=np.uint8) if want_symbols else None
        nch = (ntot + chunk_len - 1) // chunk_len if chunk_len else 1
        cl = np.zeros(NLAYMAX * nch, dtype=np.uint32)
        if cutoff is None:
            rc = self.lib.wro_encode(nx, ny, nz, _p(a, f64p), wtflag, tol, chunk_len, C.byref(hdr),
                                     _p(data, u8p), cap, _p(sym, u8p) if want_symbols else None, _p(cl, u32p))
        else:
            cmx, cmy, cmz, vals = cutoff
            cv = np.ascontiguousarray(vals, dtype=np.float64)
            assert cv.size == cmx * cmy * cmz
            rc = self.lib.wro_encode_cutoff(nx, ny, nz, _p(a, f64p), wtflag, cmx, cmy, cmz, _p(cv, f64p), chunk_len,
                                            C.byref(hdr), _p(data, u8p), cap, _p(sym, u8p) if want_symbols else None,
                                            _p(cl, u32p))
        if rc != 0:
            raise RuntimeError("oracle encode overflow")
        nlay = hdr.nlay
        return dict(header=hdr, data=data[:hdr.ntot_enc].copy(),
                    symbols=sym[:nlay * ntot].reshape(nlay, ntot).copy() if want_symbols else None,
                    chunk_lens=cl[:nlay * nch].reshape(nlay, nch).copy() if nlay else cl[:0],
                    residual=a)

    def encode_digest(self, fld, tol, chunk_len, wtflag=1, sample_chunk=-1, inplace=False):
        """The encode of a LARGE field, multi-threaded, keeping only a digest (oracle/wr_oracle.c wro_encode_digest):
        dict(header, chunk_lens[nlay, nch], stream_hash[nlay, nch], symbol_hash[nlay, nch], sample[nlay, chunk_len]).
        inplace: fld (float64, C-contiguous) is clobbered instead of copied."""
        a = fld if inplace else np.ascontiguousarray(fld, dtype=np.float64).copy()
        assert a.dtype == np.float64 and a.flags.c_contiguous
        nz, ny, nx = a.shape
        nch = (a.size + chunk_len - 1) // chunk_len
        cl = np.zeros(NLAYMAX * nch, dtype=np.uint32)
        sh = np.zeros(NLAYMAX * nch, dtype=np.uint64)
        qh = np.zeros(NLAYMAX * nch, dtype=np.uint64)
        sample = np.zeros(NLAYMAX * chunk_len, dtype=np.uint8) if sample_chunk >= 0 else None
        hdr = Header()
        u64p = C.POINTER(C.c_uint64)
        rc = self.lib.wro_encode_digest(nx, ny, nz, _p(a, f64p), wtflag, tol, chunk_len, C.byref(hdr), _p(cl, u32p),
                                        _p(sh, u64p), _p(qh, u64p), sample_chunk,
                                        _p(sample, u8p) if sample is not None else None)
        if rc != 0:
            raise RuntimeError("oracle digest encode failed (%d)" % rc)
        n = hdr.nlay
        return dict(header=hdr, chunk_lens=cl[:n * nch].reshape(n, nch), stream_hash=sh[:n * nch].reshape(n, nch),
                    symbol_hash=qh[:n * nch].reshape(n, nch),
                    sample=sample[:n * chunk_len].reshape(n, chunk_len) if sample is not None else None)

    def fnv1a_many(self, data, offs, lens):
        """FNV-1a of every byte range data[offs[i] : offs[i] + lens[i]] (multi-threaded)"""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint64)
        assert offs.size == lens.size and (offs.size == 0 or int((offs + lens).max()) <= data.size)
        out = np.zeros(offs.size, dtype=np.uint64)
        u64p = C.POINTER(C.c_uint64)
        self.lib.wro_fnv1a_many(_p(data, u8p), _p(offs, u64p), _p(lens, u64p), offs.size, _p(out, u64p))
        return out

    def decode_symbols(self, shape, hdr, sym):
        """accumulate + inverse transform of given symbol planes (nlay, ntot), multi-threaded"""
        nz, ny, nx = shape
        sym = np.ascontiguousarray(sym, dtype=np.uint8)
        assert sym.size == hdr.nlay * nx * ny * nz
        out = np.empty(shape, dtype=np.float64)
        self.lib.wro_decode_symbols_mt(nx, ny, nz, _p(out, f64p), C.byref(hdr), _p(sym, u8p))
        return out

    def decode(self, shape, hdr, data, chunk_len=0, chunk_lens=None):
        nz, ny, nx = shape
        out = np.empty(shape, dtype=np.float64)
        buf = np.zeros(len(data) + 16, dtype=np.uint8)
        buf[:len(data)] = data
        cl = None
        if chunk_len:
            cl = np.ascontiguousarray(np.asarray(chunk_lens).reshape(-1), dtype=np.uint32)
        self.lib.wro_decode(nx, ny, nz, _p(out, f64p), C.byref(hdr), _p(buf, u8p), chunk_len,
                            _p(cl, u32p) if cl is not None else None)
        return out

    def probe_field(self, shape, seed=12345, nm=24, expo=-5.0 / 12.0, round_f32=False):
        nz, ny, nx = shape
        out = np.empty(shape, dtype=np.float64)
        self.lib.wro_probe_field(nx, ny, nz, seed, nm, expo, int(round_f32), _p(out, f64p))
        return out

    def fnv1a(self, b):
        b = np.ascontiguousarray(np.frombuffer(bytes(b), dtype=np.uint8))
        return self.lib.wro_fnv1a(_p(b, u8p), b.size)

    def ind_p2w(self, lvl, n, i):
        o = [C.c_int() for _ in range(4)]
        self.lib.wro_ind_p2w(lvl, n[0], n[1], n[2], i[0], i[1], i[2], *[C.byref(x) for x in o])
        return tuple(x.value for x in o)


class Reference:
    """The unmodified reference library (oracle/_ref).  variant: 'strict' | 'fma'."""

    def __init__(self, variant="strict"):
        path = os.path.join(HERE, "_ref", "libwaverange_ref_%s.so" % variant)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.variant = variant
        L = self.lib = C.CDLL(path)
        ul = C.c_ulong
        L.waveletcdf97_3d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, f64p]
        L.ref_range_encode.argtypes = [u8p, ul, u8p, C.POINTER(ul)]
        L.ref_range_decode.argtypes = [u8p, ul, u8p, ul]
        L.encoding_wrap.argtypes = [C.c_int, C.c_int, C.c_int, f64p, C.c_int, C.c_int, C.c_int, C.c_int, f64p,
                                    f64p, f64p, f64p, u8p, u8p, C.POINTER(ul), f64p, f64p, C.POINTER(ul), u8p]
        L.decoding_wrap.argtypes = [C.c_int, C.c_int, C.c_int, f64p, f64p, f64p, f64p, u8p, u8p, C.POINTER(ul),
                                    f64p, f64p, C.POINTER(ul), u8p]
        L.setup_wr.argtypes = [C.c_int, C.c_int, C.c_int, u8p, C.POINTER(ul)]
        L.ind_p2w_3d.argtypes = [C.c_int] * 7 + [C.POINTER(C.c_int)] * 4

    @staticmethod
    def available(variant="strict"):
        return os.path.exists(os.path.join(HERE, "_ref", "libwaverange_ref_%s.so" % variant))

    def wavelet3d(self, a, lvl):
        a = np.ascontiguousarray(a, dtype=np.float64).copy()
        nz, ny, nx = a.shape
        self.lib.waveletcdf97_3d(nx, ny, nz, lvl, _p(a, f64p))
        return a

    def range_encode(self, sym):
        sym = np.ascontiguousarray(sym, dtype=np.uint8)
        padded = np.zeros(sym.size + 8, dtype=np.uint8)      # reference reads sym[n] (wrappers.cpp:89)
        padded[:sym.size] = sym
        out = np.zeros(2 * sym.size + 1024 * (sym.size // BLOCK + 2) + 4096, dtype=np.uint8)
        n = C.c_ulong(0)
        self.lib.ref_range_encode(_p(padded, u8p), sym.size, _p(out, u8p), C.byref(n))
        return out[:n.value].copy()

    def range_decode(self, stream, nsym):
        buf = np.zeros(len(stream) + 16, dtype=np.uint8)
        buf[:len(stream)] = np.frombuffer(bytes(stream), dtype=np.uint8)
        sym = np.zeros(nsym + BLOCK + 16, dtype=np.uint8)    # reference does not bound its writes
        self.lib.ref_range_decode(_p(buf, u8p), len(stream) + 8, _p(sym, u8p), nsym)
        return sym[:nsym].copy()

    def encode(self, fld, tol, wtflag=1, cutoff=None):
        a = np.ascontiguousarray(fld, dtype=np.float64).copy()
        nz, ny, nx = a.shape
        ntot = a.size
        nlaymax = C.c_uint8()
        cap = C.c_ulong()
        self.lib.setup_wr(nx, ny, nz, C.byref(nlaymax), C.byref(cap))
        data = np.zeros(cap.value + 16, dtype=np.uint8)
        cut = np.array([tol], dtype=np.float64)
        cmx = cmy = cmz = 1
        if cutoff is not None:
            cmx, cmy, cmz, vals = cutoff
            cut = np.ascontiguousarray(vals, dtype=np.float64)
        tolabs, mid, half = C.c_double(), C.c_double(), C.c_double()
        wlev, nlay = C.c_uint8(), C.c_uint8()
        ntot_enc = C.c_ulong()
        deps = np.zeros(NLAYMAX)
        minv = np.zeros(NLAYMAX)
        lens = (C.c_ulong * NLAYMAX)()
        with quiet_stdout():
            self.lib.encoding_wrap(nx, ny, nz, _p(a, f64p), wtflag, cmx, cmy, cmz, _p(cut, f64p),
                                   C.cast(C.byref(tolabs), f64p), C.cast(C.byref(mid), f64p),
                                   C.cast(C.byref(half), f64p), C.cast(C.byref(wlev), u8p),
                                   C.cast(C.byref(nlay), u8p), C.byref(ntot_enc), _p(deps, f64p), _p(minv, f64p),
                                   lens, _p(data, u8p))
        hdr = Header()
        hdr.tolabs, hdr.midval, hdr.halfspan = tolabs.value, mid.value, half.value
        hdr.wlev, hdr.nlay, hdr.ntot_enc = wlev.value, nlay.value, ntot_enc.value
        for i in range(nlay.value):
            hdr.deps[i], hdr.minval[i], hdr.len[i] = deps[i], minv[i], lens[i]
        return dict(header=hdr, data=data[:ntot_enc.value].copy(), residual=a)

    def decode(self, shape, hdr, data):
        nz, ny, nx = shape
        out = np.zeros(shape, dtype=np.float64)
        buf = np.zeros(len(data) + 16, dtype=np.uint8)
        buf[:len(data)] = data
        tolabs, mid, half = C.c_double(hdr.tolabs), C.c_double(hdr.midval), C.c_double(hdr.halfspan)
        wlev, nlay = C.c_uint8(hdr.wlev), C.c_uint8(hdr.nlay)
        ntot_enc = C.c_ulong(hdr.ntot_enc)
        deps = np.array(list(hdr.deps), dtype=np.float64)
        minv = np.array(list(hdr.minval), dtype=np.float64)
        lens = (C.c_ulong * NLAYMAX)(*list(hdr.len))
        with quiet_stdout():
            self.lib.decoding_wrap(nx, ny, nz, _p(out, f64p), C.cast(C.byref(tolabs), f64p),
                                   C.cast(C.byref(mid), f64p), C.cast(C.byref(half), f64p),
                                   C.cast(C.byref(wlev), u8p), C.cast(C.byref(nlay), u8p), C.byref(ntot_enc),
                                   _p(deps, f64p), _p(minv, f64p), lens, _p(buf, u8p))
        return out

    def ind_p2w(self, lvl, n, i):
        o = [C.c_int() for _ in range(4)]
        self.lib.ind_p2w_3d(lvl, n[0], n[1], n[2], i[0], i[1], i[2], *[C.byref(x) for x in o])
        return tuple(x.value for x in o)
