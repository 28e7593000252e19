"""Host-side mirror of the WaveRange library interface, bound to libwaverange_b200.so via ctypes.

The product is the shared library (CUDA kernels + C ABI, see include/waverange_b200.h and
include/waverange.h); this module only passes pointers.  It never falls back to a CPU
implementation: if the library has not been built, importing the bindings raises.

Two levels, as in the C ABI:
  * encoding_wrap / decoding_wrap / setup_wr -- the reference's entry points
    (reference src/core/wrappers.h:53,70,75) on HOST numpy arrays;
  * Codec -- handle-based device-pointer API (what bench.py times) and the stage-level entry
    points the parity tests use.  Device buffers are passed as integers (e.g. torch
    tensor.data_ptr()), so nothing here depends on torch.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwaverange_b200.so")
NLAYMAX = 8
BLOCKSIZE = 60000
F64, F32 = 0, 1


class Header(C.Structure):
    """wrb_header (include/waverange_b200.h): the coding metadata encoding_wrap() returns."""
    _fields_ = [("tolabs", C.c_double), ("midval", C.c_double), ("halfspanval", C.c_double),
                ("wlev", C.c_ubyte), ("nlay", C.c_ubyte), ("ntot_enc", C.c_ulong),
                ("deps_vec", C.c_double * NLAYMAX), ("minval_vec", C.c_double * NLAYMAX),
                ("len_enc_vec", C.c_ulong * NLAYMAX)]

    def as_dict(self):
        n = self.nlay
        return dict(tolabs=self.tolabs, midval=self.midval, halfspanval=self.halfspanval, wlev=self.wlev,
                    nlay=n, ntot_enc=self.ntot_enc, deps_vec=list(self.deps_vec)[:n],
                    minval_vec=list(self.minval_vec)[:n], len_enc_vec=list(self.len_enc_vec)[:n])


class FieldDesc(C.Structure):
    """wrb_field_desc (include/waverange_files.h): per-field parameters of the generic front-end."""
    _fields_ = [("nbytes", C.c_int), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nh", C.c_int),
                ("idinv", C.c_int), ("icomp", C.c_int), ("tol_base", C.c_double)]


class FieldRecord(C.Structure):
    """wrb_field_record: one field record of a .wrh header file."""
    _fields_ = [("desc", FieldDesc), ("recl", C.c_ubyte * 8), ("hdr", Header)]


class MssgCtl(C.Structure):
    """wrb_mssg_ctl (include/waverange_mssg.h): what the MSSG front-end takes from a GrADS .ctl file."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nt", C.c_int), ("undef", C.c_double), ("dset", C.c_char * 256)]


class MssgNmlst(C.Structure):
    """wrb_mssg_nmlst: grid, process grid and record table of an MSSG restart .nmlst file."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("nprocx", C.c_int), ("nprocy", C.c_int), ("ndset", C.c_int),
                ("dset", (C.c_char * 256) * 50)]


class WaveRangeError(RuntimeError):
    pass


# callbacks of the z-slab partition (include/waverange_b200.h: wrb_halo_fn, wrb_reduce_fn)
HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong)
REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)


_lib = None

# every symbol include/waverange_b200.h and include/waverange.h declare
EXPORTS = ["wrb_create", "wrb_destroy", "wrb_last_error", "wrb_set_stream", "wrb_set_chunk_blocks", "wrb_set_seek_points",
           "wrb_set_local_cutoff", "wrb_launch_count", "wrb_trim", "wrb_current_device", "wrb_layer_guess_misses", "wrb_setup", "wrb_encode_device", "wrb_decode_device",
           "wrb_encode_host", "wrb_decode_host", "wrb_decode_symbols_device", "wrb_decode_slab_symbols_device", "wrb_set_slab", "wrb_comm_unique_id", "wrb_set_comm",
           "wrb_set_slab_peers", "wrb_set_slab_order", "wrb_comm_counters", "wrb_slab_order_plane", "wrb_slab_chunk_range", "wrb_encode_slab_device", "wrb_decode_slab_device", "wrb_quantise_slab_device", "wrb_wavelet3d_device", "wrb_quantise_device",
           "wrb_range_encode_device", "wrb_range_decode_device", "wrb_ind_p2w_3d", "wrb_set_timing",
           "wrb_last_stage_ms",
           "wrb_wrh_begin", "wrb_wrh_append", "wrb_wrh_read", "wrb_file_encode", "wrb_file_decode", "wrb_file_last_error",
           "wrb_mssg_read_ctl", "wrb_mssg_read_nmlst", "wrb_mssg_header_begin", "wrb_mssg_header_time", "wrb_mssg_header_append",
           "wrb_mssg_header_read", "wrb_mssg_encode", "wrb_mssg_decode", "wrb_mssg_last_error",
           "encoding_wrap", "decoding_wrap", "setup_wr", "encoding_wrap_f", "decoding_wrap_f", "setup_wr_f"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WaveRangeError("libwaverange_b200.so is not built (run `python -m waverange_b200.build` or "
                             "__graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, d, ul = C.c_void_p, C.c_int, C.c_double, C.c_ulong
    H = C.POINTER(Header)
    L.wrb_create.argtypes = [C.POINTER(vp), i]
    L.wrb_destroy.argtypes = [vp]
    L.wrb_destroy.restype = None
    L.wrb_last_error.argtypes = [vp]
    L.wrb_last_error.restype = C.c_char_p
    L.wrb_set_stream.argtypes = [vp, vp]
    L.wrb_set_chunk_blocks.argtypes = [vp, i]
    L.wrb_set_seek_points.argtypes = [vp, i]
    L.wrb_set_local_cutoff.argtypes = [vp, i, i, i, C.POINTER(d)]
    L.wrb_launch_count.argtypes = [vp]
    L.wrb_launch_count.restype = C.c_ulonglong
    L.wrb_trim.argtypes = [vp]
    L.wrb_current_device.argtypes = [C.POINTER(i)]
    L.wrb_layer_guess_misses.argtypes = [vp]
    L.wrb_layer_guess_misses.restype = C.c_ulonglong
    L.wrb_setup.argtypes = [i, i, i, C.POINTER(C.c_ubyte), C.POINTER(ul)]
    L.wrb_setup.restype = None
    L.wrb_encode_device.argtypes = [vp, vp, i, i, i, i, i, d, H, vp, ul]
    L.wrb_decode_device.argtypes = [vp, vp, i, i, i, i, H, vp]
    L.wrb_encode_host.argtypes = [vp, vp, i, i, i, i, i, d, H, vp, ul]
    L.wrb_decode_host.argtypes = [vp, vp, i, i, i, i, H, vp]
    L.wrb_set_slab.argtypes = [vp, i, i, HALO_FN, REDUCE_FN, vp]
    L.wrb_comm_unique_id.argtypes = [C.c_char_p]
    L.wrb_set_comm.argtypes = [vp, i, i, C.c_char_p]
    L.wrb_set_slab_peers.argtypes = [vp, C.POINTER(vp), i]
    L.wrb_set_slab_order.argtypes = [vp, i]
    L.wrb_comm_counters.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.wrb_slab_order_plane.argtypes = [i] * 8
    L.wrb_slab_chunk_range.argtypes = [i, i, i, i, ul, i, C.POINTER(ul), C.POINTER(ul)]
    L.wrb_encode_slab_device.argtypes = [vp, vp, i, i, i, i, i, i, i, d, H, vp, ul]
    L.wrb_decode_slab_device.argtypes = [vp, vp, i, i, i, i, i, i, H, vp]
    L.wrb_decode_symbols_device.argtypes = [vp, vp, i, i, i, i, H, vp]
    L.wrb_decode_slab_symbols_device.argtypes = [vp, vp, i, i, i, i, i, i, H, vp]
    L.wrb_quantise_slab_device.argtypes = [vp, vp, i, i, i, i, i, i, i, d, H, vp, vp]
    L.wrb_wavelet3d_device.argtypes = [vp, vp, i, i, i, i]
    L.wrb_quantise_device.argtypes = [vp, vp, i, i, i, i, i, d, H, vp, vp]
    L.wrb_range_encode_device.argtypes = [vp, vp, ul, ul, vp, ul, C.POINTER(ul), C.POINTER(ul)]
    L.wrb_range_decode_device.argtypes = [vp, vp, C.POINTER(ul), ul, ul, vp]
    L.wrb_ind_p2w_3d.argtypes = [i] * 7 + [C.POINTER(i)] * 4
    L.wrb_ind_p2w_3d.restype = None
    L.wrb_set_timing.argtypes = [vp, i]
    L.wrb_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float)]
    cp = C.c_char_p
    L.wrb_wrh_begin.argtypes = [cp, cp, i, i, i]
    L.wrb_wrh_append.argtypes = [cp, i, C.POINTER(FieldRecord)]
    L.wrb_wrh_read.argtypes = [cp, C.POINTER(i), C.POINTER(FieldRecord), i]
    L.wrb_file_encode.argtypes = [vp, cp, cp, cp, i, i, i, C.POINTER(FieldDesc), C.POINTER(d)]
    L.wrb_file_decode.argtypes = [vp, cp, cp, cp, i, i]
    L.wrb_file_last_error.argtypes = []
    L.wrb_file_last_error.restype = cp
    L.wrb_mssg_read_ctl.argtypes = [cp, C.POINTER(MssgCtl)]
    L.wrb_mssg_read_nmlst.argtypes = [cp, C.POINTER(MssgNmlst)]
    L.wrb_mssg_header_begin.argtypes = [cp, cp, cp, i, i, i, d]
    L.wrb_mssg_header_time.argtypes = [cp, cp, C.POINTER(d)]
    L.wrb_mssg_header_append.argtypes = [cp, i, cp, H]
    L.wrb_mssg_header_read.argtypes = [cp, i, C.POINTER(d), C.POINTER(i), C.POINTER(i), vp, H, i]
    L.wrb_mssg_encode.argtypes = [vp, cp, cp, i, i, i, d, i]
    L.wrb_mssg_decode.argtypes = [vp, cp, cp, cp, i, i, i, i]
    L.wrb_mssg_last_error.argtypes = []
    L.wrb_mssg_last_error.restype = cp
    f64p, u8p, ulp = C.POINTER(d), C.POINTER(C.c_ubyte), C.POINTER(ul)
    L.encoding_wrap.argtypes = [i, i, i, f64p, i, i, i, i, f64p, f64p, f64p, f64p, u8p, u8p, ulp, f64p, f64p, ulp, u8p]
    L.encoding_wrap.restype = None
    L.decoding_wrap.argtypes = [i, i, i, f64p, f64p, f64p, f64p, u8p, u8p, ulp, f64p, f64p, ulp, u8p]
    L.decoding_wrap.restype = None
    L.setup_wr.argtypes = [i, i, i, u8p, ulp]
    L.setup_wr.restype = None
    _lib = L
    return L


def setup_wr(nx, ny, nz):
    """reference wrappers.cpp:531-541 -> (nlaymax, ntot_enc_max)"""
    n, m = C.c_ubyte(), C.c_ulong()
    lib().wrb_setup(nx, ny, nz, C.byref(n), C.byref(m))
    return n.value, m.value


def _np_ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def comm_unique_id():
    """128 bytes from ncclGetUniqueId (rank 0); hand them to every rank's Codec.set_comm"""
    buf = C.create_string_buffer(128)
    rc = lib().wrb_comm_unique_id(buf)
    if rc != 0:
        raise WaveRangeError("wrb_comm_unique_id failed (%d): NCCL not loadable?" % rc)
    return buf.raw


def slab_order_plane(nx, ny, nz, nranks, levels, rank, p, reg):
    return lib().wrb_slab_order_plane(nx, ny, nz, nranks, levels, rank, p, reg)


def slab_chunk_range(nx, ny, nz, nranks, chunk_len, rank):
    c0, c1 = C.c_ulong(), C.c_ulong()
    if lib().wrb_slab_chunk_range(nx, ny, nz, nranks, chunk_len, rank, C.byref(c0), C.byref(c1)) != 0:
        raise WaveRangeError("bad partition")
    return c0.value, c1.value


# ---- generic .wrh / .wrb files (include/waverange_files.h) -------------------------------------
def _fck(rc):
    if rc != 0:
        raise WaveRangeError("%s (code %d)" % (lib().wrb_file_last_error().decode(), rc))


def wrh_read(path):
    """all field records of a .wrh header file (reference gen_aux.cpp:554-644)"""
    n = C.c_int()
    _fck(lib().wrb_wrh_read(path.encode(), C.byref(n), None, 0))
    recs = (FieldRecord * max(1, n.value))()
    _fck(lib().wrb_wrh_read(path.encode(), C.byref(n), recs, n.value))
    return list(recs)[:n.value]


# ---- MSSG files (include/waverange_mssg.h) --------------------------------------------------------
def _mck(rc):
    if rc != 0:
        raise WaveRangeError("%s (code %d)" % (lib().wrb_mssg_last_error().decode(), rc))


def mssg_read_ctl(path):
    """GrADS control file of MSSG regular output (reference ctrl_aux.cpp:217-320) -> dict"""
    g = MssgCtl()
    _mck(lib().wrb_mssg_read_ctl(path.encode(), C.byref(g)))
    return dict(nx=g.nx, ny=g.ny, nz=g.nz, nt=g.nt, undef=g.undef, dset=g.dset.decode())


def mssg_read_nmlst(path):
    """MSSG restart namelist (reference ctrl_aux.cpp:49-213) -> dict"""
    m = MssgNmlst()
    _mck(lib().wrb_mssg_read_nmlst(path.encode(), C.byref(m)))
    return dict(nx=m.nx, ny=m.ny, nz=m.nz, nprocx=m.nprocx, nprocy=m.nprocy, ndset=m.ndset,
                dset=[m.dset[k].value.decode() for k in range(m.ndset)])


def mssg_header_read(path, filetype):
    """all of an MSSG encoding header -> (time record or None, [(id, name, Header)])  (reference ctrl_aux.cpp:538-585)"""
    n = C.c_int()
    t = (C.c_double * 15)()
    _mck(lib().wrb_mssg_header_read(path.encode(), filetype, t, C.byref(n), None, None, None, 0))
    k = max(1, n.value)
    ids, names, hdrs = (C.c_int * k)(), ((C.c_char * 256) * k)(), (Header * k)()
    _mck(lib().wrb_mssg_header_read(path.encode(), filetype, t, C.byref(n), ids, C.cast(names, C.c_void_p), hdrs, n.value))
    return (list(t) if filetype else None), [(ids[j], names[j].value.decode(), hdrs[j]) for j in range(n.value)]


def mssg_header_write(path, prefix, ext, filetype, nbytes, endianflip, tol_base, time_name, time_rec, records):
    """preamble, the time record (types 1/2) and one record per (0-based id, name, Header)
    (reference mssg_enc.cpp:273-284, 459-486, ctrl_aux.cpp:498-535)"""
    _mck(lib().wrb_mssg_header_begin(path.encode(), prefix.encode(), ext.encode(), filetype, nbytes, endianflip, tol_base))
    if filetype:
        _mck(lib().wrb_mssg_header_time(path.encode(), time_name.encode(), (C.c_double * 15)(*time_rec)))
    for idset, name, h in records:
        _mck(lib().wrb_mssg_header_append(path.encode(), idset, name.encode(), C.byref(h)))


def wrh_write(path, encoded_name, filetype, endianflip, records):
    """preamble + one record per field (reference gen_enc.cpp:508-519, gen_aux.cpp:505-551)"""
    _fck(lib().wrb_wrh_begin(path.encode(), encoded_name.encode(), filetype, endianflip, len(records)))
    for k, r in enumerate(records):
        _fck(lib().wrb_wrh_append(path.encode(), k, C.byref(r)))


def encoding_wrap(fld, tol, wtflag=1, cutoff=None):
    """The reference's encoding_wrap on a host float64 array shaped (nz, ny, nx).
    cutoff = (mx, my, mz, values): the local-precision grid (tol is then ignored).
    Returns (Header, data_enc[:ntot_enc])."""
    L = lib()
    a = np.ascontiguousarray(fld, dtype=np.float64)
    nz, ny, nx = a.shape
    _, cap = setup_wr(nx, ny, nz)
    data = np.zeros(cap, dtype=np.uint8)
    cut = np.array([tol], dtype=np.float64)
    cmx = cmy = cmz = 1
    if cutoff is not None:
        cmx, cmy, cmz, vals = cutoff
        cut = np.ascontiguousarray(vals, dtype=np.float64)
        assert cut.size == cmx * cmy * cmz
    h = Header()
    tolabs, mid, half = C.c_double(), C.c_double(), C.c_double()
    wlev, nlay, ntot_enc = C.c_ubyte(), C.c_ubyte(), C.c_ulong()
    deps, minv = np.zeros(NLAYMAX), np.zeros(NLAYMAX)
    lens = (C.c_ulong * NLAYMAX)()
    L.encoding_wrap(nx, ny, nz, _np_ptr(a, C.c_double), wtflag, cmx, cmy, cmz, _np_ptr(cut, C.c_double),
                    C.byref(tolabs), C.byref(mid), C.byref(half), C.byref(wlev), C.byref(nlay), C.byref(ntot_enc),
                    _np_ptr(deps, C.c_double), _np_ptr(minv, C.c_double), lens, _np_ptr(data, C.c_ubyte))
    h.tolabs, h.midval, h.halfspanval = tolabs.value, mid.value, half.value
    h.wlev, h.nlay, h.ntot_enc = wlev.value, nlay.value, ntot_enc.value
    for k in range(nlay.value):
        h.deps_vec[k], h.minval_vec[k], h.len_enc_vec[k] = deps[k], minv[k], lens[k]
    return h, data[:ntot_enc.value].copy()


def decoding_wrap(shape, h, data):
    """The reference's decoding_wrap; returns a float64 array of `shape` = (nz, ny, nx)."""
    L = lib()
    nz, ny, nx = shape
    out = np.empty(shape, dtype=np.float64)
    buf = np.zeros(len(data) + 64, dtype=np.uint8)
    buf[:len(data)] = data
    tolabs, mid, half = C.c_double(h.tolabs), C.c_double(h.midval), C.c_double(h.halfspanval)
    wlev, nlay, ntot_enc = C.c_ubyte(h.wlev), C.c_ubyte(h.nlay), C.c_ulong(h.ntot_enc)
    deps = np.array(list(h.deps_vec), dtype=np.float64)
    minv = np.array(list(h.minval_vec), dtype=np.float64)
    lens = (C.c_ulong * NLAYMAX)(*list(h.len_enc_vec))
    L.decoding_wrap(nx, ny, nz, _np_ptr(out, C.c_double), C.byref(tolabs), C.byref(mid), C.byref(half),
                    C.byref(wlev), C.byref(nlay), C.byref(ntot_enc), _np_ptr(deps, C.c_double),
                    _np_ptr(minv, C.c_double), lens, _np_ptr(buf, C.c_ubyte))
    return out


def ind_p2w_3d(lvl, n, idx):
    o = [C.c_int() for _ in range(4)]
    lib().wrb_ind_p2w_3d(lvl, n[0], n[1], n[2], idx[0], idx[1], idx[2], *[C.byref(x) for x in o])
    return tuple(x.value for x in o)


def parse_container(layer_bytes):
    """Split one layer of data_enc into its chunk streams.
    Returns (chunk_len, [stream bytes...]); a bare reference stream gives (0, [layer])."""
    b = bytes(layer_bytes)
    if b[:4] != b"WRCK":
        return 0, [b]
    assert int.from_bytes(b[4:8], "little") in (2, 3), "container version"
    chunk_len = int.from_bytes(b[8:16], "little")
    nch = int.from_bytes(b[24:28], "little")
    nseek = int.from_bytes(b[28:32], "little")
    lens = np.frombuffer(b, dtype="<u4", count=nch, offset=32)
    off = 32 + 4 * nch + 10 * nseek * nch          # version 3: 10-byte seek entries (version 2 is only read without them)
    out = []
    for n in lens:
        out.append(b[off:off + int(n)])
        off += int(n)
    assert off == len(b), "container length mismatch"
    return chunk_len, out


class Codec:
    """wrb_codec handle.  Pointers are plain integers (device addresses)."""

    def __init__(self, device=0, chunk_blocks=None, stream=None):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.wrb_create(C.byref(h), device)
        if rc != 0 or not h.value:
            raise WaveRangeError("wrb_create failed (%d): no usable CUDA device; there is no CPU fallback" % rc)
        self.h = h
        if chunk_blocks is not None:
            self._ck(self.L.wrb_set_chunk_blocks(self.h, chunk_blocks))
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.wrb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise WaveRangeError("%s (code %d)" % (self.L.wrb_last_error(self.h).decode(), rc))

    def file_encode(self, in_name, encoded_name, header_name, filetype, endianflip, fields, cutoff_all=None):
        """wrenc (reference gen_enc.cpp): fields = list of FieldDesc; cutoff_all: one tolerance for every field
        (the reference front-end's behaviour, see include/waverange_files.h)"""
        arr = (FieldDesc * max(1, len(fields)))(*fields)
        cut = C.byref(C.c_double(cutoff_all)) if cutoff_all is not None else None
        _fck(self.L.wrb_file_encode(self.h, in_name.encode(), encoded_name.encode(), header_name.encode(), filetype,
                                    endianflip, len(fields), arr, cut))

    def file_decode(self, encoded_name, header_name, out_name, filetype, endianflip):
        """wrdec (reference gen_dec.cpp)"""
        _fck(self.L.wrb_file_decode(self.h, encoded_name.encode(), header_name.encode(), out_name.encode(), filetype,
                                    endianflip))

    def mssg_encode(self, prefix, ext, filetype, nbytes, endianflip, tol_base, procid=0):
        """wrmssgenc (reference mssg_enc.cpp): file names derive from the prefix, relative to the working directory"""
        _mck(self.L.wrb_mssg_encode(self.h, prefix.encode(), ext.encode(), filetype, nbytes, endianflip, tol_base, procid))

    def mssg_decode(self, in_prefix, ext, out_prefix, filetype, nbytes, endianflip, procid=0):
        """wrmssgdec (reference mssg_dec.cpp)"""
        _mck(self.L.wrb_mssg_decode(self.h, in_prefix.encode(), ext.encode(), out_prefix.encode(), filetype, nbytes,
                                    endianflip, procid))

    def set_stream(self, stream_handle):
        self._ck(self.L.wrb_set_stream(self.h, C.c_void_p(stream_handle)))

    def set_chunk_blocks(self, k):
        self._ck(self.L.wrb_set_chunk_blocks(self.h, k))

    def set_local_cutoff(self, mx=0, my=0, mz=0, cutoffvec=None):
        """encoding_wrap()'s mx*my*mz > 1 branch (wrappers.cpp:343-379); no arguments: off"""
        if cutoffvec is None:
            self._ck(self.L.wrb_set_local_cutoff(self.h, 0, 0, 0, None))
        else:
            v = np.ascontiguousarray(cutoffvec, dtype=np.float64)
            assert v.size == mx * my * mz
            self._ck(self.L.wrb_set_local_cutoff(self.h, mx, my, mz, _np_ptr(v, C.c_double)))

    def set_seek_points(self, n):
        self._ck(self.L.wrb_set_seek_points(self.h, n))

    def set_timing(self, on):
        self._ck(self.L.wrb_set_timing(self.h, int(on)))

    def stage_ms(self):
        a = (C.c_float * 4)()
        self._ck(self.L.wrb_last_stage_ms(self.h, a))
        return list(a)

    def launch_count(self):
        return int(self.L.wrb_launch_count(self.h))

    def trim(self):
        self._ck(self.L.wrb_trim(self.h))

    def layer_guess_misses(self):
        return int(self.L.wrb_layer_guess_misses(self.h))

    # ---- device path -----------------------------------------------------------------------
    def encode_device(self, d_field, dtype, nx, ny, nz, tol, d_out, cap, wtflag=1):
        h = Header()
        self._ck(self.L.wrb_encode_device(self.h, d_field, dtype, nx, ny, nz, wtflag, tol, C.byref(h), d_out, cap))
        return h

    def decode_device(self, d_out, dtype, nx, ny, nz, h, d_data):
        self._ck(self.L.wrb_decode_device(self.h, d_out, dtype, nx, ny, nz, C.byref(h), d_data))

    # ---- z-slab partition -----------------------------------------------------------------
    def set_slab(self, rank, nranks, halo_cb, reduce_cb):
        """halo_cb / reduce_cb: HALO_FN / REDUCE_FN instances (kept alive by the caller)"""
        self._ck(self.L.wrb_set_slab(self.h, rank, nranks, halo_cb, reduce_cb, None))

    def set_comm(self, rank, nranks, unique_id):
        """NCCL transport inside the library + the global symbol order (include/waverange_b200.h wrb_set_comm)"""
        assert len(unique_id) == 128
        self._ck(self.L.wrb_set_comm(self.h, rank, nranks, unique_id))

    def set_slab_peers(self, codecs):
        """ranks emulated in one process: the Codec objects of all ranks in rank order"""
        arr = (C.c_void_p * max(1, len(codecs)))(*[c.h.value for c in codecs])
        self._ck(self.L.wrb_set_slab_peers(self.h, arr, len(codecs)))

    def set_slab_order(self, global_order):
        self._ck(self.L.wrb_set_slab_order(self.h, int(bool(global_order))))

    def comm_counters(self):
        a = (C.c_ulonglong * 3)()
        self._ck(self.L.wrb_comm_counters(self.h, a))
        return dict(halo_bytes=a[0], halo_calls=a[1], reduce_calls=a[2])

    def encode_slab_device(self, d_field, dtype, nx, ny, nz, z0, nzl, tol, d_out, cap, wtflag=1):
        h = Header()
        self._ck(self.L.wrb_encode_slab_device(self.h, d_field, dtype, nx, ny, nz, z0, nzl, wtflag, tol, C.byref(h), d_out, cap))
        return h

    def quantise_slab_device(self, d_field, dtype, nx, ny, nz, z0, nzl, tol, wtflag=1, d_coef=None, d_sym=None):
        h = Header()
        self._ck(self.L.wrb_quantise_slab_device(self.h, d_field, dtype, nx, ny, nz, z0, nzl, wtflag, tol, C.byref(h), d_coef, d_sym))
        return h

    def decode_slab_device(self, d_out, dtype, nx, ny, nz, z0, nzl, h, d_data):
        self._ck(self.L.wrb_decode_slab_device(self.h, d_out, dtype, nx, ny, nz, z0, nzl, C.byref(h), d_data))

    def decode_symbols_device(self, d_out, dtype, nx, ny, nz, h, d_sym):
        self._ck(self.L.wrb_decode_symbols_device(self.h, d_out, dtype, nx, ny, nz, C.byref(h), d_sym))

    def decode_slab_symbols_device(self, d_out, dtype, nx, ny, nz, z0, nzl, h, d_sym):
        self._ck(self.L.wrb_decode_slab_symbols_device(self.h, d_out, dtype, nx, ny, nz, z0, nzl, C.byref(h), d_sym))

    # ---- host path -------------------------------------------------------------------------
    def encode_host(self, fld, tol, wtflag=1, out=None):
        a = np.ascontiguousarray(fld)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        dtype = F32 if a.dtype == np.float32 else F64
        nz, ny, nx = a.shape
        _, cap = setup_wr(nx, ny, nz)
        data = out if out is not None else np.empty(cap, dtype=np.uint8)
        h = Header()
        self._ck(self.L.wrb_encode_host(self.h, a.ctypes.data, dtype, nx, ny, nz, wtflag, tol, C.byref(h),
                                        data.ctypes.data, min(cap, data.size)))
        return h, data[:h.ntot_enc]

    def decode_host(self, shape, h, data, dtype=np.float64, out=None):
        nz, ny, nx = shape
        res = out if out is not None else np.empty(shape, dtype=dtype)
        d = np.ascontiguousarray(data)
        self._ck(self.L.wrb_decode_host(self.h, res.ctypes.data, F32 if res.dtype == np.float32 else F64, nx, ny, nz,
                                        C.byref(h), d.ctypes.data if d.size else None))
        return res

    # ---- stages ----------------------------------------------------------------------------
    def wavelet3d_device(self, d_x, nx, ny, nz, lvl):
        self._ck(self.L.wrb_wavelet3d_device(self.h, d_x, nx, ny, nz, lvl))

    def quantise_device(self, d_field, dtype, nx, ny, nz, tol, wtflag=1, d_coef=None, d_sym=None):
        h = Header()
        self._ck(self.L.wrb_quantise_device(self.h, d_field, dtype, nx, ny, nz, wtflag, tol, C.byref(h), d_coef, d_sym))
        return h

    def range_encode_device(self, d_sym, n, chunk_len, d_out, cap):
        nch = 1 if chunk_len == 0 or chunk_len >= n else (n + chunk_len - 1) // chunk_len
        lens = (C.c_ulong * nch)()
        total = C.c_ulong()
        self._ck(self.L.wrb_range_encode_device(self.h, d_sym, n, chunk_len, d_out, cap, lens, C.byref(total)))
        return list(lens), total.value

    def range_decode_device(self, d_in, lens, n, chunk_len, d_sym):
        arr = (C.c_ulong * len(lens))(*lens)
        self._ck(self.L.wrb_range_decode_device(self.h, d_in, arr, n, chunk_len, d_sym))
