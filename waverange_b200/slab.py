"""z-slab partition of one large field across the GPUs of a box: the collectives.

The CUDA side (csrc/wavelet_slab.cu, csrc/codec.cu) asks for two things through C callbacks
(include/waverange_b200.h: wrb_halo_fn, wrb_reduce_fn): neighbour planes for the z lifting of every
level, and global min/max of order-preserving u64 keys.  This module supplies them:

  * DistHooks   -- torch.distributed (backend nccl on GPUs: send/recv of whole planes over NVLink and
                   all_reduce MIN/MAX; backend gloo on CPU tensors in the CPU tests);
  * LocalGroup  -- emulation of N ranks inside one process (one thread and one codec per rank, halos
                   copied between the ranks' buffers after host synchronisation).  It lets a single GPU
                   run the complete slab pipeline in the parity tests; nothing waits on the device.

Reference context: the reference has no distributed mode; the partition follows
SURVEY.md section 8e.  The rank-local coefficient order and its map to the reference's global
wavelet-space order are given by local_to_global_z().
"""
import ctypes as C
import threading

import numpy as np

from . import api

INT64_MIN = -(2 ** 63)


def partition(nz, world):
    """planes [z0, z0+nzl) owned by every rank; the slab kernels need whole, 32-aligned slabs"""
    if nz % world != 0 or (nz // world) % 32 != 0 or nz % 16 != 0:
        raise ValueError("z-slab mode needs nz divisible by 32 * world (nz=%d, world=%d)" % (nz, world))
    nzl = nz // world
    return [(r * nzl, nzl) for r in range(world)]


def local_to_global_z(nx, ny, nz, z0, nzl, levels=4):
    """gz[zl, y, x]: global wavelet-space plane of every element of a rank-local coefficient array
    (x and y are unchanged).  At level k the local box has nzl >> (k-1) planes; its upper half holds the
    z-high band (global plane = global low extent + owned pair index), its lower half outside the
    (x, y) low box holds z-low details, and the low-low-low box recurses."""
    gz = np.zeros((nzl, ny, nx), dtype=np.int64)
    xs = np.arange(nx)[None, :]
    ys = np.arange(ny)[:, None]
    done = np.zeros((ny, nx), dtype=bool)          # (x, y) already outside the recursing low box
    n0, n1, n2l, n2g, zg = nx, ny, nzl, nz, z0
    assigned = np.zeros((nzl, ny, nx), dtype=bool)
    for k in range(1, levels + 1):
        m0, m1, nl, Mg, pg = (n0 + 1) // 2, (n1 + 1) // 2, n2l // 2, n2g // 2, zg // 2
        inbox = (xs < n0) & (ys < n1) & ~done         # still inside this level's box in x, y
        # z-high band of this level: local planes [nl, 2nl) for in-box (x, y)
        for p in range(nl, 2 * nl):
            sel = inbox & ~assigned[p]
            gz[p][sel] = Mg + pg + (p - nl)
            assigned[p][sel] = True
        # z-low, but x or y high: final details of this level
        outer = inbox & ((xs >= m0) | (ys >= m1))
        for p in range(0, nl):
            sel = outer & ~assigned[p]
            gz[p][sel] = pg + p
            assigned[p][sel] = True
        done = done | outer
        n0, n1, n2l, n2g, zg = m0, m1, nl, Mg, pg
    inbox = (xs < n0) & (ys < n1) & ~done
    for p in range(0, n2l):                          # coarsest approximation
        sel = inbox & ~assigned[p]
        gz[p][sel] = zg + p
        assigned[p][sel] = True
    assert assigned.all()
    return gz


def _tensor_from_ptr(torch, ptr, nbytes, cuda):
    """uint8 tensor aliasing raw memory"""
    if cuda:
        class _Arr:
            pass
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(a, device="cuda")
    buf = (C.c_uint8 * nbytes).from_address(int(ptr))
    return torch.from_numpy(np.ctypeslib.as_array(buf))


class DistHooks:
    """halo exchange + key reduction over a torch.distributed process group"""

    def __init__(self, torch, dist, cuda=True, group=None):
        self.torch, self.dist, self.cuda, self.group = torch, dist, cuda, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.halo_cb = api.HALO_FN(self._halo)
        self.reduce_cb = api.REDUCE_FN(self._reduce)
        self.error = None
        self.halo_bytes = 0

    def halo_exchange(self, send_down, send_up, recv_lo, recv_hi):
        """uint8 tensors (or None for an empty transfer): my first planes go down to rank-1 and arrive there as its
        upper halo, my last planes go up to rank+1 as its lower halo; nothing happens at the domain ends."""
        dist = self.dist
        ops = []
        if self.rank + 1 < self.world:
            if send_up is not None:
                ops.append(dist.P2POp(dist.isend, send_up, self.rank + 1, self.group))
            if recv_hi is not None:
                ops.append(dist.P2POp(dist.irecv, recv_hi, self.rank + 1, self.group))
        if self.rank > 0:
            if send_down is not None:
                ops.append(dist.P2POp(dist.isend, send_down, self.rank - 1, self.group))
            if recv_lo is not None:
                ops.append(dist.P2POp(dist.irecv, recv_lo, self.rank - 1, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        self.halo_bytes += sum(op.tensor.numel() for op in ops if op.op == dist.irecv)

    def reduce_min(self, t):
        """t: int64 tensor, reduced in place with MIN over the group (a single collective)"""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)

    def all_gather_u8(self, t):
        """(world, n) uint8: the tensor t of every rank"""
        t = t.contiguous().reshape(-1)
        if self.cuda:
            out = self.torch.empty((self.world, t.numel()), dtype=self.torch.uint8, device=t.device)
            self.dist.all_gather_into_tensor(out, t, group=self.group)
            return out
        parts = [self.torch.empty_like(t) for _ in range(self.world)]          # gloo: list form
        self.dist.all_gather(parts, t, group=self.group)
        return self.torch.stack(parts)

    def _halo(self, user, send_down, send_up, recv_lo, recv_hi, down_bytes, up_bytes):
        try:
            t = lambda ptr, n: _tensor_from_ptr(self.torch, ptr, n, self.cuda) if n else None
            self.halo_exchange(t(send_down, down_bytes), t(send_up, up_bytes), t(recv_lo, up_bytes), t(recv_hi, down_bytes))
            return 0
        except Exception as e:          # never let an exception cross the C boundary
            self.error = e
            return 1

    def _reduce(self, user, pbuf, count):
        try:
            self.reduce_min(_tensor_from_ptr(self.torch, pbuf, 8 * count, self.cuda).view(self.torch.int64))
            return 0
        except Exception as e:
            self.error = e
            return 1


class LocalGroup:
    """N ranks in one process on one device (tests): rank r runs in its own thread with its own codec;
    the callbacks meet at a barrier after synchronising the device and copy between the ranks' buffers."""

    def __init__(self, torch, world):
        self.torch, self.world = torch, world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.gslots = [None] * world
        self.hooks = [self._make(r) for r in range(world)]

    def _make(self, rank):
        torch, world = self.torch, self.world

        def halo(user, send_down, send_up, recv_lo, recv_hi, down_bytes, up_bytes):
            t = lambda ptr, n: _tensor_from_ptr(torch, ptr, n, True) if n else None
            torch.cuda.synchronize()
            self.slots[rank] = (t(send_down, down_bytes), t(send_up, up_bytes))
            self.barrier.wait()
            if rank + 1 < world and down_bytes:
                t(recv_hi, down_bytes).copy_(self.slots[rank + 1][0])
            if rank > 0 and up_bytes:
                t(recv_lo, up_bytes).copy_(self.slots[rank - 1][1])
            torch.cuda.synchronize()
            self.barrier.wait()
            return 0

        def reduce(user, pbuf, count):
            t = _tensor_from_ptr(torch, pbuf, 8 * count, True).view(torch.int64)
            torch.cuda.synchronize()
            self.slots[rank] = t
            self.barrier.wait()
            mn = torch.stack(list(self.slots)).min(dim=0).values
            torch.cuda.synchronize()
            self.barrier.wait()                      # everyone has read before anyone writes
            t.copy_(mn)
            torch.cuda.synchronize()
            self.barrier.wait()
            return 0

        return api.HALO_FN(halo), api.REDUCE_FN(reduce)

    def all_gather_u8(self, rank, t):
        """(world, n) uint8: the tensor t of every rank (threads meet at the barrier)"""
        torch = self.torch
        torch.cuda.synchronize()
        self.gslots[rank] = t
        self.barrier.wait()
        out = torch.stack([x.reshape(-1) for x in self.gslots])
        torch.cuda.synchronize()
        self.barrier.wait()
        return out

    def rank_hooks(self, rank):
        """object with the all_gather_u8(t) method of DistHooks, for rank `rank`"""
        grp = self

        class _H:
            def all_gather_u8(self, t):
                return grp.all_gather_u8(rank, t)
        return _H()

    def run(self, fn):
        """fn(rank, halo_cb, reduce_cb) in one thread per rank; returns the list of results"""
        out, err = [None] * self.world, [None] * self.world

        def work(r):
            try:
                self.torch.cuda.set_device(0)
                out[r] = fn(r, *self.hooks[r])
            except Exception as e:      # noqa: BLE001
                err[r] = e
                self.barrier.abort()

        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out


# ---------------------------------------------------------------------------------------------------
# Global symbol order.
#
# In plain z-slab mode every rank codes the symbols of ITS coefficients in rank-local array order, so the chunk
# streams differ from those of a single-GPU run (the reconstruction does not).  The functions below put an exchange
# between the quantiser and the coder: the 1-byte symbols are brought into the wavelet-space order of the GLOBAL
# array, rank r takes a contiguous run of whole chunks of that sequence and codes it with the stage entry point
# wrb_range_encode_device -- every chunk stream is then byte for byte the one a single GPU (and the reference's
# range_encode on that sub-array) produces, and the pieces of all ranks join into an ordinary WRCK container that
# wrb_decode_device reads.  Decoding mirrors it (wrb_range_decode_device per rank, exchange back,
# wrb_decode_slab_symbols_device).
#
# The exchange is an all_gather of the layer's symbol plane (NVLink: every rank receives G x its own share) followed
# by an indexed gather of the wanted run through the inverse index map; the maps are small tables: the global plane
# of a local element depends only on its local plane and on the level at which its (x, y) leaves the low box.
# ---------------------------------------------------------------------------------------------------
CHUNK = 59999


def region_tables(nz, world, levels=4):
    """T[r, p, reg]: global wavelet-space plane of local plane p of rank r for an (x, y) position of region reg
    (reg = k in 1..levels: (x, y) leaves the low box at level k; reg = levels + 1: coarsest approximation), derived from
    local_to_global_z() on a 2^(levels+1)-wide stand-in for the (x, y) plane; inv[w, reg] = (rank, local plane)."""
    nzl = nz // world
    n = 1 << (levels + 1)
    T = np.zeros((world, nzl, levels + 2), dtype=np.int64)
    for r in range(world):
        gz = local_to_global_z(n, n, nz, r * nzl, nzl, levels)
        for k in range(1, levels + 1):
            T[r, :, k] = gz[:, 0, (n >> (k - 1)) - 1]          # x in the high half of level k, y = 0
        T[r, :, levels + 1] = gz[:, 0, 0]
    inv = np.full((nz, levels + 2, 2), -1, dtype=np.int64)
    for r in range(world):
        for reg in range(1, levels + 2):
            inv[T[r, :, reg], reg, 0] = r
            inv[T[r, :, reg], reg, 1] = np.arange(nzl)
    assert (inv[:, 1:, 0] >= 0).all()
    return T, inv


def region_map(torch, nx, ny, levels, device):
    """reg[y, x] (int64): level at which (x, y) leaves the low box, levels + 1 inside the coarsest box"""
    xs = torch.arange(nx, device=device)[None, :]
    ys = torch.arange(ny, device=device)[:, None]
    reg = torch.full((ny, nx), levels + 1, dtype=torch.int64, device=device)
    for k in range(levels, 0, -1):
        m0, m1 = nx >> k, ny >> k
        reg = torch.where((xs >= m0) | (ys >= m1), torch.full_like(reg, k), reg)
    return reg


def chunk_ranges(ntot, world, chunk=CHUNK):
    """rank r codes the chunks [cb[r], cb[r+1]) of the global symbol sequence"""
    nch = (ntot + chunk - 1) // chunk
    cb = [r * nch // world for r in range(world + 1)]
    return nch, cb


class GlobalOrder:
    """index maps of one (nx, ny, nz, world) geometry on one rank"""

    def __init__(self, torch, nx, ny, nz, rank, world, device, levels=4, chunk=CHUNK):
        self.torch, self.nx, self.ny, self.nz, self.rank, self.world, self.chunk = torch, nx, ny, nz, rank, world, chunk
        self.nzl = nz // world
        self.ntl = nx * ny * self.nzl
        self.ntot = nx * ny * nz
        T, inv = region_tables(nz, world, levels)
        self.T = torch.from_numpy(T).to(device)              # (world, nzl, levels + 2)
        self.inv = torch.from_numpy(inv).to(device)          # (nz, levels + 2, 2)
        self.reg = region_map(torch, nx, ny, levels, device)  # (ny, nx)
        self.nch, self.cb = chunk_ranges(self.ntot, world, chunk)
        self.j0 = [min(self.ntot, c * chunk) for c in self.cb]            # symbol range of every rank
        self.device = device

    def my_run_len(self):
        return self.j0[self.rank + 1] - self.j0[self.rank]

    def gather_run(self, allsym):
        """allsym: (world, ntl) uint8, rank-local planes of every rank -> my run of the global sequence"""
        torch = self.torch
        out = torch.empty(self.my_run_len(), dtype=torch.uint8, device=self.device)
        plane = self.nx * self.ny
        flat = allsym.reshape(-1)
        j_lo, j_hi = self.j0[self.rank], self.j0[self.rank + 1]
        step = 16 * plane
        regf = self.reg.reshape(-1)
        for a in range(j_lo, j_hi, step):
            b = min(j_hi, a + step)
            j = torch.arange(a, b, device=self.device)
            w = j // plane
            xy = j - w * plane
            sp = self.inv[w, regf[xy]]                        # (n, 2): source rank, local plane
            out[a - j_lo:b - j_lo] = flat[sp[:, 0] * self.ntl + sp[:, 1] * plane + xy]
        return out

    def scatter_local(self, allruns, run_pitch):
        """allruns: (world, run_pitch) uint8, the decoded runs of every rank -> my rank-local symbol plane"""
        torch = self.torch
        out = torch.empty(self.ntl, dtype=torch.uint8, device=self.device)
        plane = self.nx * self.ny
        flat = allruns.reshape(-1)
        j0 = torch.tensor(self.j0, device=self.device)
        regf = self.reg.reshape(-1)
        Tr = self.T[self.rank]
        for p0 in range(0, self.nzl, 16):
            p1 = min(self.nzl, p0 + 16)
            lidx = torch.arange(p0 * plane, p1 * plane, device=self.device)
            p = lidx // plane
            xy = lidx - p * plane
            j = Tr[p, regf[xy]] * plane + xy                  # global index of every local element
            owner = torch.bucketize(j, j0[1:], right=True)     # rank whose run holds j
            out[p0 * plane:p1 * plane] = flat[owner * run_pitch + (j - j0[owner])]
        return out


def wrck_container(chunk_len, nsym, lens, streams):
    """a WRCK v3 layer container without seek points from chunk byte lengths and the concatenated streams"""
    hdr = b"WRCK" + (3).to_bytes(4, "little") + int(chunk_len).to_bytes(8, "little") + int(nsym).to_bytes(8, "little") \
        + len(lens).to_bytes(4, "little") + (0).to_bytes(4, "little")
    return hdr + np.asarray(lens, dtype="<u4").tobytes() + streams


def encode_global(torch, codec, hooks, go, d_field_slab, dtype, tol, wtflag=1):
    """z-slab encode with the coder working on the GLOBAL symbol order.  Returns (header, pieces): pieces[l] =
    (chunk byte lengths of my chunk range, their concatenated streams as a uint8 tensor) for every layer."""
    nx, ny, nz, nzl, rank = go.nx, go.ny, go.nz, go.nzl, go.rank
    sym = torch.empty(api.NLAYMAX * go.ntl, dtype=torch.uint8, device=go.device)
    h = codec.quantise_slab_device(d_field_slab, dtype, nx, ny, nz, rank * nzl, nzl, tol, wtflag, d_sym=sym.data_ptr())
    pieces = []
    n = go.my_run_len()
    out = torch.empty(2 * n + 4096 * (go.cb[rank + 1] - go.cb[rank] + 1), dtype=torch.uint8, device=go.device)
    for l in range(h.nlay):
        allsym = hooks.all_gather_u8(sym[l * go.ntl:(l + 1) * go.ntl])
        run = go.gather_run(allsym)
        del allsym
        if n > 0:
            lens, total = codec.range_encode_device(run.data_ptr(), n, go.chunk, out.data_ptr(), out.numel())
            pieces.append((lens, out[:total].clone()))
        else:
            pieces.append(([], out[:0].clone()))
    return h, pieces


def decode_global(torch, codec, hooks, go, h, pieces, d_out_slab, dtype):
    """inverse of encode_global: pieces as returned there (this rank's chunk range of every layer)"""
    nx, ny, nz, nzl, rank = go.nx, go.ny, go.nz, go.nzl, go.rank
    n = go.my_run_len()
    pitch = max(go.j0[r + 1] - go.j0[r] for r in range(go.world))
    sym = torch.empty(max(1, h.nlay) * go.ntl, dtype=torch.uint8, device=go.device)
    run = torch.zeros(pitch, dtype=torch.uint8, device=go.device)
    for l in range(h.nlay):
        lens, streams = pieces[l]
        if n > 0:
            buf = torch.zeros(streams.numel() + 64, dtype=torch.uint8, device=go.device)
            buf[:streams.numel()] = streams
            codec.range_decode_device(buf.data_ptr(), lens, n, go.chunk, run.data_ptr())
        allruns = hooks.all_gather_u8(run)
        sym[l * go.ntl:(l + 1) * go.ntl] = go.scatter_local(allruns, pitch)
        del allruns
    codec.decode_slab_symbols_device(d_out_slab, dtype, nx, ny, nz, rank * nzl, nzl, h, sym.data_ptr())
