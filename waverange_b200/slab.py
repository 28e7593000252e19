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

    def reduce_min(self, t, total=False):
        """t: int64 tensor, reduced in place with MIN (total: SUM) over the group (a single collective)"""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if total else self.dist.ReduceOp.MIN, group=self.group)

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
            self.reduce_min(_tensor_from_ptr(self.torch, pbuf, 8 * abs(count), self.cuda).view(self.torch.int64), total=count < 0)
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
            t = _tensor_from_ptr(torch, pbuf, 8 * abs(count), True).view(torch.int64)
            torch.cuda.synchronize()
            self.slots[rank] = t
            self.barrier.wait()
            mn = torch.stack(list(self.slots)).sum(dim=0) if count < 0 else torch.stack(list(self.slots)).min(dim=0).values
            torch.cuda.synchronize()
            self.barrier.wait()                      # everyone has read before anyone writes
            t.copy_(mn)
            torch.cuda.synchronize()
            self.barrier.wait()
            return 0

        return api.HALO_FN(halo), api.REDUCE_FN(reduce)

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
        real = [e for e in err if e is not None and "callback failed" not in str(e)]      # the cause, not the broken barrier
        for e in real + [e for e in err if e is not None]:
            raise e
        return out


# ---------------------------------------------------------------------------------------------------
# Global symbol order.
#
# With wrb_set_comm (NCCL) or wrb_set_slab_peers (ranks emulated in one process) the LIBRARY exchanges the 1-byte
# symbols between the quantiser and the coder (csrc/slab_order.cu: every rank gathers its run of the global
# wavelet-space sequence straight from the peers' symbol planes over NVLink) and rank r codes the chunks
# [r * nchunks / world, (r + 1) * nchunks / world) of that sequence: every chunk stream is byte for byte the single-GPU
# run's and the reference's range_encode of that sub-array.  Each rank's blob is a self-contained WRCK container of its
# run; the helpers below split the pieces into chunk streams and join them into one ordinary container.
# local_to_global_z() above is the independent restatement of the index map the tests check the library against.
# ---------------------------------------------------------------------------------------------------
CHUNK = 59999


def chunk_ranges(ntot, world, chunk=CHUNK):
    """rank r codes the chunks [cb[r], cb[r+1]) of the global symbol sequence"""
    nch = (ntot + chunk - 1) // chunk
    cb = [r * nch // world for r in range(world + 1)]
    return nch, cb


def region_of(nx, ny, x, y, levels=4):
    """level at which (x, y) leaves the low box (1..levels), levels + 1 inside the coarsest box"""
    for k in range(1, levels + 1):
        if x >= (nx + (1 << k) - 1) >> k or y >= (ny + (1 << k) - 1) >> k:
            return k
    return levels + 1


def piece_streams(h, blob):
    """the chunk streams of every layer of one rank's piece: [[bytes, ...] per layer]"""
    out, off = [], 0
    for l in range(h.nlay):
        _, streams = api.parse_container(blob[off:off + h.len_enc_vec[l]])
        out.append(streams)
        off += h.len_enc_vec[l]
    return out


def wrck_container(chunk_len, nsym, lens, streams):
    """a WRCK v3 layer container without seek points from chunk byte lengths and the concatenated streams"""
    hdr = b"WRCK" + (3).to_bytes(4, "little") + int(chunk_len).to_bytes(8, "little") + int(nsym).to_bytes(8, "little") \
        + len(lens).to_bytes(4, "little") + (0).to_bytes(4, "little")
    return hdr + np.asarray(lens, dtype="<u4").tobytes() + streams


def join_pieces(h, pieces, ntot, chunk=CHUNK):
    """One ordinary container of the whole field from the ranks' pieces (pieces[r] = piece_streams of rank r): returns
    (header with the joined lengths, bytes).  What a writer of a single .wrb file does with the ranks' outputs."""
    hj = api.Header.from_buffer_copy(bytes(h))
    blob = b""
    for l in range(h.nlay):
        streams = [s for p in pieces for s in p[l]]
        layer = wrck_container(chunk if ntot > chunk else ntot, ntot, [len(s) for s in streams], b"".join(streams))
        hj.len_enc_vec[l] = len(layer)
        blob += layer
    hj.ntot_enc = len(blob)
    return hj, blob


def set_comm_from_dist(codec, torch, dist, device):
    """NCCL transport inside the library for a torch.distributed job: rank 0 draws the NCCL id, torch.distributed only
    broadcasts its 128 bytes; every collective of the codec is then issued by the library itself."""
    rank, world = dist.get_rank(), dist.get_world_size()
    idt = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(api.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    codec.set_comm(rank, world, idt.cpu().numpy().tobytes())


def stream_crcs(h, blob_host):
    """crc32 of every chunk stream of a piece / container: [[crc, ...] per layer]"""
    import zlib
    return [[zlib.crc32(s) for s in layer] for layer in piece_streams(h, blob_host)]
