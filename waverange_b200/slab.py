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

    def halo_planes(self, t, nown, lo, hi):
        """t: [lo + nown + hi, plane_bytes] view of the halo'd buffer.  The rank above needs my last `lo`
        planes, the rank below my first `hi` planes; nothing happens at the domain ends."""
        dist = self.dist
        ops = []
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, t[nown:nown + lo], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, t[lo + nown:lo + nown + hi], self.rank + 1, self.group))
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, t[lo:lo + hi], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, t[0:lo], self.rank - 1, self.group))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        self.halo_bytes += sum(op.tensor.numel() for op in ops if op.op == dist.irecv)

    def reduce_min(self, t):
        """t: int64 tensor, reduced in place with MIN over the group (a single collective)"""
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)

    def _halo(self, user, buf, elem_bytes, plane_elems, nown, lo, hi):
        try:
            pb = elem_bytes * plane_elems
            t = _tensor_from_ptr(self.torch, buf, (lo + nown + hi) * pb, self.cuda).view(lo + nown + hi, pb)
            self.halo_planes(t, nown, lo, hi)
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
        self.hooks = [self._make(r) for r in range(world)]

    def _make(self, rank):
        torch, world = self.torch, self.world

        def halo(user, buf, elem_bytes, plane_elems, nown, lo, hi):
            pb = elem_bytes * plane_elems
            t = _tensor_from_ptr(torch, buf, (lo + nown + hi) * pb, True).view(lo + nown + hi, pb)
            torch.cuda.synchronize()
            self.slots[rank] = t
            self.barrier.wait()
            if rank + 1 < world:
                t[lo + nown:lo + nown + hi].copy_(self.slots[rank + 1][lo:lo + hi])
            if rank > 0:
                t[0:lo].copy_(self.slots[rank - 1][nown:nown + lo])
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
