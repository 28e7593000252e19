"""CPU tests of the z-slab (multi-GPU) host logic: partition arithmetic, the rank-local -> global
wavelet-space index map, and the two collectives (halo exchange, min/max key reduction) over a real
torch.distributed process group (gloo, world size 2, 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_rules():
    from waverange_b200 import slab
    assert slab.partition(2048, 8) == [(r * 256, 256) for r in range(8)]
    assert slab.partition(128, 2) == [(0, 64), (64, 64)]
    with pytest.raises(ValueError):
        slab.partition(100, 2)
    with pytest.raises(ValueError):
        slab.partition(96, 2)          # 48 planes per rank: not a multiple of 32


@pytest.mark.parametrize("world", [1, 2, 4])
def test_local_to_global_map_is_a_bijection(world, oracle):
    """Every global wavelet-space plane of every (x, y) column is owned by exactly one rank, and the map
    agrees with where the reference's own transform puts the data: a field that depends on z only through
    which rank owns the plane keeps, after mapping, the support structure of the global transform."""
    from waverange_b200 import slab
    nx, ny, nz = 20, 12, 128
    seen = np.zeros((nz, ny, nx), dtype=np.int32)
    for z0, nzl in slab.partition(nz, world):
        gz = slab.local_to_global_z(nx, ny, nz, z0, nzl)
        assert gz.shape == (nzl, ny, nx)
        yy, xx = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
        for p in range(nzl):
            np.add.at(seen, (gz[p], yy, xx), 1)
    assert (seen == 1).all()
    # the map follows ind_p2w: physical plane z of rank-local data ends in the same band as the reference says
    lvl, _, _, w = oracle.ind_p2w(4, (nx, ny, nz), (0, 0, 70))
    gz1 = slab.local_to_global_z(nx, ny, nz, 64, 64) if world == 2 else None
    if gz1 is not None:
        assert w in set(gz1[:, 0, 0].tolist())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from waverange_b200 import slab
    from waverange_b200.csrc_keys import dkey_np       # same key function as the device code
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    hk = slab.DistHooks(torch, dist, cuda=False)
    ok = True
    # ---- halo exchange through the C-callback signature, on raw host memory ----
    lo, hi, nown, plane = 4, 3, 6, 10
    buf = np.full((lo + nown + hi, plane), -1.0)
    buf[lo:lo + nown] = (100 * rank + np.arange(nown))[:, None] + np.arange(plane)[None, :] / 100.0
    pb = plane * 8
    base = buf.ctypes.data
    # my first `hi` planes go down, my last `lo` planes go up; the halos arrive before / after the own planes
    rc = hk.halo_cb(None, base + lo * pb, base + nown * pb, base, base + (lo + nown) * pb, hi * pb, lo * pb)
    ok &= rc == 0 and hk.error is None
    own = lambda r: (100 * r + np.arange(nown))[:, None] + np.arange(plane)[None, :] / 100.0
    if rank > 0:
        ok &= np.array_equal(buf[:lo], own(rank - 1)[nown - lo:])
    else:
        ok &= (buf[:lo] == -1).all()                       # domain end: untouched
    if rank + 1 < world:
        ok &= np.array_equal(buf[lo + nown:], own(rank + 1)[:hi])
    else:
        ok &= (buf[lo + nown:] == -1).all()
    # ---- min/max reduction of order-preserving keys (negative values and keys >= 2^63 included) ----
    # the codec packs (min key, ~max key) with the top bit flipped into one int64 buffer (codec.cu pack_keys_kernel)
    allv = np.array([[-3.5, 2.0, 1e-300], [-1e300, 7.25, -0.125]])
    vals = allv[rank % 2]
    flip = np.uint64(1 << 63)
    buf = np.concatenate([dkey_np(vals) ^ flip, (~dkey_np(vals)) ^ flip]).view(np.int64).copy()
    rc = hk.reduce_cb(None, buf.ctypes.data, buf.size)
    ok &= rc == 0 and hk.error is None
    got = buf.view(np.uint64)
    kmin, kmax = got[:3] ^ flip, ~(got[3:] ^ flip)
    ok &= np.array_equal(kmin, dkey_np(allv.min(axis=0))) and np.array_equal(kmax, dkey_np(allv.max(axis=0)))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_halo_exchange_and_key_reduction_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(128, 32, 48), (256, 40, 16), (512, 18, 22)])
def test_library_index_map_matches_the_restatement(world, shape, product_lib):
    """The closed-form index map of the library (csrc/slab_order.cu, through wrb_slab_order_plane) against the
    independent restatement local_to_global_z(), which follows the in-box de-interleave of the reference
    (waveletcdf97_3d.c:128-135,256-263) level by level; and the chunk ranges of the ranks."""
    from waverange_b200 import api, slab
    nz, ny, nx = shape
    if (nz // world) % 32 != 0 or nz % world:
        pytest.skip("partition rule")
    nzl = nz // world
    reg = np.array([[slab.region_of(nx, ny, x, y) for x in range(nx)] for y in range(ny)])
    for r in range(world):
        gz = slab.local_to_global_z(nx, ny, nz, r * nzl, nzl)
        for k in range(1, 6):
            ys, xs = np.nonzero(reg == k)
            if ys.size == 0:
                continue
            for p in range(nzl):
                w = api.slab_order_plane(nx, ny, nz, world, 4, r, p, k)
                assert (gz[p, ys, xs] == w).all(), (r, p, k)
    nch, cb = slab.chunk_ranges(nx * ny * nz, world, 997)
    for r in range(world):
        assert api.slab_chunk_range(nx, ny, nz, world, 997, r) == (cb[r], cb[r + 1])


def test_join_pieces_gives_an_ordinary_container():
    """slab.join_pieces: the ranks' chunk streams, in rank order, as one container the parser (and the decoder) reads"""
    from waverange_b200 import api, slab
    rng = np.random.default_rng(1)
    streams = [[bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in ns] for ns in ((9, 600), (17,), (5, 5, 80))]
    h = api.Header()
    h.nlay = 1
    hj, blob = slab.join_pieces(h, [[s] for s in streams], 6 * 59999 - 3)
    cl, got = api.parse_container(np.frombuffer(blob, dtype=np.uint8))
    assert cl == 59999 and got == [x for s in streams for x in s] and hj.ntot_enc == len(blob) == hj.len_enc_vec[0]


def test_wrck_container_helper_matches_the_parser():
    """slab.wrck_container (joins the chunk streams the ranks coded in global symbol order) writes what
    api.parse_container -- and the decoder -- read: version 3, no seek points, u32 lengths, streams back to back"""
    import numpy as np
    from waverange_b200 import api, slab
    rng = np.random.default_rng(0)
    streams = [bytes(rng.integers(0, 256, n, dtype=np.uint8)) for n in (7, 513, 1, 9000)]
    blob = slab.wrck_container(59999, 3 * 59999 + 17, [len(s) for s in streams], b"".join(streams))
    assert blob[:4] == b"WRCK" and int.from_bytes(blob[4:8], "little") == 3
    assert int.from_bytes(blob[24:28], "little") == 4 and int.from_bytes(blob[28:32], "little") == 0
    chunk_len, got = api.parse_container(np.frombuffer(blob, dtype=np.uint8))
    assert chunk_len == 59999 and got == streams
