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
    # ---- symbol exchange of the global-order mode: all_gather of the rank-local planes, gather of the rank's run ----
    nz, ny, nx = 64, 16, 24
    nzl = nz // world
    G = np.random.default_rng(9).integers(0, 256, size=(nz, ny, nx), dtype=np.uint8)      # same on both ranks
    ys, xs = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    gz = slab.local_to_global_z(nx, ny, nz, rank * nzl, nzl)
    mine = torch.from_numpy(G[gz, ys[None], xs[None]].reshape(-1).copy())
    go = slab.GlobalOrder(torch, nx, ny, nz, rank, world, torch.device("cpu"), chunk=997)
    allsym = hk.all_gather_u8(mine)
    ok &= tuple(allsym.shape) == (world, mine.numel())
    run = go.gather_run(allsym).numpy()
    ok &= np.array_equal(run, G.reshape(-1)[go.j0[rank]:go.j0[rank + 1]])
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


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("shape", [(128, 32, 48), (64, 40, 16)])
def test_global_order_maps(world, shape):
    """GlobalOrder (compact region tables) against local_to_global_z: the run a rank gathers is its slice of the
    global wavelet-space sequence, and scattering the runs back gives every rank its local planes"""
    import torch
    from waverange_b200 import slab
    nz, ny, nx = shape
    if (nz // world) % 32 != 0:
        pytest.skip("partition rule")
    nzl = nz // world
    rng = np.random.default_rng(world + nx)
    G = rng.integers(0, 256, size=(nz, ny, nx), dtype=np.uint8)          # symbols in global wavelet-space order
    ys, xs = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    local = []
    for r in range(world):
        gz = slab.local_to_global_z(nx, ny, nz, r * nzl, nzl)
        local.append(G[gz, ys[None], xs[None]].reshape(-1))
    allsym = torch.from_numpy(np.stack(local))
    gos = [slab.GlobalOrder(torch, nx, ny, nz, r, world, torch.device("cpu"), chunk=997) for r in range(world)]
    runs = []
    for r, go in enumerate(gos):
        run = go.gather_run(allsym).numpy()
        assert np.array_equal(run, G.reshape(-1)[go.j0[r]:go.j0[r + 1]])
        runs.append(run)
    assert sum(len(x) for x in runs) == G.size and gos[0].j0[-1] == G.size
    pitch = max(len(x) for x in runs)
    allruns = torch.zeros((world, pitch), dtype=torch.uint8)
    for r, x in enumerate(runs):
        allruns[r, :len(x)] = torch.from_numpy(x)
    for r, go in enumerate(gos):
        assert np.array_equal(go.scatter_local(allruns, pitch).numpy(), local[r])


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
