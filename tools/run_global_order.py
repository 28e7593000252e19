#!/usr/bin/env python3
"""z-slab partition with the coder on the GLOBAL symbol order (waverange_b200/slab.py encode_global / decode_global)
over real NCCL: checks, against a single-GPU encode of the whole field on rank 0, that the chunk streams the ranks
produce are byte for byte the single-GPU run's, and times the mode.

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29521 tools/run_global_order.py [--edge 512] [--tol 1e-4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from waverange_b200 import api, slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", dest="n", type=int, default=512)
    ap.add_argument("--tol", type=float, default=1e-4)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = a.n
    nzl = n // world
    z0 = rank * nzl
    field = bench.synth_field(torch, n, 1234, dev, torch.float32, nz_total=n, z0=z0, nzl=nzl)
    stream = torch.cuda.current_stream()
    codec = api.Codec(device=local, stream=stream.cuda_stream)
    hooks = slab.DistHooks(torch, dist, cuda=True)
    codec.set_slab(rank, world, hooks.halo_cb, hooks.reduce_cb)
    go = slab.GlobalOrder(torch, n, n, n, rank, world, dev)
    rec = torch.empty(n * n * nzl, dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    times = []
    for it in range(3):
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record(stream)
        h, pieces = slab.encode_global(torch, codec, hooks, go, field.data_ptr(), api.F32, a.tol)
        ev[1].record(stream)
        slab.decode_global(torch, codec, hooks, go, h, pieces, rec.data_ptr(), api.F32)
        ev[2].record(stream)
        torch.cuda.synchronize()
        times.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])))
    err = (rec.view_as(field).double() - field.double()).abs().max()
    amax = field.double().abs().max()
    dist.all_reduce(err, op=dist.ReduceOp.MAX); dist.all_reduce(amax, op=dist.ReduceOp.MAX)
    # all pieces and the whole field to rank 0, which encodes it alone
    mine = [(list(l), s.cpu().numpy().tobytes()) for l, s in pieces]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    whole = torch.empty((world, nzl, n, n), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(whole, field.contiguous())
    same = None
    if rank == 0:
        solo = api.Codec(device=local)
        _, cap = api.setup_wr(n, n, n)
        cap = min(cap, 5 * n ** 3 + (1 << 20))
        blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
        h1 = solo.encode_device(whole.data_ptr(), api.F32, n, n, n, a.tol, blob.data_ptr(), cap)
        data = blob[:h1.ntot_enc].cpu().numpy()
        same = h1.nlay == h.nlay and list(h1.deps_vec)[:h.nlay] == list(h.deps_vec)[:h.nlay]
        off = 0
        for l in range(h1.nlay):
            _, streams = api.parse_container(data[off:off + h1.len_enc_vec[l]])
            off += h1.len_enc_vec[l]
            joined = b"".join(g[l][1] for g in gathered)
            lens = [x for g in gathered for x in g[l][0]]
            same = same and lens == [len(s) for s in streams] and joined == b"".join(streams)
        t = np.array(times[1:]).mean(axis=0)
        print(json.dumps({"config": "%d^3 float32, tol %g, z-slabs over %d GPUs, global symbol order" % (n, a.tol, world),
                          "chunk_streams_equal_single_gpu": bool(same), "rel_linf_error": err.item() / amax.item(),
                          "within_tolerance": bool(err.item() <= 1.1 * a.tol * amax.item()), "nlay": int(h.nlay),
                          "encode_global_ms": float(t[0]), "decode_global_ms": float(t[1])}), flush=True)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
