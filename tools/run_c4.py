#!/usr/bin/env python3
"""BASELINE.json configs[3]: one 2048^3 float32 field, z-slab partitioned over the GPUs of a box.

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_c4.py [--edge 2048] [--tol 1e-4] [--reps 3]

Every rank owns n/N planes, compresses and decompresses its slab (NCCL halo exchange per wavelet level, one
all_reduce of the extrema per layer) and the run checks what can be checked without the oracle at this size:
the round trip meets the tolerance, and -- the transform and the layer parameters being those of the GLOBAL field --
the header doubles and a checksum of the reconstruction are the same for every N (compare the printed JSON lines of
runs with different N).  Rank 0 prints one JSON line with the device-timed throughput.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from waverange_b200 import api, slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", dest="n", type=int, default=2048)
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = a.n
    nzl = n // world
    z0 = rank * nzl
    # the slab of the global field (same generator as bench.py, evaluated for planes [z0, z0 + nzl))
    field = torch.empty((nzl, n, n), dtype=torch.float32, device=dev)
    step = 64
    for zs in range(0, nzl, step):
        m = min(step, nzl - zs)
        field[zs:zs + m] = bench.synth_field(torch, n, 1234, dev, torch.float32, nz_total=n, z0=z0 + zs, nzl=m)
    stream = torch.cuda.current_stream()
    codec = api.Codec(device=local, stream=stream.cuda_stream)
    codec.set_timing(True)
    hooks = slab.DistHooks(torch, dist, cuda=True)
    codec.set_slab(rank, world, hooks.halo_cb, hooks.reduce_cb)
    ntl = n * n * nzl
    cap = ntl * 5 + (1 << 20)
    blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    rec = torch.empty(ntl, dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    enc_ms, dec_ms = [], []
    h = None
    for it in range(a.reps + 1):
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record(stream)
        h = codec.encode_slab_device(field.data_ptr(), api.F32, n, n, n, z0, nzl, a.tol, blob.data_ptr(), cap)
        ev[1].record(stream)
        codec.decode_slab_device(rec.data_ptr(), api.F32, n, n, n, z0, nzl, h, blob.data_ptr())
        ev[2].record(stream)
        torch.cuda.synchronize()
        if it > 0:
            enc_ms.append(ev[0].elapsed_time(ev[1])); dec_ms.append(ev[1].elapsed_time(ev[2]))
    if hooks.error is not None:
        raise hooks.error
    err = (rec.view_as(field).double() - field.double()).abs().max()
    amax = field.double().abs().max()
    # order-independent checksum of the reconstruction: sum of the float bit patterns as int64
    chk = rec.view(torch.int32).to(torch.int64).sum()
    size = torch.tensor([float(h.ntot_enc)], device=dev, dtype=torch.float64)
    t = torch.tensor([sum(enc_ms) / len(enc_ms), sum(dec_ms) / len(dec_ms)], device=dev, dtype=torch.float64)
    dist.all_reduce(err, op=dist.ReduceOp.MAX); dist.all_reduce(amax, op=dist.ReduceOp.MAX)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM); dist.all_reduce(size, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = bool(err.item() <= 1.10 * a.tol * amax.item())
    if rank == 0:
        nbytes = 4 * n ** 3
        print(json.dumps({"config": "%d^3 float32, tol %g, z-slabs over %d GPUs" % (n, a.tol, world), "n_gpus": world,
                          "compress_gbs": nbytes / (t[0].item() * 1e-3) / 1e9, "decompress_gbs": nbytes / (t[1].item() * 1e-3) / 1e9,
                          "encode_ms": t[0].item(), "decode_ms": t[1].item(), "rel_linf_error": err.item() / amax.item(),
                          "within_tolerance": ok, "nlay": int(h.nlay), "ntot_enc_all_ranks": int(size.item()),
                          "ratio": nbytes / size.item(), "tolabs": h.tolabs, "midval": h.midval,
                          "deps_vec": list(h.deps_vec)[:h.nlay], "minval_vec": list(h.minval_vec)[:h.nlay],
                          "reconstruction_checksum": int(chk.item()), "halo_bytes_rank0": int(hooks.halo_bytes)}), flush=True)
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
