#!/usr/bin/env python3
"""BASELINE.json configs[3]: one large float32 field, z-slab partitioned over the GPUs of a box, coded in the GLOBAL
symbol order with every collective issued by the library (NCCL from C++, wrb_set_comm).

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/run_c4.py [--edge 2048] [--tol 1e-4] [--reps 3] [--single-gpu-check]

Every rank owns n/N planes.  Compress: NCCL halo exchange per wavelet level, all-reduce of the extrema per layer,
symbol exchange over NVLink into the global wavelet-space order, chunk coder on the rank's run of whole chunks;
decompress mirrors it.  Checks without an oracle at this size: the round trip meets the tolerance; header doubles, the
checksum of the reconstruction and (--single-gpu-check, when the whole field fits one GPU) the crc32 of EVERY chunk
stream are those of a single-GPU encode of the same field on rank 0.  Rank 0 prints one JSON line.
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


KMAX = 24


def slab_field(n, dev, z0, nzl, nz_total):
    field = torch.empty((nzl, n, n), dtype=torch.float32, device=dev)
    step = 64
    for zs in range(0, nzl, step):
        m = min(step, nzl - zs)
        field[zs:zs + m] = bench.synth_field(torch, n, 1234, dev, torch.float32, nz_total=nz_total, z0=z0 + zs, nzl=m, kmax=KMAX)
    return field


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", dest="n", type=int, default=2048)
    ap.add_argument("--nz", type=int, default=0, help="z extent (default: the edge)")
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--single-gpu-check", action="store_true")
    ap.add_argument("--local-order", action="store_true", help="rank-local symbol order (no symbol exchange)")
    ap.add_argument("--cap", type=float, default=3.0, help="capacity of the rank's output buffer in bytes per grid point (the library "
                    "fails cleanly when a piece does not fit); 1.0 for the runs that have to fit 2048^3 on two GPUs")
    ap.add_argument("--kmax", type=int, default=0, help="largest wave number per box edge of the synthetic field (default 24 * edge / 512: "
                    "the 512^3 benchmark field's spectrum per grid point, i.e. the same compressibility; 24 gives a field that is "
                    "4x smoother per grid point at 2048^3 and codes 49:1)")
    a = ap.parse_args()
    global KMAX
    KMAX = a.kmax or max(24, (24 * a.n) // 512)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # descriptor 1 -> stderr while NCCL may print its banner; the JSON line goes to the saved descriptor
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    n, nz = a.n, (a.nz or a.n)
    nzl = nz // world
    z0 = rank * nzl
    field = slab_field(n, dev, z0, nzl, nz)
    torch.cuda.empty_cache()                     # the generator's temporaries go back to the driver: the library needs the room
    stream = torch.cuda.current_stream()
    codec = api.Codec(device=local, stream=stream.cuda_stream)
    codec.set_timing(True)
    slab.set_comm_from_dist(codec, torch, dist, dev)
    if a.local_order:
        codec.set_slab_order(False)
    ntl = n * n * nzl
    cap = int(ntl * a.cap) + (1 << 20)
    blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    rec = torch.empty(ntl, dtype=torch.float32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    enc_ms, dec_ms, se, sd = [], [], [], []
    h = None
    for it in range(a.reps + 1):
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record(stream)
        h = codec.encode_slab_device(field.data_ptr(), api.F32, n, n, nz, z0, nzl, a.tol, blob.data_ptr(), cap)
        e = codec.stage_ms()
        ev[1].record(stream)
        codec.decode_slab_device(rec.data_ptr(), api.F32, n, n, nz, z0, nzl, h, blob.data_ptr())
        d = codec.stage_ms()
        ev[2].record(stream)
        torch.cuda.synchronize()
        if it > 0:
            enc_ms.append(ev[0].elapsed_time(ev[1])); dec_ms.append(ev[1].elapsed_time(ev[2])); se.append(e); sd.append(d)
    # checks in z-pieces (whole-slab float64 temporaries would be four times the slab)
    recv = rec.view_as(field)
    err = torch.zeros((), dtype=torch.float64, device=dev)
    amax = torch.zeros((), dtype=torch.float64, device=dev)
    chk = torch.zeros((), dtype=torch.int64, device=dev)   # order-independent checksum: sum of the float bit patterns
    for zs in range(0, nzl, 32):
        fp, rp = field[zs:zs + 32].double(), recv[zs:zs + 32]
        err = torch.maximum(err, (rp.double() - fp).abs().max())
        amax = torch.maximum(amax, fp.abs().max())
        chk += rp.contiguous().view(torch.int32).to(torch.int64).sum()
    del fp, rp
    size = torch.tensor([float(h.ntot_enc)], device=dev, dtype=torch.float64)
    t = torch.tensor([sum(enc_ms) / len(enc_ms), sum(dec_ms) / len(dec_ms)] + [sum(x[i] for x in se) / len(se) for i in range(4)]
                     + [sum(x[i] for x in sd) / len(sd) for i in range(4)], device=dev, dtype=torch.float64)
    dist.all_reduce(err, op=dist.ReduceOp.MAX); dist.all_reduce(amax, op=dist.ReduceOp.MAX)
    dist.all_reduce(chk, op=dist.ReduceOp.SUM); dist.all_reduce(size, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    in_use = float(torch.cuda.mem_get_info(dev)[1] - torch.cuda.mem_get_info(dev)[0])
    mem = torch.tensor([float(torch.cuda.max_memory_allocated(dev)), in_use, in_use - float(torch.cuda.memory_reserved(dev))],
                       device=dev, dtype=torch.float64)
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    ok = bool(err.item() <= 1.10 * a.tol * amax.item())
    check = None
    if a.single_gpu_check:
        crcs = slab.stream_crcs(h, blob[:h.ntot_enc].cpu().numpy())
        allc = [None] * world
        dist.gather_object(crcs, allc if rank == 0 else None, dst=0)
        del rec, blob, field
        codec.trim()
        torch.cuda.empty_cache()
        if rank == 0:
            whole = slab_field(n, dev, 0, nz, nz)
            c1 = api.Codec(device=local, stream=stream.cuda_stream)
            cap1 = whole.numel() * 3 + (1 << 20)
            blob1 = torch.empty(cap1 + 64, dtype=torch.uint8, device=dev)
            h1 = c1.encode_device(whole.data_ptr(), api.F32, n, n, nz, a.tol, blob1.data_ptr(), cap1)
            want = slab.stream_crcs(h1, blob1[:h1.ntot_enc].cpu().numpy())
            rec1 = torch.empty(whole.numel(), dtype=torch.float32, device=dev)
            c1.decode_device(rec1.data_ptr(), api.F32, n, n, nz, h1, blob1.data_ptr())
            got = [[x for r in range(world) for x in allc[r][l]] for l in range(h.nlay)]
            check = {"chunk_streams_equal_single_gpu": bool(h1.nlay == h.nlay and got == want),
                     "chunk_streams": sum(len(x) for x in want),
                     "header_doubles_equal": bool(list(h1.deps_vec) == list(h.deps_vec) and list(h1.minval_vec) == list(h.minval_vec)
                                                  and h1.tolabs == h.tolabs),
                     "reconstruction_checksum_equal": bool(int(rec1.view(torch.int32).to(torch.int64).sum().item()) == int(chk.item())),
                     "single_gpu_bytes": int(h1.ntot_enc)}
            ok = ok and all(v for k, v in check.items() if k.endswith("equal") or k.endswith("gpu"))
        dist.barrier()
    if rank == 0:
        nbytes = 4 * n * n * nz
        cnt = codec.comm_counters()
        line = {"config": "%dx%dx%d float32 (kmax %d), tol %g, z-slabs over %d GPUs, %s symbol order" % (n, n, nz, KMAX, a.tol, world, "rank-local" if a.local_order else "global"),
                "n_gpus": world, "compress_gbs": nbytes / (t[0].item() * 1e-3) / 1e9, "decompress_gbs": nbytes / (t[1].item() * 1e-3) / 1e9,
                "encode_ms": t[0].item(), "decode_ms": t[1].item(),
                "stages_ms": {"encode": dict(zip(["transform", "quantise+exchange", "range_encode", "assemble"], t[2:6].tolist())),
                              "decode": dict(zip(["parse", "range_decode+exchange", "dequantise", "inverse_transform"], t[6:10].tolist()))},
                "rel_linf_error": err.item() / amax.item(), "within_tolerance": ok, "nlay": int(h.nlay),
                "ntot_enc_all_ranks": int(size.item()), "ratio": nbytes / size.item(), "tolabs": h.tolabs, "midval": h.midval,
                "deps_vec": list(h.deps_vec)[:h.nlay], "minval_vec": list(h.minval_vec)[:h.nlay],
                "reconstruction_checksum": int(chk.item()), "nccl_rank0": cnt,
                "torch_peak_bytes_max": mem[0].item(), "device_bytes_in_use_max": mem[1].item(),
                "device_bytes_outside_torch_max": mem[2].item(), "check": check}
        os.write(real, (json.dumps(line) + "\n").encode())
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
