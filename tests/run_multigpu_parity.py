#!/usr/bin/env python3
"""Multi-GPU parity against the ORACLE on real ranks (run under torchrun; not collected by pytest):

    gpurun --gpus N -- python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29531 tests/run_multigpu_parity.py [--edge 512] [--tol 1e-4] [--dtype f32]

An edge^3 field in z-slabs over the N ranks, NCCL issued by the library (wrb_set_comm), global symbol order.  Rank 0
runs the oracle's multi-threaded digest encode of the WHOLE field on the host (oracle/wr_oracle.c wro_encode_digest)
and checks: header doubles bit-equal on every rank, every chunk length and the FNV-1a hash of EVERY chunk stream of all
ranks (in rank order) equal to the oracle's, round trip within tolerance.  One JSON line; exit code 0 iff all equal."""
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

L1 = 59999


def piece_hashes(oracle, h, blob, runlen):
    """chunk byte lengths and FNV-1a of every chunk stream of one rank's piece, per layer"""
    nch = (runlen + L1 - 1) // L1
    lens_all, hash_all, off = [], [], 0
    data = blob[:h.ntot_enc].cpu().numpy()
    for l in range(h.nlay):
        assert bytes(data[off:off + 4]) == b"WRCK" and int.from_bytes(bytes(data[off + 24:off + 28]), "little") == nch
        nseek = int.from_bytes(bytes(data[off + 28:off + 32]), "little")
        lens = np.frombuffer(bytes(data[off + 32:off + 32 + 4 * nch]), dtype="<u4").astype(np.uint64)
        starts = np.uint64(off + 32 + 4 * nch + 10 * nseek * nch) + np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
        hash_all.append(oracle.fnv1a_many(data, starts, lens))
        lens_all.append(lens)
        off += h.len_enc_vec[l]
    return lens_all, hash_all


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", dest="n", type=int, default=512)
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--dtype", default="f32")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    from oracle.binding import Restatement
    oracle = Restatement()
    n = a.n
    tdt = torch.float32 if a.dtype == "f32" else torch.float64
    code = api.F32 if a.dtype == "f32" else api.F64
    nzl = n // world
    z0 = rank * nzl
    field = bench.synth_field(torch, n, 1234, dev, tdt, nz_total=n, z0=z0, nzl=nzl)
    codec = api.Codec(device=local, stream=torch.cuda.current_stream().cuda_stream)
    slab.set_comm_from_dist(codec, torch, dist, dev)
    ntl = n * n * nzl
    cap = ntl * field.element_size() + (8 << 20)
    blob = torch.zeros(cap + 64, dtype=torch.uint8, device=dev)
    h = codec.encode_slab_device(field.data_ptr(), code, n, n, n, z0, nzl, a.tol, blob.data_ptr(), cap)
    rec = torch.empty(ntl, dtype=tdt, device=dev)
    codec.decode_slab_device(rec.data_ptr(), code, n, n, n, z0, nzl, h, blob.data_ptr())
    err = (rec.view_as(field).double() - field.double()).abs().max()
    amax = field.double().abs().max()
    dist.all_reduce(err, op=dist.ReduceOp.MAX); dist.all_reduce(amax, op=dist.ReduceOp.MAX)
    nch, cb = slab.chunk_ranges(n ** 3, world)
    runlen = min(n ** 3, cb[rank + 1] * L1) - min(n ** 3, cb[rank] * L1)
    lens, hashes = piece_hashes(oracle, h, blob, runlen)
    mine = dict(hdr=[h.tolabs, h.midval, h.halfspanval] + list(h.deps_vec) + list(h.minval_vec), nlay=int(h.nlay),
                lens=[x.tolist() for x in lens], hashes=[x.tolist() for x in hashes])
    allr = [None] * world
    dist.gather_object(mine, allr if rank == 0 else None, dst=0)
    ok = True
    line = None
    if rank == 0:
        whole = bench.synth_field(torch, n, 1234, dev, tdt, nz_total=n, z0=0, nzl=n).cpu().numpy().astype(np.float64)
        want = oracle.encode_digest(whole, a.tol, L1, inplace=True)
        hw = want["header"]
        ref = [hw.tolabs, hw.midval, hw.halfspan] + list(hw.deps) + list(hw.minval)
        hdr_ok = all(r["nlay"] == hw.nlay and np.array(r["hdr"]).tobytes() == np.array(ref).tobytes() for r in allr)
        bad = 0
        for l in range(hw.nlay):
            glens = np.array([x for r in allr for x in r["lens"][l]], dtype=np.uint64)
            ghash = np.array([x for r in allr for x in r["hashes"][l]], dtype=np.uint64)
            if glens.size != nch or not np.array_equal(glens, want["chunk_lens"][l].astype(np.uint64)):
                bad += nch
            else:
                bad += int((ghash != want["stream_hash"][l]).sum())
        ok = hdr_ok and bad == 0 and err.item() <= 1.10 * a.tol * amax.item()
        line = {"config": "%d^3 %s tol %g, %d ranks, NCCL in the library, global symbol order" % (n, a.dtype, a.tol, world),
                "nlay": int(hw.nlay), "chunk_streams": int(nch * hw.nlay), "chunk_streams_differing_from_oracle": bad,
                "header_doubles_equal_oracle": bool(hdr_ok), "rel_linf_error": err.item() / amax.item(), "ok": bool(ok),
                "ntot_enc_oracle_chunked": int(hw.ntot_enc), "ntot_enc_all_ranks": int(sum(sum(sum(x) for x in r["lens"]) for r in allr))}
        os.write(real, (json.dumps(line) + "\n").encode())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    return 0 if flag.item() else 1


if __name__ == "__main__":
    sys.exit(main())
