#!/usr/bin/env python3
"""bench.py -- WaveRange hot path on B200: raw-field compress/decompress GB/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one compress followed by one decompress of the synthetic field of BASELINE.json
configs[1] (512^3 float32 turbulence-like field, relative tolerance 1e-4).  `value` is raw-field
GB/s (1 GB = 1e9 B), counting the field's native bytes once per direction:
    value = 2 * ntot * 4 B / (t_compress + t_decompress),
device-timed with CUDA events, field and coded stream resident in HBM.  `e2e` is the same metric
through the host-buffer C-ABI calls (wrb_encode_host / wrb_decode_host, pinned host memory, H2D and
D2H copies inside the timed region).  With N > 1 (torchrun) every rank codes its own field
(independent fields shard with no data-path collective) -> weak scaling.

--impl reference times the reference's own CPU implementation (oracle/_ref, stock-style FMA
build; single-threaded like the reference) on the whole 512^3 field (about 20 s per step on one core; with N > 1
on one rank's 512^3 share of the N-times larger field, a bounded sample).

At N = 1 the line also carries a `configs` block (outside the headline's timed region): the other configurations of
BASELINE.json -- C1 256^3 incl. the wall clock of the reference's wrenc/wrdec beside this repo's, C3 1024^3 float64
(one field and the four fields of a backup), C5 512^3 float64 at three tolerances -- each with device GB/s, the
fraction of the HBM roofline of its transform + quantise scope, layer count and size.  --no-configs skips it.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_FIELD = 512
TOL = 1e-4
SAMPLE = 256            # edge of the CPU-baseline sample cube, only used when the whole-field leg fails (host memory)
METRIC = "raw-field compress+decompress throughput (device-timed)"
KERNEL_SOURCES = ["wavelet_fused.cu", "wavelet_pairs.cuh", "wr_common.cuh", "quant.cu"]     # the roofline scope's kernels


def kernels_sha():
    """hash of the sources of the kernels the roofline is quoted on: an ncu traffic figure is only printed when it
    was captured from exactly these sources"""
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "waverange_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(n, nlay):
    """dram__bytes_read.sum + dram__bytes_write.sum of the forward-transform + quantise kernels of ONE compress, from the
    newest profiles/traffic_*.json (written by tools/ncu_traffic.py from an `ncu --set full` capture) that was taken
    on the current kernel sources at this size and layer count; None (-> null) otherwise."""
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json"))):
        try:
            t = json.load(open(p))
            t["_file"] = p
        except Exception:
            continue
        if t.get("kernels_sha") == kernels_sha() and t.get("edge") == n and t.get("nlay") == nlay:
            best = t
    if best is None:
        return None, None
    return int(best["dram_bytes"]), {"file": "profiles/" + os.path.basename(best["_file"]), "commit": best.get("commit"),
                                     "kernels_sha": best.get("kernels_sha")}


UNIT = "GB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_field(torch, n, seed, device, dtype, nm=48, expo=-5.0 / 6.0, nz_total=None, z0=0, nzl=None, kmax=24):
    """Turbulence-like field: sum of nm plane waves with random integer wavevectors, random phases,
    amplitude |k|^expo, evaluated in f64 on the device slab by slab through
    separable complex tables, then rounded to `dtype`.  Deterministic for a given seed."""
    rng = np.random.default_rng(seed)
    k = rng.integers(1, kmax + 1, size=(nm, 3)).astype(np.float64)
    k *= rng.choice([-1.0, 1.0], size=(nm, 3))
    ph = rng.uniform(0, 2 * np.pi, nm)
    amp = (k ** 2).sum(1) ** (expo / 2.0 * 1.0)
    x = np.arange(n, dtype=np.float64) / n
    nz_total = nz_total or n
    nzl = nzl or nz_total
    zc = (z0 + np.arange(nzl, dtype=np.float64)) / n        # same wave numbers per unit length along z
    tab = [np.exp(2j * np.pi * np.outer(k[:, 0], x)), np.exp(2j * np.pi * np.outer(k[:, 1], x)),
           np.exp(2j * np.pi * np.outer(k[:, 2], zc))]
    coef = torch.from_numpy(amp * np.exp(1j * ph)).to(device)
    X = torch.from_numpy(tab[0]).to(device)
    Y = torch.from_numpy(tab[1]).to(device)
    Z = torch.from_numpy(tab[2]).to(device)
    out = torch.empty((nzl, n, n), dtype=dtype, device=device)
    slab = 16
    for zs in range(0, nzl, slab):
        zc = Z[:, zs:zs + slab] * coef[:, None]                                  # (nm, slab)
        yz = torch.einsum("mz,my->mzy", zc, Y)                                   # (nm, slab, n)
        f = torch.einsum("mzy,mx->zyx", yz, X).imag
        out[zs:zs + slab] = f.to(dtype)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile(prefix="wrb_clocks_", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    @staticmethod
    def _stamp(text):
        """nvidia-smi timestamp 'YYYY/MM/DD HH:MM:SS.mmm' (local time) -> seconds since the epoch"""
        import datetime
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        """Median clock and throttle reasons of the samples taken inside [t0, t1] (time.time() at the ends of the timed
        region; nvidia-smi runs since before the warm-up, so it is sampling when the region starts).  A region shorter
        than the sampling interval can hold none: then the samples of the 50 ms around it are used and `window` says so."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                p = [s.strip() for s in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    rows.append((self._stamp(p[0]), float(p[1]), float(p[2]), p[5:9]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        window = "timed region"
        pick = rows
        if t0 is not None and t1 is not None and any(r[0] is not None for r in rows):
            pick = [r for r in rows if r[0] is not None and t0 <= r[0] <= t1]
            if len(pick) < 2:
                pick = [r for r in rows if r[0] is not None and t0 - 0.05 <= r[0] <= t1 + 0.05]
                window = "timed region +- 50 ms (region shorter than the sampling interval)"
        reasons = set()
        for r in pick:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if pick:
            out.update(sm_mhz=statistics.median([r[1] for r in pick]), sm_max_mhz=max(r[2] for r in pick),
                       reasons=sorted(reasons), samples=len(pick), window=window, samples_whole_run=len(rows))
        return out


def workload_config(n, world):
    """`config` of the JSON line: the same dict in both arms (it describes the workload, not the result)"""
    if world == 1:
        shape = "%d^3" % n
        fields = "one %dx%dx%d field" % (n, n, n)
    else:
        shape = "%dx%dx%d" % (n, n, n * world)
        fields = ("one %dx%dx%d field, z-slab partitioned over %d GPUs (%d planes each): NCCL halo exchange per level "
                  "(4 planes below / 3 above), all-reduce of the extrema per layer, symbol exchange over NVLink into the "
                  "global wavelet-space order (every chunk stream = the single-GPU run's)" % (n, n, n * world, world, n))
    tag = " (BASELINE.json configs[1])" if (n == N_FIELD and world == 1) else \
          (" (configs[1]'s field extended along z, weak scaling; configs[3]'s partition)" if n == N_FIELD else " (profiling size)")
    return {"workload": "%s float32 turbulence-like field, tol 1e-4%s" % (shape, tag),
            "field_bytes": n * n * n * world * 4, "tolerance": TOL, "fields": fields,
            "l2": "inputs (0.5 GB field per GPU, 1-2 GB of coefficients and symbols) exceed the 126 MB L2; no explicit flush"}


def reference_lib():
    from oracle.binding import Reference, Restatement
    if Reference.available("fma"):
        return Reference("fma"), "reference"
    if Reference.available("strict"):
        return Reference("strict"), "reference"
    return Restatement(), "port"


def cpu_time_sample(sample_f64, tol):
    """encode+decode of the sample on one host core with the reference's own code; returns seconds"""
    lib, kind = reference_lib()
    t0 = time.perf_counter()
    enc = lib.encode(sample_f64, tol)
    t1 = time.perf_counter()
    lib.decode(sample_f64.shape, enc["header"], enc["data"])
    t2 = time.perf_counter()
    return kind, t1 - t0, t2 - t1, enc["header"]


def run_reference(args, rank, world):
    """The reference's own code (oracle/_ref) on the host, single-threaded like the reference, on the WHOLE 512^3 field of
    the N = 1 workload -- the same config as our arm.  With N > 1 our arm codes an N-times larger field in z-slabs; the
    reference (about 20 s per 512^3 on one core) then codes one rank's 512^3 share of it per step: a bounded sample."""
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n = args.size if dev == "cuda" else min(args.size, SAMPLE)      # without a GPU (this container): a small smoke run
    fld = synth_field(torch, n, 1234, dev, torch.float32, nz_total=n * world, z0=0, nzl=n)
    sample = fld.contiguous().cpu().numpy().astype(np.float64)
    del fld
    times = []
    kind = "port"
    for it in range(args.warmup + args.steps):
        kind, te, td, h = cpu_time_sample(sample, TOL)
        if it >= args.warmup:
            times.append((te, td))
    te = statistics.mean(t[0] for t in times)
    td = statistics.mean(t[1] for t in times)
    nbytes = sample.size * 4
    val = 2 * nbytes / (te + td) / 1e9
    what = ("the whole %d^3 field" % n) if world == 1 else \
           ("rank 0's %d^3 slab of the %dx%dx%d field (1/%d of the workload per step)" % (n, n, n, n * world, world))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (te + td) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "input_dtype": "f32", "data": "synthetic",
            "config": workload_config(n, world),
            "result": {"nlay": int(h.nlay), "ntot_enc": int(h.ntot_enc), "ratio_vs_f32": nbytes / max(1, int(h.ntot_enc))},
            "compress_gbs": nbytes / te / 1e9, "decompress_gbs": nbytes / td / 1e9,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": "%s, encoding_wrap %.2f s + decoding_wrap %.2f s in process, 1 thread (the "
                                       "reference is single-threaded); host has %d cores" % (what, te, td, os.cpu_count())},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from waverange_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.size
    nz = ny = nx = n
    ntot = n ** 3
    slab_mode = world > 1
    nz_total, z0 = n * world, n * rank       # N > 1: ONE field of n x n x (n*N), z-slab partitioned (weak scaling)
    field = synth_field(torch, n, 1234, dev, torch.float32, nz_total=nz_total, z0=z0, nzl=n)
    nbytes = ntot * 4

    sampler = ClockSampler(local_rank)          # started before the warm-up: nvidia-smi needs ~0.1 s to produce its first line
    if rank == 0:
        sampler.start()
    stream = torch.cuda.current_stream()
    codec = api.Codec(device=local_rank, stream=stream.cuda_stream)
    codec.set_timing(True)
    if slab_mode:
        # every collective is issued by the library (NCCL from C++); torch.distributed only carries the 128-byte NCCL id.
        # The ranks code the GLOBAL symbol order: each chunk stream is the one a single GPU (and the reference) produces.
        from waverange_b200 import slab
        slab.set_comm_from_dist(codec, torch, dist, dev)

    def encode_dev(src_ptr, dst_ptr):
        if slab_mode:
            return codec.encode_slab_device(src_ptr, api.F32, nx, ny, nz_total, z0, nz, TOL, dst_ptr, cap)
        return codec.encode_device(src_ptr, api.F32, nx, ny, nz, TOL, dst_ptr, cap)

    def decode_dev(dst_ptr, hh, src_ptr):
        if slab_mode:
            return codec.decode_slab_device(dst_ptr, api.F32, nx, ny, nz_total, z0, nz, hh, src_ptr)
        return codec.decode_device(dst_ptr, api.F32, nx, ny, nz, hh, src_ptr)
    _, cap = api.setup_wr(nx, ny, nz)
    cap = min(cap, ntot * 5 + (1 << 20))          # f32 field at tol 1e-4 codes to < 2 B/pt; keep HBM use modest
    blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    recon = torch.empty(ntot, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    enc_ms, dec_ms, stage_enc, stage_dec = [], [], [], []
    h = None
    # ---- device-resident arm ----------------------------------------------------------------
    for it in range(args.warmup):
        h = encode_dev(field.data_ptr(), blob.data_ptr())
        decode_dev(recon.data_ptr(), h, blob.data_ptr())
    barrier()
    launches0 = codec.launch_count()
    t_wall0 = time.perf_counter()
    t_region0 = time.time()
    tot_ev0 = torch.cuda.Event(enable_timing=True)
    tot_ev1 = torch.cuda.Event(enable_timing=True)
    tot_ev0.record(stream)
    for it in range(args.steps):
        ev[0].record(stream)
        h = encode_dev(field.data_ptr(), blob.data_ptr())
        se = codec.stage_ms()
        ev[1].record(stream)
        decode_dev(recon.data_ptr(), h, blob.data_ptr())
        sd = codec.stage_ms()
        ev[2].record(stream)
        torch.cuda.synchronize()
        enc_ms.append(ev[0].elapsed_time(ev[1])); dec_ms.append(ev[1].elapsed_time(ev[2]))
        stage_enc.append(se); stage_dec.append(sd)
    tot_ev1.record(stream)
    barrier()
    launches = codec.launch_count() - launches0
    step_ms = tot_ev0.elapsed_time(tot_ev1) / args.steps
    wall_ms = (time.perf_counter() - t_wall0) * 1e3 / args.steps
    clocks = sampler.stop(t_region0, time.time()) if rank == 0 else None

    # N > 1, for context only: the same steps in the rank-local symbol order (no symbol exchange; the pieces are then NOT
    # the reference encoder's chunk streams of the whole field) -- what the global order costs
    local_ms = float('nan')
    if slab_mode and not args.no_check:
        codec.set_slab_order(False)
        for it in range(2):
            hl = encode_dev(field.data_ptr(), blob.data_ptr())
            decode_dev(recon.data_ptr(), hl, blob.data_ptr())
        barrier()
        le0, le1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        le0.record(stream)
        for it in range(args.steps):
            hl = encode_dev(field.data_ptr(), blob.data_ptr())
            decode_dev(recon.data_ptr(), hl, blob.data_ptr())
        le1.record(stream)
        barrier()
        local_ms = le0.elapsed_time(le1) / args.steps
        codec.set_slab_order(True)
        h = encode_dev(field.data_ptr(), blob.data_ptr())            # blob / recon back to the global-order result
        decode_dev(recon.data_ptr(), h, blob.data_ptr())
        barrier()

    # correctness guard inside the bench: the reconstruction meets the requested tolerance
    errt = torch.stack([(recon.view(n, n, n).double() - field.double()).abs().max(), field.double().abs().max()])
    if world > 1:
        dist.all_reduce(errt, op=dist.ReduceOp.MAX)
    err, amax = errt[0].item(), errt[1].item()
    assert err <= TOL * amax * 1.0000001, "reconstruction error %.3e exceeds tolerance" % (err / amax)

    # ---- end-to-end arm: host buffers through the C ABI (pinned memory, copies timed) ---------
    h_field = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
    h_field.copy_(field)
    h_blob = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    h_rec = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
    np_field, np_blob, np_rec = h_field.numpy(), h_blob.numpy(), h_rec.numpy()
    e2e_ms = []
    for it in range(0 if args.no_e2e else max(1, min(args.warmup, 2)) + args.steps):
        barrier()
        t0 = time.perf_counter()
        if slab_mode:     # same traffic as wrb_encode_host / wrb_decode_host: H2D field, D2H stream, H2D stream, D2H field
            d_in = h_field.to(dev, non_blocking=True)
            hh = encode_dev(d_in.data_ptr(), blob.data_ptr())
            h_blob[:hh.ntot_enc].copy_(blob[:hh.ntot_enc], non_blocking=True)
            torch.cuda.synchronize()
            blob[:hh.ntot_enc].copy_(h_blob[:hh.ntot_enc], non_blocking=True)
            decode_dev(recon.data_ptr(), hh, blob.data_ptr())
            h_rec.copy_(recon.view(n, n, n), non_blocking=True)
        else:
            hh, data = codec.encode_host(np_field, TOL, out=np_blob)
            codec.decode_host((n, n, n), hh, data, out=np_rec)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if it >= max(1, min(args.warmup, 2)):
            e2e_ms.append((t1 - t0) * 1e3)
    e2e_step = statistics.mean(e2e_ms) if e2e_ms else float('nan')
    if e2e_ms:
        assert np.array_equal(np_rec.ravel()[:4096], recon[:4096].cpu().numpy())
    # what the host link gives each rank when all ranks copy at once (explains the e2e scaling: the ranks share the host's
    # memory system and PCIe root ports, the codec itself scales with the device-timed value)
    copy_gbs = None
    if e2e_ms and slab_mode:
        d_tmp = torch.empty_like(field)
        t_copy = []
        for direction in (0, 1):
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                if direction == 0:
                    d_tmp.copy_(h_field, non_blocking=True)
                else:
                    h_rec.copy_(d_tmp, non_blocking=True)
            torch.cuda.synchronize()
            t_copy.append((time.perf_counter() - t0) / 3)
        tc = torch.tensor(t_copy, device=dev, dtype=torch.float64)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        copy_gbs = {"h2d_per_rank_min": nbytes / tc[0].item() / 1e9, "d2h_per_rank_min": nbytes / tc[1].item() / 1e9,
                    "note": "pinned 0.5 GB per rank, all ranks copying at the same time"}
        del d_tmp

    # ---- N > 1: the chunk streams of all ranks against a single-GPU encode of the same (whole) field on rank 0 --------
    streams_equal = None
    if slab_mode and not args.no_check:
        from waverange_b200 import slab
        crcs = slab.stream_crcs(h, blob[:h.ntot_enc].cpu().numpy())
        allc = [None] * world
        dist.gather_object((crcs, list(h.deps_vec), list(h.minval_vec), int(h.nlay)), allc if rank == 0 else None, dst=0)
        if rank == 0:
            whole = synth_field(torch, n, 1234, dev, torch.float32, nz_total=nz_total, z0=0, nzl=nz_total)
            c1 = api.Codec(device=local_rank, stream=stream.cuda_stream)
            cap1 = whole.numel() * 3 + (1 << 20)
            blob1 = torch.empty(cap1 + 64, dtype=torch.uint8, device=dev)
            h1 = c1.encode_device(whole.data_ptr(), api.F32, nx, ny, nz_total, TOL, blob1.data_ptr(), cap1)
            want = slab.stream_crcs(h1, blob1[:h1.ntot_enc].cpu().numpy())
            got = [[x for r in range(world) for x in allc[r][0][l]] for l in range(int(h1.nlay))] if all(a[3] == h1.nlay for a in allc) else None
            streams_equal = {"chunk_streams_equal_single_gpu": bool(got == want),
                             "header_doubles_equal_single_gpu": bool(all(a[1] == list(h1.deps_vec) and a[2] == list(h1.minval_vec) for a in allc)),
                             "chunk_streams": sum(len(x) for x in want), "single_gpu_bytes": int(h1.ntot_enc)}
            c1.close()
            del whole, blob1
        dist.barrier()

    # ---- max over ranks ------------------------------------------------------------------------
    vals = torch.tensor([step_ms, statistics.mean(enc_ms), statistics.mean(dec_ms), e2e_step, local_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    step_ms, enc_mean, dec_mean, e2e_step, local_ms = [float(v) for v in vals.tolist()]
    if rank != 0:
        return

    hbm, peak_src = peaks()
    nlay = int(h.nlay)
    se = [statistics.mean(s[i] for s in stage_enc) for i in range(4)]
    sd = [statistics.mean(s[i] for s in stage_dec) for i in range(4)]
    a_c = ntot * (4 + 8 + 9 * nlay)                    # algorithmic bytes, SURVEY.md section 8d
    t_wq = (se[0] + se[1]) * 1e-3
    achieved = a_c / t_wq / 1e9
    value = world * 2 * nbytes / (step_ms * 1e-3) / 1e9
    e2e_val = world * 2 * nbytes / (e2e_step * 1e-3) / 1e9

    # CPU baseline (rank 0, N = 1 only): the reference's own code on the WHOLE 512^3 workload -- about 20 s on one core, the
    # reference being single-threaded -- or on a 256^3 corner of a profiling-size field
    cpu = None

    def cpu_leg(edge):
        what = "the whole %d^3 field" % n if edge == n else "%d^3 sub-cube of the same field" % edge
        sample = field.view(n, n, n)[:edge, :edge, :edge].contiguous().cpu().numpy().astype(np.float64)
        kind, te, td, href = cpu_time_sample(sample, TOL)
        # the same data through the GPU path: coded size against the reference's single-stream layers
        hs, _ = codec.encode_host(sample.astype(np.float32), TOL)
        rc = {"sample": what, "reference_bytes": int(href.ntot_enc), "ours_bytes": int(hs.ntot_enc),
              "size_overhead": hs.ntot_enc / max(1, href.ntot_enc) - 1.0, "nlay_equal": int(hs.nlay) == int(href.nlay)}
        return rc, {"value": 2 * sample.size * 4 / (te + td) / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                    "sample": "%s, encoding_wrap %.2f s + decoding_wrap %.2f s, 1 thread "
                              "(reference is single-threaded); host has %d cores" % (what, te, td, os.cpu_count())}

    if world == 1 and not args.no_cpu:
        try:
            ratio_check, cpu = cpu_leg(n if n == N_FIELD else min(n, SAMPLE))
        except Exception as exc:                      # e.g. host memory: the bounded sample instead, and say so
            print("bench: cpu_baseline on the whole field failed (%r); using the %d^3 sample" % (exc, SAMPLE), file=sys.stderr)
            ratio_check, cpu = cpu_leg(min(n, SAMPLE))

    traffic, traffic_src = measured_traffic(n, nlay) if not slab_mode else (None, None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "input_dtype": "f32",          # arithmetic of the path (transform + quantiser in f64, coder in u32) / the field's own type
        "data": "synthetic",
        "config": workload_config(n, world),
        "result": {"nlay": nlay, "ntot_enc": int(h.ntot_enc), "ratio_vs_f32": nbytes / max(1, int(h.ntot_enc)),
                   "chunk_symbols": 59999, "layer_guess_misses": codec.layer_guess_misses()},
        "compress_gbs": world * nbytes / (enc_mean * 1e-3) / 1e9,
        "decompress_gbs": world * nbytes / (dec_mean * 1e-3) / 1e9,
        "rel_linf_error": err / amax,
        "stages_ms": {"encode": dict(zip(["transform", "quantise", "range_encode", "assemble"], se)),
                      "decode": dict(zip(["parse", "range_decode", "dequantise", "inverse_transform"], sd)),
                      "encode_total": enc_mean, "decode_total": dec_mean},
        "roofline": {"bound": "hbm", "scope": "forward wavelet + quantise kernels of one compress (stage events)",
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     # dram__bytes_read+write of those kernels per compress from an ncu --set full capture of THESE kernel
                     # sources (profiles/traffic_*.json, tools/ncu_traffic.py); null when no capture matches them
                     "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes": a_c, "peak_source": peak_src,
                     "decode_scope": {"what": "inverse transform incl. dequantiser (A_d = ntot * (nlay + 4))",
                                      "achieved": ntot * (nlay + 4) / max(1e-9, sd[3] * 1e-3) / 1e9,
                                      "frac": ntot * (nlay + 4) / max(1e-9, sd[3] * 1e-3) / 1e9 / hbm},
                     "coder_scope": {"what": "range coder, A_rc = ntot * nlay + ntot_enc (no HBM target: serial chains)",
                                     "encode_frac": (ntot * nlay + int(h.ntot_enc)) / max(1e-9, se[2] * 1e-3) / 1e9 / hbm,
                                     "decode_frac": (ntot * nlay + int(h.ntot_enc)) / max(1e-9, sd[1] * 1e-3) / 1e9 / hbm}},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": world * (nbytes + int(h.ntot_enc)),
                "d2h_bytes_per_step": world * (nbytes + int(h.ntot_enc)), "ms_per_step": e2e_step,
                "api": "wrb_encode_host + wrb_decode_host (f32 pinned host buffers)" if world == 1 else
                       "pinned H2D + wrb_encode_slab_device + D2H, H2D + wrb_decode_slab_device + D2H per rank",
                "host_link": copy_gbs},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "host_wall_ms_per_step": wall_ms,
    }
    if local_ms == local_ms:
        line["rank_local_order"] = {"value": world * 2 * nbytes / (local_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": local_ms,
                                    "note": "same steps without the symbol exchange (round 1's mode): pieces are not the "
                                            "reference's chunk streams of the whole field; not the headline"}
    if streams_equal is not None:
        line["global_order_check"] = streams_equal
        line["nccl_rank0"] = codec.comm_counters()
    if cpu is not None:
        line["cpu_baseline"] = cpu
        line["ratio_check"] = ratio_check
    if world == 1 and not args.no_configs and n == N_FIELD:
        del field, recon, blob
        codec.trim()
        torch.cuda.empty_cache()
        try:
            line["configs"] = other_configs(torch, codec, dev, hbm)
        except Exception as exc:                          # the headline line must not be lost to a side measurement
            line["configs"] = {"error": repr(exc)}
    emit(line)


def _time_field(torch, codec, api, field, n, tol, hbm, reps=3, warm=1):
    """device-timed compress + decompress of one n^3 field already in HBM: means over `reps`, stage times, roofline
    fraction of the forward-transform + quantise scope (A_c) and of the inverse (A_d)"""
    code = api.F64 if field.dtype == torch.float64 else api.F32
    esz = field.element_size()
    ntot = n ** 3
    _, cap = api.setup_wr(n, n, n)
    cap = min(cap, ntot * esz + (64 << 20))
    blob = torch.empty(cap + 64, dtype=torch.uint8, device=field.device)
    rec = torch.empty(ntot, dtype=field.dtype, device=field.device)
    stream = torch.cuda.current_stream()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    enc, dec, se, sd = [], [], [], []
    h = None
    for it in range(warm + reps):
        ev[0].record(stream)
        h = codec.encode_device(field.data_ptr(), code, n, n, n, tol, blob.data_ptr(), cap)
        a = codec.stage_ms()
        ev[1].record(stream)
        codec.decode_device(rec.data_ptr(), code, n, n, n, h, blob.data_ptr())
        b = codec.stage_ms()
        ev[2].record(stream)
        torch.cuda.synchronize()
        if it >= warm:
            enc.append(ev[0].elapsed_time(ev[1])); dec.append(ev[1].elapsed_time(ev[2])); se.append(a); sd.append(b)
    amax = field.abs().max().item()
    err = 0.0
    step = max(1, n // 8) * n * n
    fv = field.view(-1)
    for a0 in range(0, ntot, step):
        err = max(err, (rec[a0:a0 + step].double() - fv[a0:a0 + step].double()).abs().max().item())
    nlay = int(h.nlay)
    te, td = statistics.mean(enc), statistics.mean(dec)
    t_wq = statistics.mean(x[0] + x[1] for x in se) * 1e-3
    t_inv = statistics.mean(x[3] for x in sd) * 1e-3
    a_c, a_d = ntot * (esz + 8 + 9 * nlay), ntot * (nlay + esz)
    nch = (ntot + 59998) // 59999
    nseek = max(0, (int(h.len_enc_vec[0]) and int.from_bytes(blob[28:32].cpu().numpy().tobytes(), "little")))
    out = {"compress_gbs": ntot * esz / te / 1e6, "decompress_gbs": ntot * esz / td / 1e6,
           "value_gbs": 2 * ntot * esz / (te + td) / 1e6, "compress_ms": te, "decompress_ms": td,
           "nlay": nlay, "ntot_enc": int(h.ntot_enc), "ratio": ntot * esz / max(1, int(h.ntot_enc)),
           "rel_linf_error": err / amax,
           "stages_ms": {"transform": statistics.mean(x[0] for x in se), "quantise": statistics.mean(x[1] for x in se),
                         "range_encode": statistics.mean(x[2] for x in se), "assemble": statistics.mean(x[3] for x in se),
                         "range_decode": statistics.mean(x[1] for x in sd), "inverse_transform": statistics.mean(x[3] for x in sd)},
           "roofline_frac": a_c / t_wq / 1e9 / hbm, "roofline_frac_inverse": a_d / t_inv / 1e9 / hbm,
           # the container's own bytes on top of the chunk streams (layer header, chunk lengths, seek entries): what it
           # adds to the reference's single-stream layers besides ~7 bytes of framing per chunk
           "container_table_bytes": nlay * (32 + nch * (4 + 10 * nseek)), "seek_points": nseek}
    del blob, rec
    return out


def other_configs(torch, codec, dev, hbm):
    """BASELINE.json configs other than the headline's, measured after it (device-timed like `value`; C1 also through
    the executables beside the reference's).  The reference's single-threaded CPU code is timed only where it takes
    seconds (C1); C2's is the headline's cpu_baseline."""
    import shutil
    from waverange_b200 import api
    out = {}
    # ---- C1: 256^3 float32, tol 1e-5: device, and wrenc/wrdec wall clock against the reference's executables -------
    n, tol = 256, 1e-5
    f = synth_field(torch, n, 1234, dev, torch.float32)
    c1 = _time_field(torch, codec, api, f, n, tol, hbm)
    c1["workload"] = "256^3 float32, tol 1e-5, C layout (configs[0])"
    bin_dir = os.path.join(ROOT, "waverange_b200", "bin")
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    tmp = tempfile.mkdtemp(prefix="wrb_c1_")
    try:
        f.cpu().numpy().tofile(os.path.join(tmp, "data.bin"))
        enc_args = ["data.bin", "data.wrb", "data.wrh", "2", "0", "1", "1", str(n), str(n), str(n), str(tol)]
        dec_args = ["data.wrb", "data.wrh", "datarec.bin", "2", "0"]

        def cli(enc, dec, sub):
            d = os.path.join(tmp, sub)
            os.makedirs(d)
            os.symlink(os.path.join(tmp, "data.bin"), os.path.join(d, "data.bin"))
            t0 = time.perf_counter()
            subprocess.run([enc] + enc_args, cwd=d, check=True, stdout=subprocess.DEVNULL)
            t1 = time.perf_counter()
            subprocess.run([dec] + dec_args, cwd=d, check=True, stdout=subprocess.DEVNULL)
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1, os.path.getsize(os.path.join(d, "data.wrb")), os.path.join(d, "datarec.bin")
        if os.path.exists(os.path.join(bin_dir, "wrenc")):
            cli(os.path.join(bin_dir, "wrenc"), os.path.join(bin_dir, "wrdec"), "warm")        # page cache, driver start-up
            te, td, size, rec_ours = cli(os.path.join(bin_dir, "wrenc"), os.path.join(bin_dir, "wrdec"), "ours")
            c1["cli"] = {"wrenc_s": te, "wrdec_s": td, "wrb_bytes": size,
                         "note": "process start + CUDA context creation + file I/O + codec, wall clock"}
            if os.path.exists(os.path.join(ref_dir, "wrenc_ref")):
                re_, rd, rsize, rec_ref = cli(os.path.join(ref_dir, "wrenc_ref"), os.path.join(ref_dir, "wrdec_ref"), "ref")
                c1["cli_reference"] = {"wrenc_s": re_, "wrdec_s": rd, "wrb_bytes": rsize, "cores": 1,
                                       "note": "the reference's own executables (oracle/_ref, unmodified sources), same file"}
                c1["cli"]["size_vs_reference"] = size / rsize - 1.0
                c1["cli"]["reconstruction_equal_to_reference"] = open(rec_ours, "rb").read() == open(rec_ref, "rb").read()
                c1["cli"]["speedup_wall"] = (re_ + rd) / (te + td)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out["C1"] = c1
    del f
    # ---- C5: 512^3 float64 at both ends and the middle of the tolerance sweep ----------------------------------------
    n = 512
    f = synth_field(torch, n, 5, dev, torch.float64)
    c5 = {"workload": "512^3 float64, tolerance sweep (configs[4]); 1e-16 is the 8-layer coder stress case", "points": {}}
    for tol in (1e-3, 1e-8, 1e-16):
        c5["points"]["%g" % tol] = _time_field(torch, codec, api, f, n, tol, hbm, reps=2)
    out["C5"] = c5
    del f
    codec.trim()
    torch.cuda.empty_cache()
    # ---- C3: four 1024^3 float64 fields (ux, uy, uz, p), tol 1e-8, one after another through one handle ------------
    n, tol = 1024, 1e-8
    per = []
    for k in range(4):
        f = synth_field(torch, n, 100 + k, dev, torch.float64, expo=(-7.0 / 6.0 if k == 3 else -5.0 / 6.0))
        per.append(_time_field(torch, codec, api, f, n, tol, hbm, reps=2 if k == 0 else 1, warm=1 if k == 0 else 0))
        del f
    tot_ms = sum(p["compress_ms"] + p["decompress_ms"] for p in per)
    out["C3"] = {"workload": "4 x 1024^3 float64 (ux, uy, uz, p), tol 1e-8 (configs[2]), fields back to back",
                 "one_field": per[0], "four_fields_value_gbs": 2 * 4 * n ** 3 * 8 / tot_ms / 1e6,
                 "four_fields_ms": tot_ms, "nlay": [p["nlay"] for p in per], "ratio": [p["ratio"] for p in per],
                 "note": "independent fields: with one GPU per field (4 GPUs) they run concurrently with no data-path "
                         "collective; tools/multi_field.py overlaps them on one GPU with one stream per field"}
    out["C4"] = {"workload": "2048^3 float32 in z-slabs over 2/4/8 GPUs (configs[3])",
                 "note": "needs several GPUs: `bench.py --gpus N` (weak-scaling stand-in, driver's SCALE run) and "
                         "tools/run_c4.py; results archived under profiles/"}
    codec.trim()
    return out


_REAL_STDOUT = None


def emit(line):
    """the one JSON line, on the process's real stdout"""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--size", type=int, default=N_FIELD, help="field edge (default 512 = BASELINE configs[1]; other sizes are for profiling only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (profiling only)")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (C1, C3, C5 beside the headline)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the comparison of the chunk streams with a single-GPU encode")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        # stdout must carry exactly one JSON line.  NCCL prints its version banner straight to file descriptor 1 (the
        # torch-bundled build does so whatever NCCL_DEBUG_FILE says), so descriptor 1 points at stderr while the run
        # lasts and the JSON line goes to the saved descriptor (emit()).
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
