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
build; single-threaded like the reference) on a bounded 256^3 sample of the same field.
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
SAMPLE = 256            # edge of the CPU-baseline sample cube when the whole workload would take too long (2 s per pass)
METRIC = "raw-field compress+decompress throughput (device-timed)"
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the forward-wavelet + quantise kernels of one compress of
# the default workload, from the ncu --set full capture profiles/r1j_ncu_raw_512.csv:
# forward levels 1-4: 1.578 + 0.249 + 0.017 + 0.002 GB, quantise 3 x (1.076 + 0.131) GB
NCU_TRAFFIC_512 = int((1.578 + 0.249 + 0.017 + 0.002 + 3 * (1.076 + 0.131)) * 1e9)
UNIT = "GB/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def synth_field(torch, n, seed, device, dtype, nm=48, expo=-5.0 / 6.0, nz_total=None, z0=0, nzl=None):
    """Turbulence-like field: sum of nm plane waves with random integer wavevectors, random phases,
    amplitude |k|^expo, evaluated in f64 on the device slab by slab through
    separable complex tables, then rounded to `dtype`.  Deterministic for a given seed."""
    rng = np.random.default_rng(seed)
    kmax = 24
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
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    n = N_FIELD if dev == "cuda" else SAMPLE
    fld = synth_field(torch, n, 1234, dev, torch.float32)
    # the whole 512^3 workload (17 s of single-threaded CPU work per step) when the run stays within a few minutes,
    # else the 256^3 corner of it
    edge = n if (n == N_FIELD and args.steps + args.warmup <= 10) else SAMPLE
    sample = fld[:edge, :edge, :edge].contiguous().cpu().numpy().astype(np.float64)
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
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (te + td) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "512^3 float32 turbulence-like field, tol 1e-4 (BASELINE.json configs[1])",
                       "sample": ("the whole field" if edge == N_FIELD else "%d^3 corner sub-cube of the same field" % edge)},
            "compress_gbs": nbytes / te / 1e9, "decompress_gbs": nbytes / td / 1e9,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                             "sample": "%d^3 f32-valued %s, encoding_wrap+decoding_wrap in process, 1 thread "
                                       "(the reference is single-threaded); host has %d cores"
                                       % (edge, "field (the whole workload)" if edge == N_FIELD else "sub-cube", os.cpu_count())},
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
    hooks = None
    if slab_mode:
        from waverange_b200 import slab
        hooks = slab.DistHooks(torch, dist, cuda=True)
        codec.set_slab(rank, world, hooks.halo_cb, hooks.reduce_cb)

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

    # ---- max over ranks ------------------------------------------------------------------------
    vals = torch.tensor([step_ms, statistics.mean(enc_ms), statistics.mean(dec_ms), e2e_step], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    step_ms, enc_mean, dec_mean, e2e_step = [float(v) for v in vals.tolist()]
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

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%d^3 float32 turbulence-like field, tol 1e-4%s" % (n, " (BASELINE.json configs[1])" if n == N_FIELD else " (profiling size)"),
                   "field_bytes": nbytes, "tolerance": TOL, "nlay": nlay, "ntot_enc": int(h.ntot_enc),
                   "ratio_vs_f32": nbytes / max(1, int(h.ntot_enc)), "chunk_symbols": 59999,
                   "fields": ("one %dx%dx%d field" % (n, n, n)) if world == 1 else
                             ("one %dx%dx%d field, z-slab partitioned over %d GPUs: NCCL halo exchange per level "
                              "(4 planes below / 3 above), all_reduce of extrema per layer" % (n, n, n * world, world)),
                   "l2": "working set (0.5 GB field + 2 GB scratch) exceeds the 126 MB L2; no explicit flush"},
        "compress_gbs": world * nbytes / (enc_mean * 1e-3) / 1e9,
        "decompress_gbs": world * nbytes / (dec_mean * 1e-3) / 1e9,
        "rel_linf_error": err / amax,
        "stages_ms": {"encode": dict(zip(["transform", "quantise", "range_encode", "assemble"], se)),
                      "decode": dict(zip(["parse", "range_decode", "dequantise", "inverse_transform"], sd)),
                      "encode_total": enc_mean, "decode_total": dec_mean},
        "roofline": {"bound": "hbm", "scope": "forward wavelet + quantise kernels of one compress (stage events)",
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     # dram__bytes_read+write of those kernels from the ncu --set full captures of this workload
                     # (profiles/r1j_ncu_raw_512.csv), per compress
                     "traffic": (NCU_TRAFFIC_512 if (n == N_FIELD and nlay == 3 and not slab_mode) else None),
                     "algorithmic_bytes": a_c, "peak_source": peak_src},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": world * (nbytes + int(h.ntot_enc)),
                "d2h_bytes_per_step": world * (nbytes + int(h.ntot_enc)), "ms_per_step": e2e_step,
                "api": "wrb_encode_host + wrb_decode_host (f32 pinned host buffers)" if world == 1 else
                       "pinned H2D + wrb_encode_slab_device + D2H, H2D + wrb_decode_slab_device + D2H per rank"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "host_wall_ms_per_step": wall_ms,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
        line["ratio_check"] = ratio_check
    emit(line)


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
