#!/usr/bin/env python3
"""A/B timing of library variants in ONE process (development aid; run under gpurun).

    python tools/ab_stages.py CASE[,CASE...] VARIANT[;VARIANT...] [reps]

CASE     = edge:dtype:tol         e.g. 512:f32:1e-4  1024:f64:1e-8
VARIANT  = comma-separated WRB_* settings, "-" for the defaults, e.g.  "-;WRB_INV_IMPL=fused;WRB_INV_NOTMA=1"
Prints the stage times of every (case, variant) and a checksum of the coded bytes and of the reconstruction: variants
of one case must agree bit for bit."""
import os
import statistics
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from waverange_b200 import api  # noqa: E402

cases = sys.argv[1].split(",")
variants = sys.argv[2].split(";") if len(sys.argv) > 2 else ["-"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
for case in cases:
    edge, dt, tol = case.split(":")
    n, tol = int(edge), float(tol)
    tdt = torch.float32 if dt == "f32" else torch.float64
    code = api.F32 if dt == "f32" else api.F64
    field = bench.synth_field(torch, n, 1234 if dt == "f32" else 5, dev, tdt)
    _, cap = api.setup_wr(n, n, n)
    cap = min(cap, field.numel() * field.element_size() + (64 << 20))
    blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    rec = torch.empty(n ** 3, dtype=tdt, device=dev)
    for var in variants:
        saved = {}
        if var != "-":
            for kv in var.split(","):
                k, v = kv.split("=")
                saved[k] = os.environ.get(k)
                os.environ[k] = v
        codec = api.Codec(device=0, stream=torch.cuda.current_stream().cuda_stream)
        codec.set_timing(True)
        se, sd, te, td = [], [], [], []
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for it in range(reps + 2):
            ev[0].record()
            h = codec.encode_device(field.data_ptr(), code, n, n, n, tol, blob.data_ptr(), cap)
            e = codec.stage_ms()
            ev[1].record()
            codec.decode_device(rec.data_ptr(), code, n, n, n, h, blob.data_ptr())
            d = codec.stage_ms()
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                se.append(e); sd.append(d); te.append(ev[0].elapsed_time(ev[1])); td.append(ev[1].elapsed_time(ev[2]))
        crc_b = zlib.crc32(blob[:h.ntot_enc].cpu().numpy().tobytes())
        crc_r = 0
        step = max(1, n // 8) * n * n
        for a0 in range(0, n ** 3, step):
            crc_r = zlib.crc32(rec[a0:a0 + step].cpu().numpy().tobytes(), crc_r)
        m = lambda xs, i: statistics.mean(x[i] for x in xs)
        print("%-16s %-44s nlay %d enc %.3f [fwd %.3f q %.3f rc %.3f asm %.3f] dec %.3f [parse %.3f rd %.3f deq %.3f inv %.3f] "
              "crc blob %08x rec %08x" % (case, var, h.nlay, statistics.mean(te), m(se, 0), m(se, 1), m(se, 2), m(se, 3),
                                          statistics.mean(td), m(sd, 0), m(sd, 1), m(sd, 2), m(sd, 3), crc_b, crc_r), flush=True)
        codec.close()
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    del field, blob, rec
    torch.cuda.empty_cache()
