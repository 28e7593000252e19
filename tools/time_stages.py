#!/usr/bin/env python3
"""Stage timings of one compress + decompress of the bench field (development aid; run under gpurun).
   python tools/time_stages.py [size] [reps]   -- honours the WRB_* environment switches of the library"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from waverange_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
field = bench.synth_field(torch, n, 1234, dev, torch.float32)
codec = api.Codec(device=0, stream=torch.cuda.current_stream().cuda_stream)
codec.set_timing(True)
_, cap = api.setup_wr(n, n, n)
cap = min(cap, n ** 3 * 5 + (1 << 20))
blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
rec = torch.empty(n ** 3, dtype=torch.float32, device=dev)
se, sd = [], []
for it in range(reps + 2):
    h = codec.encode_device(field.data_ptr(), api.F32, n, n, n, bench.TOL, blob.data_ptr(), cap)
    e = codec.stage_ms()
    codec.decode_device(rec.data_ptr(), api.F32, n, n, n, h, blob.data_ptr())
    d = codec.stage_ms()
    if it >= 2:
        se.append(e); sd.append(d)
err = ((rec.view(n, n, n).double() - field.double()).abs().max() / field.double().abs().max()).item()
tag = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("WRB_"))
print("[%s] enc %s | dec %s | err %.3e" % (tag, ["%.3f" % statistics.mean(x[i] for x in se) for i in range(4)],
                                           ["%.3f" % statistics.mean(x[i] for x in sd) for i in range(4)], err))
