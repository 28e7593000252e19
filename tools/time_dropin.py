#!/usr/bin/env python3
"""Wall-clock of the reference entry points (encoding_wrap / decoding_wrap of libwaverange_b200.so) on ordinary
(pageable) host arrays, the way a program written for the reference calls them, next to the pinned-buffer host API.
   python tools/time_dropin.py [edge]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from waverange_b200 import api  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
tol = 1e-4
dev = torch.device("cuda", 0)
f32 = bench.synth_field(torch, n, 1234, dev, torch.float32, nz_total=n, z0=0, nzl=n).cpu().numpy().reshape(n, n, n)
f64 = f32.astype(np.float64)                       # what a reference caller holds: a pageable double array
out = {"edge": n, "tolerance": tol}
for rep in range(3):
    t0 = time.perf_counter()
    h, data = api.encoding_wrap(f64, tol)
    t1 = time.perf_counter()
    rec = api.decoding_wrap((n, n, n), h, data)
    t2 = time.perf_counter()
    out["dropin_f64_pageable_ms"] = [round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)]
assert np.abs(rec - f64).max() <= tol * np.abs(f64).max()
c = api.Codec(device=0)
hp = torch.empty((n, n, n), dtype=torch.float64, pin_memory=True)
hp.copy_(torch.from_numpy(f64))
hb = torch.empty(f64.nbytes, dtype=torch.uint8, pin_memory=True)
hr = torch.empty((n, n, n), dtype=torch.float64, pin_memory=True)
for rep in range(3):
    t0 = time.perf_counter()
    h, data = c.encode_host(hp.numpy(), tol, out=hb.numpy())
    t1 = time.perf_counter()
    c.decode_host((n, n, n), h, data, out=hr.numpy())
    t2 = time.perf_counter()
    out["host_api_f64_pinned_ms"] = [round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)]
out["field_GB"] = f64.nbytes / 1e9
print(json.dumps(out))
