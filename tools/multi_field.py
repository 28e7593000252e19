#!/usr/bin/env python3
"""Independent fields (BASELINE.json configs[2]: the ux, uy, uz, p fields of a backup) on ONE GPU: one codec handle
and one stream per field, driven from one host thread each (the C ABI releases nothing it shares).  The coder kernels
occupy a tenth of the machine, so fields in flight on different streams overlap and the per-field latency of the
range coder is hidden.   python tools/multi_field.py [edge] [nfields] [f32|f64] [tol]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from waverange_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dt = torch.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else torch.float32
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-4
dev = torch.device("cuda", 0)
code = api.F64 if dt == torch.float64 else api.F32
fields = [bench.synth_field(torch, n, 100 + k, dev, dt, expo=(-7.0 / 6.0 if k == 3 else -5.0 / 6.0)) for k in range(nf)]
streams = [torch.cuda.Stream() for _ in range(nf)]
codecs = [api.Codec(device=0, stream=s.cuda_stream) for s in streams]
_, cap = api.setup_wr(n, n, n)
cap = min(cap, fields[0].numel() * fields[0].element_size() + (64 << 20))
blobs = [torch.empty(cap + 64, dtype=torch.uint8, device=dev) for _ in range(nf)]
recs = [torch.empty(n ** 3, dtype=dt, device=dev) for _ in range(nf)]
hdrs = [None] * nf


def job(k):
    hdrs[k] = codecs[k].encode_device(fields[k].data_ptr(), code, n, n, n, tol, blobs[k].data_ptr(), cap)
    codecs[k].decode_device(recs[k].data_ptr(), code, n, n, n, hdrs[k], blobs[k].data_ptr())


def run(parallel):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if parallel:
        ts = [threading.Thread(target=job, args=(k,)) for k in range(nf)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    else:
        for k in range(nf):
            job(k)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for _ in range(2):
    run(False); run(True)
ts = min(run(False) for _ in range(3))
tp = min(run(True) for _ in range(3))
nbytes = 2 * nf * fields[0].numel() * fields[0].element_size()
err = max(((recs[k].view_as(fields[k]).double() - fields[k].double()).abs().max() / fields[k].double().abs().max()).item() for k in range(nf))
print("%d x %d^3 %s tol %g: one after another %.2f ms (%.1f GB/s), %d streams %.2f ms (%.1f GB/s), nlay %s, max rel err %.2e"
      % (nf, n, str(dt).split(".")[1], tol, ts * 1e3, nbytes / ts / 1e9, nf, tp * 1e3, nbytes / tp / 1e9, [h.nlay for h in hdrs], err))
