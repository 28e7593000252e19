#!/usr/bin/env python3
"""One warm-up and ONE measured compress + decompress of the benchmark field, for ncu captures (development aid).
    python tools/profile_encode.py [edge] [f32|f64] [tol] [encodes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from waverange_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
tol = float(sys.argv[3]) if len(sys.argv) > 3 else bench.TOL
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
code = api.F64 if dt == torch.float64 else api.F32
field = bench.synth_field(torch, n, 1234 if dt == torch.float32 else 5, dev, dt)
codec = api.Codec(device=0, stream=torch.cuda.current_stream().cuda_stream)
_, cap = api.setup_wr(n, n, n)
cap = min(cap, field.numel() * field.element_size() + (64 << 20))
blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
rec = torch.empty(n ** 3, dtype=dt, device=dev)
for it in range(reps):
    h = codec.encode_device(field.data_ptr(), code, n, n, n, tol, blob.data_ptr(), cap)
    codec.decode_device(rec.data_ptr(), code, n, n, n, h, blob.data_ptr())
torch.cuda.synchronize()
print("nlay %d ntot_enc %d" % (h.nlay, h.ntot_enc))
