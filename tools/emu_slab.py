#!/usr/bin/env python3
"""N ranks EMULATED on one GPU (one thread + codec per rank, waverange_b200.slab.LocalGroup) at a chosen size: the whole
z-slab pipeline incl. the global symbol order, for profiling the exchange kernels under ncu and for checking sizes that
need no real NVLink (development aid).   python tools/emu_slab.py [edge] [world] [nz] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from waverange_b200 import api, slab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nz = int(sys.argv[3]) if len(sys.argv) > 3 else n
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda", 0)
nzl = nz // world
grp = slab.LocalGroup(torch, world)
codecs = [api.Codec(device=0) for _ in range(world)]
fields = [bench.synth_field(torch, n, 1234, dev, torch.float32, nz_total=nz, z0=r * nzl, nzl=nzl) for r in range(world)]


def rank_fn(r, halo_cb, reduce_cb):
    c = codecs[r]
    c.set_slab(r, world, halo_cb, reduce_cb)
    c.set_slab_peers(codecs)
    c.set_timing(True)
    ntl = n * n * nzl
    cap = ntl * 3 + (1 << 20)
    blob = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    rec = torch.empty(ntl, dtype=torch.float32, device=dev)
    out = None
    for it in range(reps):
        h = c.encode_slab_device(fields[r].data_ptr(), api.F32, n, n, nz, r * nzl, nzl, bench.TOL, blob.data_ptr(), cap)
        se = c.stage_ms()
        c.decode_slab_device(rec.data_ptr(), api.F32, n, n, nz, r * nzl, nzl, h, blob.data_ptr())
        sd = c.stage_ms()
        out = (h.nlay, h.ntot_enc, se, sd, ((rec.view_as(fields[r]).double() - fields[r].double()).abs().max() / fields[r].double().abs().max()).item())
    return out


t0 = time.time()
res = grp.run(rank_fn)
for r, x in enumerate(res):
    print("rank %d nlay %d bytes %d enc %s dec %s err %.3e" % (r, x[0], x[1], ["%.3f" % v for v in x[2]], ["%.3f" % v for v in x[3]], x[4]))
print("wall %.2f s" % (time.time() - t0))
