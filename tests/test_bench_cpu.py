"""CPU tests of bench.py's host logic: the JSON line reaches the real stdout even while file descriptor 1 is parked
on stderr (NCCL's banner), and nvidia-smi samples are assigned to the timed region by their timestamps."""
import os

import pytest
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_emit_survives_a_parked_stdout(tmp_path):
    code = (
        "import os, sys\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "sys.stdout.flush()\n"
        "bench._REAL_STDOUT = os.dup(1)\n"
        "os.dup2(2, 1)\n"
        "os.write(1, b'NCCL version banner\\n')\n"
        "print('python-level chatter')\n"
        "bench.emit({'metric': 'x', 'value': 1})\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"metric": "x", "value": 1}\n'
    assert "NCCL version banner" in r.stderr and "python-level chatter" in r.stderr


def test_clock_samples_are_filtered_by_timestamp(tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    now = time.time()

    def stamp(t):
        lt = time.localtime(t)
        return time.strftime("%Y/%m/%d %H:%M:%S", lt) + ".%03d" % int((t % 1) * 1000)

    assert abs(bench.ClockSampler._stamp(stamp(now)) - now) < 0.002
    assert bench.ClockSampler._stamp("garbage") is None
    rows = [(now - 1.0, 1200, "Not Active"), (now + 0.010, 1965, "Not Active"), (now + 0.020, 1950, "Not Active"),
            (now + 0.030, 1965, "Active"), (now + 2.0, 900, "Not Active")]
    path = tmp_path / "clocks.csv"
    path.write_text("".join("%s, %d, 1965, 400.0, 0x0, Not Active, Not Active, Not Active, %s\n" % (stamp(t), mhz, cap)
                            for t, mhz, cap in rows))

    class Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

        def kill(self):
            pass

    s = bench.ClockSampler(0)
    s.proc, s.path = Done(), str(path)
    out = s.stop(now, now + 0.040)
    assert out["samples"] == 3 and out["samples_whole_run"] == 5 and out["window"] == "timed region"
    assert out["sm_mhz"] == 1965 and out["sm_max_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"]
    # a region shorter than the sampling interval falls back to its 50 ms surroundings and says so
    path.write_text("".join("%s, %d, 1965, 400.0, 0x0, Not Active, Not Active, Not Active, %s\n" % (stamp(t), mhz, cap)
                            for t, mhz, cap in rows))
    s = bench.ClockSampler(0)
    s.proc, s.path = Done(), str(path)
    out = s.stop(now + 0.012, now + 0.014)
    assert out["samples"] == 3 and out["window"].startswith("timed region +- 50 ms")


def test_roofline_traffic_comes_from_a_capture_of_the_current_kernels():
    """roofline.traffic is only printed while the ncu capture under profiles/ was taken from the kernel sources that are
    in the tree; the committed capture must be that one (recapture after touching quant.cu / wavelet_fused.cu), it must be
    the sum of the two kernel families in the committed CSV, and a capture of other sources must be ignored."""
    import csv
    import glob
    import json
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "profiles", "traffic_*.json"))
    assert files, "no traffic capture committed"
    traffic, src = bench.measured_traffic(512, 3)
    if not any(json.load(open(f)).get("kernels_sha") == bench.kernels_sha() for f in files):
        assert (traffic, src) == (None, None)          # kernels changed since the capture: the line must say null
        pytest.skip("the kernels changed since the last ncu traffic capture (bench.py prints traffic: null)")
    assert traffic is not None and src["kernels_sha"] == bench.kernels_sha()
    t = json.load(open(os.path.join(root, src["file"])))
    assert t["dram_bytes"] == traffic and 1.0 <= traffic / t["algorithmic_bytes"] < 1.15      # one DRAM pass, little else
    # the same figure from the raw CSV (dram__bytes_read + dram__bytes_write of the forward + quantise launches)
    rows = list(csv.reader(open(os.path.join(root, "profiles", t["source"].replace("r3n_", "r2_")))))
    hdr = next(r for r in rows if len(r) > 5 and r[0] == "ID")
    kn, mn, mv, mu = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = sum(float(r[mv].replace(",", "")) * scale[r[mu]] for r in rows
              if len(r) == len(hdr) and r[0] != "ID" and r[mn].startswith("dram__bytes_")
              and ("fwd_level_fused" in r[kn] or "quantise_kernel" in r[kn]))
    assert abs(tot - traffic) <= 1e-6 * traffic
    assert bench.measured_traffic(512, 4) == (None, None)                                      # another layer count: no capture
