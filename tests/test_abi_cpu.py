"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the public
headers declare, and its host-only entry points agree with the oracle.  No compute call needs a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    names = set(re.findall(r"\b(wrb_[a-z0-9_]+|encoding_wrap(?:_f)?|decoding_wrap(?:_f)?|setup_wr(?:_f)?)\s*\(", txt))
    return {n for n in names if n not in ("wrb_codec", "wrb_header", "wrb_field_desc", "wrb_field_record", "wrb_mssg_ctl", "wrb_mssg_nmlst")}


def test_library_exports_every_declared_symbol(product_lib):
    lib = C.CDLL(product_lib)
    want = declared_functions("waverange_b200.h") | declared_functions("waverange.h") | declared_functions("waverange_files.h") | declared_functions("waverange_mssg.h")
    assert len(want) >= 40
    missing = [n for n in sorted(want) if not hasattr(lib, n)]
    assert not missing, missing
    from waverange_b200 import api
    assert want == set(api.EXPORTS)


def test_setup_wr_matches_reference_formula(product_lib, oracle):
    from waverange_b200 import api
    for n in [(1, 1, 1), (10, 10, 10), (256, 256, 256), (2048, 2048, 2048), (7, 11, 13)]:
        nlaymax, cap = api.setup_wr(*n)
        ntot = n[0] * n[1] * n[2]
        assert nlaymax == 8 and cap == 8 * max(1024, ntot)     # wrappers.cpp:531-541


def test_ind_p2w_matches_golden(product_lib, golden):
    from waverange_b200 import api
    for row in golden["p2w"][::5]:
        n, i, want = row[:3], row[3:6], row[6:]
        assert api.ind_p2w_3d(4, tuple(int(v) for v in n), tuple(int(v) for v in i)) == tuple(int(v) for v in want)


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device the product refuses to run instead of silently using a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from waverange_b200 import api
    with pytest.raises(api.WaveRangeError):
        api.Codec(device=0)


def test_product_does_not_reference_oracle():
    """The oracle is test infrastructure: nothing under waverange_b200/ may import or link it."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "waverange_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"\boracle\b|wr_oracle|_ref/", txt) and f != "api.py":
                    bad.append(f)
                if f == "api.py" and re.search(r"import\s+oracle|from\s+oracle|wr_oracle", txt):
                    bad.append(f)
    assert not bad, bad


def test_public_headers_compile_as_plain_c(tmp_path):
    """the device / file / MSSG headers are a C ABI: they must compile as C99 (waverange.h carries the reference's C++
    reference parameters and is checked as C++)"""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    inc = os.path.join(ROOT, "include")
    for h in ("waverange_b200.h", "waverange_files.h", "waverange_mssg.h"):
        src = tmp_path / (h + ".c")
        src.write_text('#include "%s"\nint main(void) { return 0; }\n' % os.path.join(inc, h))
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    src = tmp_path / "ref.cpp"
    src.write_text('#include "%s"\nint main() { return 0; }\n' % os.path.join(inc, "waverange.h"))
    r = subprocess.run(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_reference_front_ends_link_against_the_library_and_refuse_to_run_without_a_gpu(product_lib, tmp_path):
    """oracle/_ref/*_dropin = the reference's unmodified front-end sources linked against libwaverange_b200.so: the
    link succeeds (every symbol they need is exported with the reference's ABI) and, without a CUDA device, the run
    fails loudly instead of falling back to a CPU path"""
    import subprocess
    import torch
    from oracle import build_oracle
    if os.path.isdir("/root/reference/src"):
        build_oracle.build_dropin()
    exe = os.path.join(ROOT, "oracle", "_ref", "wrenc_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/wrenc_dropin not built")
    r = subprocess.run(["ldd", exe], capture_output=True, text=True)
    assert "libwaverange_b200.so" in r.stdout and "not found" not in r.stdout
    if torch.cuda.is_available():
        return
    np.arange(16 ** 3, dtype=np.float64).tofile(tmp_path / "data.bin")
    r = subprocess.run([exe, "data.bin", "data.wrb", "data.wrh", "2", "0", "1", "2", "16", "16", "16", "1e-6"], cwd=tmp_path,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_every_environment_switch_is_documented():
    """Every WRB_* variable the library or its Python mirror reads is listed in INTEGRATION.md's table."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for path in glob.glob(os.path.join(root, "waverange_b200", "csrc", "**", "*.*"), recursive=True) + \
            glob.glob(os.path.join(root, "waverange_b200", "*.py")):
        if path.endswith((".cu", ".cpp", ".cuh", ".h", ".py")):
            text = open(path, errors="replace").read()
            names.update(re.findall(r'getenv\("(WRB_[A-Z_0-9]+)"\)', text))
            names.update(re.findall(r'environ(?:\.get)?[\(\[]\s*["\'](WRB_[A-Z_0-9]+)', text))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    assert names, "no switches found: the scan is broken"
    missing = sorted(n for n in names if n not in doc)
    assert not missing, "undocumented environment switches: %s" % missing
