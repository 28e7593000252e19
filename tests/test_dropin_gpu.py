"""The drop-in claim end to end: the reference's OWN front-end sources (src/generic/gen_enc.cpp, gen_dec.cpp,
src/mssg/mssg_enc.cpp, mssg_dec.cpp and their aux files), compiled unmodified and linked against
libwaverange_b200.so instead of the reference's library (oracle/build_oracle.py: oracle/_ref/*_dropin), run on the
GPU and reproduce the files the all-reference executables wrote (tests/golden/files_v1, mssg_v1).  Nothing but the
six entry points of include/waverange.h connects the two sides."""
import glob
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
GEN = os.path.join(ROOT, "tests", "golden", "files_v1")
MSSG = os.path.join(ROOT, "tests", "golden", "mssg_v1")
GEN_CASES = sorted(os.path.basename(d) for d in glob.glob(os.path.join(GEN, "*")))
MSSG_CASES = sorted(os.path.basename(d) for d in glob.glob(os.path.join(MSSG, "*")))

needs_dropin = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "wrenc_dropin")),
                                  reason="oracle/_ref/*_dropin not built (needs /root/reference at build time)")


def run(exe, args, cwd, env):
    r = subprocess.run([os.path.join(REF, exe)] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (exe, r.stdout[-1500:], r.stderr[-1500:])
    return r


def same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def header_params(path):
    lines = open(path).read().split("\n")
    return lines[3].rsplit(":", 1)[1].strip(), ("0" if lines[4].strip() == "No endian conversion" else "1")


@needs_dropin
@pytest.mark.parametrize("layout", ["stock", "chunked"])
@pytest.mark.parametrize("name", GEN_CASES)
def test_reference_generic_front_end_on_our_library(product_lib, torch_cuda, tmp_path, name, layout):
    g = os.path.join(GEN, name)
    env = dict(os.environ)
    if layout == "stock":
        env["WRB_CHUNK_BLOCKS"] = "0"
    else:
        env.pop("WRB_CHUNK_BLOCKS", None)
    for f in ("data.bin", "inmeta"):
        shutil.copy(os.path.join(g, f), tmp_path)
    run("wrenc_dropin", [], tmp_path, env)                    # the reference's wrenc reads `inmeta` from the working directory
    os.remove(tmp_path / "inmeta")
    if layout == "stock":                                     # one stream per layer: the reference's files, byte for byte
        assert same(tmp_path / "data.wrh", os.path.join(g, "data.wrh"))
        assert same(tmp_path / "data.wrb", os.path.join(g, "data.wrb"))
    filetype, flip = header_params(os.path.join(g, "data.wrh"))
    run("wrdec_dropin", ["data.wrb", "data.wrh", "datarec.bin", filetype, flip], tmp_path, env)
    assert same(tmp_path / "datarec.bin", os.path.join(g, "datarec.bin"))     # same symbols either way


def mssg_params(name):
    p = open(os.path.join(MSSG, name, "params.txt")).read().split()
    return dict(prefix=p[0], ext=p[1], filetype=p[2], prec=p[3], flip=p[4], tol=p[5], procid=p[6])


@needs_dropin
@pytest.mark.parametrize("name", MSSG_CASES)
def test_reference_mssg_front_end_on_our_library(product_lib, torch_cuda, tmp_path, name):
    g, p = os.path.join(MSSG, name), mssg_params(name)
    env = dict(os.environ, WRB_CHUNK_BLOCKS="0")
    files = sorted(os.listdir(g))
    inputs = [f for f in files if f.startswith(p["prefix"] + ".")]
    encoded = [f for f in files if f.startswith(p["prefix"] + "_")]
    decoded = [f for f in files if f.startswith("dec.")]
    for f in inputs:
        shutil.copy(os.path.join(g, f), tmp_path)
    run("wrmssgenc_dropin", [p["prefix"], p["ext"], p["filetype"], p["prec"], p["flip"], p["tol"], p["procid"]], tmp_path, env)
    assert encoded
    for f in encoded:
        assert same(tmp_path / f, os.path.join(g, f)), f
    run("wrmssgdec_dropin", [p["prefix"], p["ext"], "dec", p["filetype"], p["prec"], p["flip"], p["procid"]], tmp_path, env)
    assert decoded
    for f in decoded:
        assert same(tmp_path / f, os.path.join(g, f)), f
