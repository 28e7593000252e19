"""GPU parity tests of the generic file layer and the wrenc/wrdec front-ends (NEXT-1;
BASELINE.json configs[0]): against files written by the reference's own wrenc/wrdec
(tests/golden/files_v1) and, where oracle/_ref travelled, against the reference binaries run on the spot."""
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = sorted(os.path.basename(d) for d in glob.glob(os.path.join(ROOT, "tests", "golden", "files_v1", "*")))
BIN = os.path.join(ROOT, "waverange_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def case_dir(name):
    return os.path.join(ROOT, "tests", "golden", "files_v1", name)


def header_params(path):
    lines = open(path).read().split("\n")
    return int(lines[3].rsplit(":", 1)[1]), (0 if lines[4].strip() == "No endian conversion" else 1)


def last_tolerance(g):
    return [float(l.split("=")[1]) for l in open(os.path.join(g, "inmeta")) if l.startswith("&tolerance")][-1]


def fields_of(recs):
    from waverange_b200 import api
    return [api.FieldDesc(r.desc.nbytes, r.desc.nx, r.desc.ny, r.desc.nz, r.desc.nh, r.desc.idinv, r.desc.icomp, r.desc.tol_base)
            for r in recs]


@pytest.mark.parametrize("name", CASES)
def test_stock_layout_files_are_byte_identical_to_the_reference(product_lib, torch_cuda, tmp_path, name):
    """chunk_blocks = 0: data.wrh and data.wrb equal what the reference's wrenc wrote; decoding the reference's
    files gives the bytes the reference's wrdec wrote"""
    from waverange_b200 import api
    g = case_dir(name)
    filetype, flip = header_params(os.path.join(g, "data.wrh"))
    recs = api.wrh_read(os.path.join(g, "data.wrh"))
    c = api.Codec(device=0, chunk_blocks=0)
    wrb, wrh, rec = str(tmp_path / "data.wrb"), str(tmp_path / "data.wrh"), str(tmp_path / "datarec.bin")
    # the reference's wrenc codes every field with the tolerance it parsed last (gen_enc.cpp:497-500)
    c.file_encode(os.path.join(g, "data.bin"), wrb, wrh, filetype, flip, fields_of(recs), cutoff_all=last_tolerance(g))
    assert open(wrh, "rb").read().replace(wrb.encode(), b"data.wrb") == open(os.path.join(g, "data.wrh"), "rb").read()
    assert open(wrb, "rb").read() == open(os.path.join(g, "data.wrb"), "rb").read()
    c.file_decode(os.path.join(g, "data.wrb"), os.path.join(g, "data.wrh"), rec, filetype, flip)
    assert open(rec, "rb").read() == open(os.path.join(g, "datarec.bin"), "rb").read()
    c.close()


@pytest.mark.parametrize("name", CASES)
def test_chunked_files_round_trip(product_lib, torch_cuda, tmp_path, name):
    """default layout (WRCK chunk containers): same header doubles as the reference, reconstruction identical to
    the reference's (the symbols are the same, only their packaging differs)"""
    from waverange_b200 import api
    g = case_dir(name)
    filetype, flip = header_params(os.path.join(g, "data.wrh"))
    want = api.wrh_read(os.path.join(g, "data.wrh"))
    c = api.Codec(device=0)
    wrb, wrh, rec = str(tmp_path / "data.wrb"), str(tmp_path / "data.wrh"), str(tmp_path / "datarec.bin")
    c.file_encode(os.path.join(g, "data.bin"), wrb, wrh, filetype, flip, fields_of(want), cutoff_all=last_tolerance(g))
    got = api.wrh_read(wrh)
    for a, b in zip(got, want):
        if a.desc.icomp:
            assert (a.hdr.tolabs, a.hdr.midval, a.hdr.halfspanval, a.hdr.wlev, a.hdr.nlay) == \
                   (b.hdr.tolabs, b.hdr.midval, b.hdr.halfspanval, b.hdr.wlev, b.hdr.nlay)
            n = a.hdr.nlay
            assert list(a.hdr.deps_vec)[:n] == list(b.hdr.deps_vec)[:n] and list(a.hdr.minval_vec)[:n] == list(b.hdr.minval_vec)[:n]
    c.file_decode(wrb, wrh, rec, filetype, flip)
    assert open(rec, "rb").read() == open(os.path.join(g, "datarec.bin"), "rb").read()
    c.close()


def synth(n, seed=5):
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 1, n)
    f = np.zeros((n, n, n))
    for _ in range(12):
        k = rng.integers(1, 9, 3)
        p = rng.uniform(0, 6.28, 3)
        f += (np.sin(2 * np.pi * k[0] * x + p[0])[:, None, None] * np.sin(2 * np.pi * k[1] * x + p[1])[None, :, None]
              * np.sin(2 * np.pi * k[2] * x + p[2])[None, None, :]) / np.sqrt((k ** 2).sum())
    return f.astype(np.float32)


def test_cli_round_trip_config0(product_lib, torch_cuda, tmp_path):
    """BASELINE.json configs[0] through the executables: wrenc/wrdec round trip of a float32 C-layout field at
    tolerance 1e-5 at the config's own size, 256^3, reference command lines unchanged; the same job through the
    reference's executables gives the same reconstruction, and in stock layout the same files"""
    n, tol = 256, 1e-5
    f = synth(n)
    f.tofile(tmp_path / "data.bin")
    env = dict(os.environ)
    r = subprocess.run([os.path.join(BIN, "wrenc"), "data.bin", "data.wrb", "data.wrh", "2", "0", "1", "1", str(n), str(n), str(n), str(tol)],
                       cwd=tmp_path, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([os.path.join(BIN, "wrdec"), "data.wrb", "data.wrh", "datarec.bin", "2", "0"], cwd=tmp_path, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    g = np.fromfile(tmp_path / "datarec.bin", dtype=np.float32).reshape(f.shape)
    err = np.abs(g.astype(np.float64) - f).max() / np.abs(f).max()
    assert err <= tol                                   # relative L-infinity tolerance in the reference's sense
    ratio = f.nbytes / os.path.getsize(tmp_path / "data.wrb")
    assert ratio > 2.0
    if os.path.exists(os.path.join(REF, "wrenc_ref")):   # same job through the reference's executables
        ref_dir = tmp_path / "ref"
        ref_dir.mkdir()
        shutil.copy(tmp_path / "data.bin", ref_dir)
        subprocess.run([os.path.join(REF, "wrenc_ref"), "data.bin", "data.wrb", "data.wrh", "2", "0", "1", "1", str(n), str(n), str(n), str(tol)],
                       cwd=ref_dir, check=True, stdout=subprocess.DEVNULL)
        subprocess.run([os.path.join(REF, "wrdec_ref"), "data.wrb", "data.wrh", "datarec.bin", "2", "0"], cwd=ref_dir, check=True,
                       stdout=subprocess.DEVNULL)
        assert open(ref_dir / "datarec.bin", "rb").read() == open(tmp_path / "datarec.bin", "rb").read()
        ref_size = os.path.getsize(ref_dir / "data.wrb")
        assert os.path.getsize(tmp_path / "data.wrb") <= 1.01 * ref_size      # ratio within 1 % of the reference
        # stock layout: our files are the reference's files, and the reference's wrdec reads them
        env0 = dict(env, WRB_CHUNK_BLOCKS="0")
        st = tmp_path / "stock"
        st.mkdir()
        shutil.copy(tmp_path / "data.bin", st)
        subprocess.run([os.path.join(BIN, "wrenc"), "data.bin", "data.wrb", "data.wrh", "2", "0", "1", "1", str(n), str(n), str(n), str(tol)],
                       cwd=st, env=env0, check=True, stdout=subprocess.DEVNULL)
        assert open(st / "data.wrb", "rb").read() == open(ref_dir / "data.wrb", "rb").read()
        assert open(st / "data.wrh", "rb").read() == open(ref_dir / "data.wrh", "rb").read()
        subprocess.run([os.path.join(REF, "wrdec_ref"), "data.wrb", "data.wrh", "datarec.bin", "2", "0"], cwd=st, check=True,
                       stdout=subprocess.DEVNULL)
        assert open(st / "datarec.bin", "rb").read() == open(ref_dir / "datarec.bin", "rb").read()


def test_cli_inmeta(product_lib, torch_cuda, tmp_path):
    """the `inmeta` parameter file (reference gen_enc.cpp:111-350) drives wrenc exactly as it drives the reference"""
    g = case_dir("c_f32_multi")
    for f in ("data.bin", "inmeta"):
        shutil.copy(os.path.join(g, f), tmp_path)
    env = dict(os.environ, WRB_CHUNK_BLOCKS="0")
    r = subprocess.run([os.path.join(BIN, "wrenc")], cwd=tmp_path, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(tmp_path / "data.wrh", "rb").read() == open(os.path.join(g, "data.wrh"), "rb").read()
    assert open(tmp_path / "data.wrb", "rb").read() == open(os.path.join(g, "data.wrb"), "rb").read()
