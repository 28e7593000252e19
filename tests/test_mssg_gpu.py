"""GPU parity tests of the MSSG file layer and the wrmssgenc/wrmssgdec front-ends (NEXT-3):
against files written by the reference's own wrmssgenc/wrmssgdec (tests/golden/mssg_v1) and, where oracle/_ref
travelled, against the reference binaries run on the spot."""
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "mssg_v1")
CASES = sorted(os.path.basename(d) for d in glob.glob(os.path.join(GOLD, "*")))
BIN = os.path.join(ROOT, "waverange_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def params(name):
    p = open(os.path.join(GOLD, name, "params.txt")).read().split()
    return dict(prefix=p[0], ext=p[1], filetype=p[2], prec=p[3], flip=p[4], tol=p[5], procid=p[6])


def inputs_of(name):
    p = params(name)
    return [f for f in sorted(os.listdir(os.path.join(GOLD, name)))
            if f.startswith(p["prefix"] + ".") or f == "inmeta"]


def encoded_of(name):
    p = params(name)
    return [f for f in sorted(os.listdir(os.path.join(GOLD, name))) if f.startswith(p["prefix"] + "_")]


def decoded_of(name):
    return [f for f in sorted(os.listdir(os.path.join(GOLD, name))) if f.startswith("dec.")]


def run(exe, args, cwd, env=None):
    r = subprocess.run([exe] + args, cwd=cwd, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def enc_args(p):
    return [p["prefix"], p["ext"], p["filetype"], p["prec"], p["flip"], p["tol"], p["procid"]]


def dec_args(p, out="dec"):
    return [p["prefix"], p["ext"], out, p["filetype"], p["prec"], p["flip"], p["procid"]]


@pytest.mark.parametrize("name", CASES)
def test_stock_layout_files_are_byte_identical_to_the_reference(product_lib, torch_cuda, tmp_path, name):
    """WRB_CHUNK_BLOCKS=0: header and encoded file equal what the reference's wrmssgenc wrote (mask records, padded
    fields, time record, trivial records), and decoding the reference's files gives the bytes its wrmssgdec wrote"""
    g, p = os.path.join(GOLD, name), params(name)
    env = dict(os.environ, WRB_CHUNK_BLOCKS="0")
    for f in inputs_of(name):
        if f != "inmeta":
            shutil.copy(os.path.join(g, f), tmp_path)
    run(os.path.join(BIN, "wrmssgenc"), enc_args(p), tmp_path, env)
    for f in encoded_of(name):
        assert open(tmp_path / f, "rb").read() == open(os.path.join(g, f), "rb").read(), f
    # decode the REFERENCE's encoded files
    d2 = tmp_path / "fromref"
    d2.mkdir()
    for f in inputs_of(name) + encoded_of(name):
        if f != "inmeta" and not f.endswith((".grd",)) and ".p_" not in f:
            shutil.copy(os.path.join(g, f), d2)
    run(os.path.join(BIN, "wrmssgdec"), dec_args(p), d2, env)
    for f in decoded_of(name):
        assert open(d2 / f, "rb").read() == open(os.path.join(g, f), "rb").read(), f


@pytest.mark.parametrize("name", CASES)
def test_chunked_files_round_trip(product_lib, torch_cuda, tmp_path, name):
    """default layout (WRCK chunk containers): same header doubles as the reference, reconstruction identical to the
    reference's (the symbols are the same, only their packaging differs) -- through the library calls"""
    from waverange_b200 import api
    g, p = os.path.join(GOLD, name), params(name)
    for f in inputs_of(name):
        if f != "inmeta":
            shutil.copy(os.path.join(g, f), tmp_path)
    cwd = os.getcwd()
    os.chdir(tmp_path)                      # DSET and the .p_ files are relative to the working directory, as in the reference
    try:
        c = api.Codec(device=0)
        ft, nb = int(p["filetype"]), 4 if p["prec"] == "1" else 8
        c.mssg_encode(p["prefix"], p["ext"], ft, nb, int(p["flip"]), float(p["tol"]), int(p["procid"]))
        hname = [f for f in encoded_of(name) if "_h" in f][0]
        t_got, got = api.mssg_header_read(hname, ft)
        t_want, want = api.mssg_header_read(os.path.join(g, hname), ft)
        assert t_got == t_want and len(got) == len(want)
        for (ia, na, a), (ib, nb_, b) in zip(got, want):
            assert (ia, na) == (ib, nb_)
            assert (a.tolabs, a.midval, a.halfspanval, a.wlev, a.nlay) == (b.tolabs, b.midval, b.halfspanval, b.wlev, b.nlay)
            n = a.nlay
            assert list(a.deps_vec)[:n] == list(b.deps_vec)[:n] and list(a.minval_vec)[:n] == list(b.minval_vec)[:n]
        c.mssg_decode(p["prefix"], p["ext"], "dec", ft, nb, int(p["flip"]), int(p["procid"]))
        c.close()
    finally:
        os.chdir(cwd)
    for f in decoded_of(name):
        assert open(tmp_path / f, "rb").read() == open(os.path.join(g, f), "rb").read(), f


def test_cli_inmeta(product_lib, torch_cuda, tmp_path):
    """the `inmeta` parameter file (reference mssg_enc.cpp:104-189) drives wrmssgenc as it drives the reference, in the
    namelist-like and in the old one-value-per-line format; wrmssgdec takes its answers from standard input"""
    name = "regout_f32_mask"
    g, p = os.path.join(GOLD, name), params(name)
    env = dict(os.environ, WRB_CHUNK_BLOCKS="0")
    for variant in ("new", "old"):
        d = tmp_path / variant
        d.mkdir()
        for f in inputs_of(name):
            shutil.copy(os.path.join(g, f), d)
        if variant == "old":
            (d / "inmeta").write_text("\n".join(enc_args(p)) + "\n")
        run(os.path.join(BIN, "wrmssgenc"), [], d, env)
        for f in encoded_of(name):
            assert open(d / f, "rb").read() == open(os.path.join(g, f), "rb").read(), (variant, f)
        os.remove(d / "inmeta")
        r = subprocess.run([os.path.join(BIN, "wrmssgdec")], cwd=d, env=env, input="\n".join(dec_args(p)) + "\n",
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        for f in decoded_of(name):
            assert open(d / f, "rb").read() == open(os.path.join(g, f), "rb").read(), (variant, f)


def test_masked_field_semantics(product_lib, torch_cuda, tmp_path):
    """masked points come back as UNDEF exactly, all others within the tolerance of the padded field
    (reference mssg_enc.cpp:305-365, mssg_dec.cpp:236-310)"""
    name = "regout_f32_mask"
    g, p = os.path.join(GOLD, name), params(name)
    for f in inputs_of(name):
        if f != "inmeta":
            shutil.copy(os.path.join(g, f), tmp_path)
    run(os.path.join(BIN, "wrmssgenc"), enc_args(p), tmp_path)
    run(os.path.join(BIN, "wrmssgdec"), dec_args(p), tmp_path)
    a = np.fromfile(tmp_path / "n_tm.grd", dtype=">f4").astype(np.float64)
    b = np.fromfile(tmp_path / "dec.grd", dtype=">f4").astype(np.float64)
    masked = a < -998.0
    assert masked.any() and (b[masked] == -999.0).all() and not (b[~masked] < 0).any()
    assert np.abs(a[~masked] - b[~masked]).max() <= float(p["tol"]) * np.abs(a[~masked]).max()
    assert os.path.getsize(tmp_path / "n_tm_f.enc") < 0.7 * os.path.getsize(tmp_path / "n_tm.grd")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "wrmssgenc_ref")), reason="oracle/_ref did not travel")
@pytest.mark.parametrize("filetype,prec,flip", [(0, 2, 1), (1, 1, 0), (2, 2, 1)])
def test_against_reference_binaries_on_fresh_data(product_lib, torch_cuda, tmp_path, filetype, prec, flip):
    """a larger data set than the committed fixtures (64 x 48 x 20): stock-layout files byte-identical to the
    reference's, the reference's wrmssgdec reads them, and both decoders write the same bytes"""
    rng = np.random.default_rng(filetype * 10 + prec)
    nx, ny, nz = 64, 48, 20
    dt = (">" if flip else "<") + ("f4" if prec == 1 else "f8")
    x, y, z = np.meshgrid(np.linspace(0, 1, nx), np.linspace(0, 1, ny), np.linspace(0, 1, nz), indexing="ij")

    def fld(k):
        f = np.sin(2 * np.pi * (k + 1) * x + k) * np.cos(2 * np.pi * y * (k + 2)) * np.exp(-z * (1 + k)) + 0.001 * rng.standard_normal(x.shape)
        return np.ascontiguousarray(f.transpose(2, 1, 0))           # (nz, ny, nx)

    ours, refd = tmp_path / "ours", tmp_path / "ref"
    ours.mkdir(), refd.mkdir()
    if filetype == 0:
        prefix, nt = "reg", 2
        fields = [fld(0), fld(1)]
        fields[1][:3, :10, :] = -9999.0
        np.stack(fields).astype(dt).tofile(ours / "reg.grd")
        (ours / "reg.ctl").write_text("DSET ^reg.grd\nUNDEF -9999.0\nXDEF %d LINEAR 0 1\nYDEF %d LINEAR 0 1\nZDEF %d LINEAR 0 1\nTDEF %d LINEAR 0 1\n"
                                      % (nx, ny, nz, nt))
    else:
        prefix, npx, npy = "res", 2, 3
        (ours / "res.nmlst").write_text("&a nx = %d, ny = %d, nr = %d /\n&b nproc = %d, dim_size = %d, %d /\n&c var = 'time', rec = 1\n"
                                        " var = 'u', rec = 2\n var = 'w', rec = 3\n/\n" % (nx, ny, nz, npx * npy, npx, npy))
        glob_ = [np.zeros((nz, ny, nx)), fld(2), fld(3)]
        nxl, nyl = nx // npx, ny // npy
        for py in range(npy):
            for px in range(npx):
                recs = [a[:, py * nyl:(py + 1) * nyl, px * nxl:(px + 1) * nxl].copy() for a in glob_]
                recs[0].reshape(-1)[:15] = np.arange(15) * 1.5 + 0.1
                np.stack(recs).astype(dt).tofile(ours / ("res.p_%04d" % (px + npx * py)))
    for f in os.listdir(ours):
        shutil.copy(ours / f, refd)
    args = [prefix, ".enc", str(filetype), str(prec), str(flip), "1e-5", "3" if filetype == 2 else "0"]
    dargs = [prefix, ".enc", "dec", str(filetype), str(prec), str(flip), args[-1]]
    env = dict(os.environ, WRB_CHUNK_BLOCKS="0")
    run(os.path.join(BIN, "wrmssgenc"), args, ours, env)
    subprocess.run([os.path.join(REF, "wrmssgenc_ref")] + args, cwd=refd, check=True, stdout=subprocess.DEVNULL)
    made = sorted(f for f in os.listdir(refd) if f.endswith(".enc"))
    assert len(made) == 2
    for f in made:
        assert open(ours / f, "rb").read() == open(refd / f, "rb").read(), f
    run(os.path.join(BIN, "wrmssgdec"), dargs, ours, env)
    subprocess.run([os.path.join(REF, "wrmssgdec_ref")] + dargs, cwd=refd, check=True, stdout=subprocess.DEVNULL)
    outs = sorted(f for f in os.listdir(refd) if f.startswith("dec."))
    assert outs
    for f in outs:
        assert open(ours / f, "rb").read() == open(refd / f, "rb").read(), f
