#!/usr/bin/env python3
"""Generate tests/golden/mssg_v1/ with the UNMODIFIED reference MSSG front-ends.

Run in the build container after `python oracle/build_oracle.py` (which compiles oracle/_ref/wrmssgenc_ref and
wrmssgdec_ref from /root/reference/src/mssg with -O2 -ffp-contract=off):

    python tests/golden/make_golden_mssg.py

For every case: the input files (control file + raw data), the command-line parameters (params.txt) and what the
reference's wrmssgenc and wrmssgdec made of them (PREFIX_h*.enc, PREFIX_f*.enc, dec.*).  The cases cover the three
file types (regular GrADS output with and without masked points, merged and divided restart files), both
precisions, both endian settings, the regional (nx, ny) and the global (npg, i_over, j_over) namelist forms, a
constant (trivial) record and the `inmeta` parameter file.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "mssg_v1")
REF = os.path.join(ROOT, "oracle", "_ref")


def smooth(shape, seed):
    """(nz, ny, nx) field"""
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    f = np.zeros(shape)
    for _ in range(6):
        k = rng.integers(1, 4, size=3)
        ph = rng.uniform(0, 6.28, size=3)
        t = np.ones(shape)
        for d in range(3):
            t = t * np.sin(2 * np.pi * k[d] * g[d] + ph[d])
        f += t / np.sqrt((k ** 2).sum())
    return 280.0 + 15.0 * f + 0.01 * rng.standard_normal(shape)


def tofile(path, a, dt, flip):
    a = np.ascontiguousarray(a.astype(dt))
    with open(path, "wb") as f:
        f.write(a.byteswap().tobytes() if flip else a.tobytes())


CTL = """DSET ^{dset}
TITLE MSSG regular output
OPTIONS big_endian
UNDEF {undef}
XDEF {nx} LINEAR 0.0 1.0
YDEF {ny} LINEAR 0.0 1.0
ZDEF {nz} LEVELS 0 1 2 3 4 5 6 7 8 9 10 11
TDEF {nt} LINEAR 00:00Z01JAN2000 1hr
VARS 1
t {nz} 99 temperature
ENDVARS
"""

NMLST_REGIONAL = """&grid
 nx = {nx}, ny = {ny}, nr = {nz}
/
&mpi
 nproc = {nproc}, dim_size = {npx}, {npy}
/
&restart_records
 var = 'time', rec = 1
 var = 'u', rec = 2
 var = 'theta', rec = 3
 var = 'flag', rec = 4
/
"""

# global (Yin-Yang) form: nx = 3 npg - 4 + 2 i_over, ny = 2 (npg + 2 j_over)   (reference ctrl_aux.cpp:156-177)
NMLST_GLOBAL = """&grid
 npg = {npg}, i_over = {io}, j_over = {jo}, nr = {nz}
/
&mpi
 nproc = {nproc}, dim_size = {npx}, {npy}
/
&restart_records
 var = 'time', rec = 1
 var = 'u', rec = 2
 var = 'qv', rec = 3
/
"""


def regular(d, prefix, dt, flip, nx, ny, nz, nt, masked, undef=-999.0):
    open(os.path.join(d, prefix + ".ctl"), "w").write(CTL.format(dset=prefix + ".grd", undef=undef, nx=nx, ny=ny, nz=nz, nt=nt))
    fields = []
    for it in range(nt):
        f = smooth((nz, ny, nx), 10 + it)
        if it in masked:          # "land" below a terrain surface: lower levels of a corner region
            zz, yy, xx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
            f[(zz < 1 + (xx + yy) // 8)] = undef
        fields.append(f)
    tofile(os.path.join(d, prefix + ".grd"), np.stack(fields), dt, flip)


def restart(d, prefix, dt, flip, nx, ny, nz, npx, npy, names, text):
    open(os.path.join(d, prefix + ".nmlst"), "w").write(text)
    nxl, nyl = nx // npx, ny // npy
    glob = []
    for k, name in enumerate(names):
        if name == "time":
            g = np.zeros((nz, ny, nx))
        elif name == "flag":
            g = np.full((nz, ny, nx), 3.0)          # constant record: ntot_enc = 0
        else:
            g = smooth((nz, ny, nx), 30 + k) - 270.0
        glob.append(g)
    for py in range(npy):
        for px in range(npx):
            recs = []
            for k, name in enumerate(names):
                loc = glob[k][:, py * nyl:(py + 1) * nyl, px * nxl:(px + 1) * nxl].copy()
                if name == "time":                  # every process writes the same time record at the start of its file
                    loc.reshape(-1)[:15] = [3600.0 * 7 + 0.125, 0.1, 2000, 1, 1, 7, 0, 0.5, 1e-3, 42, 1.0 / 3.0, 2.5e7, -1, 0, 9]
                recs.append(loc)
            tofile(os.path.join(d, "%s.p_%04d" % (prefix, px + npx * py)), np.stack(recs), dt, flip)


# name, filetype, precision flag (1 single, 2 double), flip, tolerance, procid, use inmeta?, builder
CASES = [
    ("regout_f32_mask", 0, 1, 1, 1e-5, 0, True,
     lambda d: regular(d, "n_tm", "f4", 1, 24, 20, 6, 3, masked={0, 2})),
    ("regout_f64", 0, 2, 0, 1e-7, 0, False,
     lambda d: regular(d, "out", "f8", 0, 20, 18, 10, 2, masked={1})),
    ("united_f64", 1, 2, 1, 1e-6, 0, False,
     lambda d: restart(d, "res", "f8", 1, 32, 32, 8, 2, 2, ["time", "u", "theta", "flag"],
                       NMLST_REGIONAL.format(nx=32, ny=32, nz=8, nproc=4, npx=2, npy=2))),
    ("divided_f32", 2, 1, 0, 1e-4, 1, False,
     lambda d: restart(d, "res", "f4", 0, 32, 32, 8, 2, 2, ["time", "u", "theta", "flag"],
                       NMLST_REGIONAL.format(nx=32, ny=32, nz=8, nproc=4, npx=2, npy=2))),
    ("united_global_f32", 1, 1, 1, 1e-3, 0, False,
     lambda d: restart(d, "yy", "f4", 1, 32, 32, 6, 2, 1, ["time", "u", "qv"],
                       NMLST_GLOBAL.format(npg=10, io=3, jo=3, nz=6, nproc=2, npx=2, npy=1))),
]


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for name, filetype, prec, flip, tol, procid, inmeta, builder in CASES:
        d = os.path.join(OUT, name)
        os.makedirs(d)
        builder(d)
        prefix = [f for f in os.listdir(d) if f.endswith((".ctl", ".nmlst"))][0].rsplit(".", 1)[0]
        open(os.path.join(d, "params.txt"), "w").write("%s .enc %d %d %d %r %d\n" % (prefix, filetype, prec, flip, tol, procid))
        inputs = sorted(os.listdir(d))
        with tempfile.TemporaryDirectory() as tmp:
            for f in inputs:
                shutil.copy(os.path.join(d, f), tmp)
            args = [prefix, ".enc", str(filetype), str(prec), str(flip), repr(tol), str(procid)]
            if inmeta:
                meta = ["# MSSG regular output, see examples/mssg/regout/inmeta", "&prefix_name = %s" % prefix, "&ext_name = .enc",
                        "&file_type = %d" % filetype, "&input_data_type = %d" % prec, "&endian_conversion = %d" % flip,
                        "&tolerance = %r" % tol, "&id_of_proc = %d" % procid]
                open(os.path.join(tmp, "inmeta"), "w").write("\n".join(meta) + "\n")
                shutil.copy(os.path.join(tmp, "inmeta"), d)
                subprocess.run([os.path.join(REF, "wrmssgenc_ref")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
                os.remove(os.path.join(tmp, "inmeta"))
            else:
                subprocess.run([os.path.join(REF, "wrmssgenc_ref")] + args, cwd=tmp, check=True, stdout=subprocess.DEVNULL)
            subprocess.run([os.path.join(REF, "wrmssgdec_ref"), prefix, ".enc", "dec", str(filetype), str(prec), str(flip), str(procid)],
                           cwd=tmp, check=True, stdout=subprocess.DEVNULL)
            for f in sorted(os.listdir(tmp)):
                if f not in inputs:
                    shutil.copy(os.path.join(tmp, f), d)
        print(name, {f: os.path.getsize(os.path.join(d, f)) for f in sorted(os.listdir(d))})


if __name__ == "__main__":
    sys.exit(main())
