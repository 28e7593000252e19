#!/usr/bin/env python3
"""Generate tests/golden/files_v1/ with the UNMODIFIED reference front-ends.

Run in the build container after `python oracle/build_oracle.py` (which compiles
oracle/_ref/wrenc_ref and wrdec_ref from /root/reference/src/generic with -O2 -ffp-contract=off):

    python tests/golden/make_golden_files.py

For every case: the raw input file, the `inmeta` parameter file, and what the reference's wrenc and
wrdec made of them (data.wrh, data.wrb, datarec.bin).  The cases cover the three file types, the
endian flip, the reversed index order, nh > 1, single and double precision, a raw (uncompressed)
field, a constant (trivial) field and several fields per file.
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden", "files_v1")
REF = os.path.join(ROOT, "oracle", "_ref")


def smooth(shape, seed):
    rng = np.random.default_rng(seed)
    g = np.meshgrid(*[np.linspace(0, 1, n) for n in shape], indexing="ij")
    f = np.zeros(shape)
    for _ in range(6):
        k = rng.integers(1, 5, size=len(shape))
        ph = rng.uniform(0, 6.28, size=len(shape))
        t = np.ones(shape)
        for d in range(len(shape)):
            t = t * np.sin(2 * np.pi * k[d] * g[d] + ph[d])
        f += t / np.sqrt((k ** 2).sum())
    return f + 0.01 * rng.standard_normal(shape)


def record(values, filetype, flip):
    """one Fortran/C record in file order"""
    b = values.tobytes()
    if flip:
        b = values.byteswap().tobytes()
    if filetype == 2:
        return b
    m = np.array([len(b)], dtype=np.uint32 if filetype == 0 else np.uint64)
    mb = m.byteswap().tobytes() if flip else m.tobytes()
    return mb + b + mb


# (name, filetype, flip, [fields]); a field = (dtype, (nx, ny, nz, nh), idinv, icomp, tol, kind)
CASES = [
    ("c_f64", 2, 0, [("f8", (24, 20, 18, 1), 0, 1, 1e-6, "smooth")]),
    ("c_f32_multi", 2, 0, [("f4", (32, 16, 16, 1), 0, 1, 1e-4, "smooth"), ("f4", (12, 10, 8, 1), 0, 0, 0.0, "smooth"),
                           ("f8", (16, 16, 16, 1), 0, 1, 1e-9, "const"), ("f8", (10, 9, 8, 3), 0, 1, 1e-3, "smooth")]),
    ("f77_4_flip", 0, 1, [("f4", (20, 18, 16, 1), 0, 1, 1e-5, "smooth"), ("f8", (16, 12, 10, 1), 1, 1, 1e-7, "smooth")]),
    ("f77_8_inv", 1, 0, [("f8", (14, 12, 10, 2), 1, 1, 1e-5, "smooth")]),
]


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    for name, filetype, flip, fields in CASES:
        d = os.path.join(OUT, name)
        os.makedirs(d)
        raw = b""
        meta = ["&in_name = data.bin", "&out_name = data.wrb", "&header_name = data.wrh", "&file_type = %d" % filetype,
                "&endian_conversion = %d" % flip, "&number_of_field = %d" % len(fields)]
        for k, (dt, (nx, ny, nz, nh), idinv, icomp, tol, kind) in enumerate(fields):
            a = smooth((nh, nz, ny, nx), 100 + k) if kind == "smooth" else np.full((nh, nz, ny, nx), 2.5)
            a = a.astype(dt)
            filevals = np.ascontiguousarray(a.transpose(3, 2, 1, 0)) if idinv else a      # reversed order: x slowest
            raw += record(filevals.ravel(), filetype, flip)
            meta += ["%%field = %d" % k, "&input_data_type = %d" % (1 if dt == "f4" else 2), "&nx = %d" % nx, "&ny = %d" % ny,
                     "&nz = %d" % nz, "&nh = %d" % nh, "&order = %d" % idinv, "&compress = %d" % icomp, "&tolerance = %r" % tol, "/"]
        open(os.path.join(d, "data.bin"), "wb").write(raw)
        open(os.path.join(d, "inmeta"), "w").write("\n".join(meta) + "\n")
        with tempfile.TemporaryDirectory() as tmp:
            for f in ("data.bin", "inmeta"):
                shutil.copy(os.path.join(d, f), tmp)
            subprocess.run([os.path.join(REF, "wrenc_ref")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
            os.remove(os.path.join(tmp, "inmeta"))
            subprocess.run([os.path.join(REF, "wrdec_ref"), "data.wrb", "data.wrh", "datarec.bin", str(filetype), str(flip)],
                           cwd=tmp, check=True, stdout=subprocess.DEVNULL)
            for f in ("data.wrh", "data.wrb", "datarec.bin"):
                shutil.copy(os.path.join(tmp, f), d)
        print(name, {f: os.path.getsize(os.path.join(d, f)) for f in sorted(os.listdir(d))})


if __name__ == "__main__":
    sys.exit(main())
