"""Build libwaverange_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m waverange_b200.build [--force]

Flags that matter for parity: -fmad=false (device) and -ffp-contract=off (host) keep every + and *
individually rounded, like the canonical (-ffp-contract=off) build of the reference.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwaverange_b200.so")
SOURCES = ["codec.cu", "wavelet.cu", "wavelet_fused.cu", "wavelet_inv_fused.cu", "wavelet_inv2.cu", "wavelet_slab.cu", "slab_comm.cu", "slab_order.cu", "quant.cu", "rangecoder.cu", "compat.cpp", "wrfile.cpp", "mssgfile.cpp"]
HEADERS = ["wr_common.cuh", "wr_kernels.h", "slab_comm.h", "wavelet_pairs.cuh", "../../include/waverange_b200.h", "../../include/waverange.h",
           "../../include/waverange_files.h", "../../include/waverange_mssg.h", "cli/wrenc.cpp", "cli/wrdec.cpp",
           "cli/wrmssgenc.cpp", "cli/wrmssgdec.cpp"]
BIN = os.path.join(HERE, "bin")
CLI = ["wrenc", "wrdec", "wrmssgenc", "wrmssgdec"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default",
    "-Xptxas", "-v",
]


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [__file__]
    return any(os.path.getmtime(os.path.normpath(d)) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc()] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            failed = True
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed or verbose:
        print("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed, see waverange_b200/csrc/build.log")
    cmd = [nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                   "-Xlinker", "-Bsymbolic", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    for o in objs:
        os.remove(o)
    # the generic and MSSG front-ends (reference bin/generic/wrenc, wrdec, bin/mssg/wrmssgenc, wrmssgdec): thin mains over
    # wrb_file_encode / wrb_file_decode / wrb_mssg_encode / wrb_mssg_decode
    os.makedirs(BIN, exist_ok=True)
    for name in CLI:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", os.path.join(BIN, name), os.path.join(CSRC, "cli", name + ".cpp"),
                               "-L" + HERE, "-lwaverange_b200", "-Wl,-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
