#!/usr/bin/env python3
"""Build the parity oracle (TEST INFRASTRUCTURE ONLY).

1. oracle/libwr_oracle.so      <- oracle/wr_oracle.c, the C restatement
   (gcc -O2 -ffp-contract=off: ISO evaluation, no FMA contraction).
2. oracle/_ref/*.so            <- the UNMODIFIED reference, compiled from the
   sources where they lie under /root/reference/src (nothing copied), only when
   that directory exists (it does not on the GPU box; the prebuilt .so travel):
     libwaverange_ref_strict.so  -O2 -ffp-contract=off        (canonical parity oracle)
     libwaverange_ref_fma.so     the reference's stock flags (config.mk:25,30)
                                 with -march=x86-64-v3 in place of -march=native so
                                 that the binary runs on whatever host the GPU box
                                 has; keeps gcc's default FMA contraction, which is
                                 the property of the stock build that matters
                                 (SURVEY.md section 8c).  Used as the CPU timing baseline.
     wrenc_ref, wrdec_ref        the reference's generic front-ends (strict flags), used by the
                                 .wrh/.wrb file-format parity tests
     wrmssgenc_ref, wrmssgdec_ref  the reference's MSSG front-ends (strict flags), used by the MSSG
                                 file-layout parity tests
     *_dropin                    the reference's front-end sources (gen_enc.cpp, gen_dec.cpp, mssg_enc.cpp,
                                 mssg_dec.cpp + their aux files) compiled unmodified and linked against
                                 waverange_b200/libwaverange_b200.so INSTEAD of the reference's library objects:
                                 the drop-in claim at link level (tests/test_dropin_gpu.py runs them on the GPU)
   The reference's own Makefiles are not run: three C/C++ files are compiled
   directly (src/core/Makefile:6-23 lists the same three objects).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("WR_REFERENCE_DIR", "/root/reference")


def run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_restatement(force=False):
    src = os.path.join(HERE, "wr_oracle.c")
    out = os.path.join(HERE, "libwr_oracle.so")
    if force or newer(out, [src]):
        # -fopenmp only parallelises the *_mt / digest entry points over independent lines, elements and chunks
        run(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", out, src, "-lm"])
    return out


def build_ref(force=False):
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        print("reference sources not present (%s): keeping prebuilt oracle/_ref" % src)
        return None
    outdir = os.path.join(HERE, "_ref")
    os.makedirs(outdir, exist_ok=True)
    wav = os.path.join(src, "waveletcdf97_3d", "waveletcdf97_3d.c")
    rc = os.path.join(src, "rangecod", "rangecod.c")
    wrap = os.path.join(src, "core", "wrappers.cpp")
    shim = os.path.join(HERE, "ref_shim.cpp")
    variants = {
        "strict": (["-O2", "-ffp-contract=off"], ["-O2", "-ffp-contract=off"]),
        "fma": (["-O2", "-ftree-vectorize", "-fomit-frame-pointer", "-funroll-loops", "-march=x86-64-v3"],
                ["-O2", "-ftree-vectorize", "-D__STDC_LIMIT_MACROS", "-march=x86-64-v3"]),
    }
    outs = {}
    for name, (cflags, cxxflags) in variants.items():
        out = os.path.join(outdir, "libwaverange_ref_%s.so" % name)
        outs[name] = out
        if not (force or newer(out, [wav, rc, wrap, shim, __file__])):
            continue
        objs = []
        for s in (wav, rc):
            o = os.path.join(outdir, "%s_%s.o" % (os.path.basename(s)[:-2], name))
            run(["gcc", "-c", "-fPIC", "-w"] + cflags + ["-o", o, s])
            objs.append(o)
        o = os.path.join(outdir, "shim_%s.o" % name)
        run(["g++", "-c", "-fPIC", "-w"] + cxxflags + ["-DWR_REF_WRAPPERS=\"%s\"" % wrap, "-o", o, shim])
        objs.append(o)
        run(["g++", "-shared", "-o", out] + objs)
        for o in objs:
            os.remove(o)
    # the reference's generic front-ends (strict flags), for the file-format parity tests:
    # src/generic/gen_enc.cpp | gen_dec.cpp + gen_aux.cpp + the three library sources (Makefile:22-25)
    gen = os.path.join(src, "generic")
    strict = ["-O2", "-ffp-contract=off", "-w"]
    for exe, main in (("wrenc_ref", "gen_enc.cpp"), ("wrdec_ref", "gen_dec.cpp")):
        out = os.path.join(outdir, exe)
        srcs = [os.path.join(gen, main), os.path.join(gen, "gen_aux.cpp"), wrap]
        if force or newer(out, srcs + [wav, rc, __file__]):
            cobjs = []
            for s in (wav, rc):
                o = os.path.join(outdir, "%s_cli.o" % os.path.basename(s)[:-2])
                run(["gcc", "-c"] + strict + ["-o", o, s])
                cobjs.append(o)
            run(["g++"] + strict + ["-D__STDC_LIMIT_MACROS", "-o", out] + srcs + cobjs)
            for o in cobjs:
                os.remove(o)
        outs[exe] = out
    # the reference's MSSG front-ends: src/mssg/mssg_enc.cpp | mssg_dec.cpp + ctrl_aux.cpp + the library (Makefile:24-28)
    mssg = os.path.join(src, "mssg")
    for exe, main in (("wrmssgenc_ref", "mssg_enc.cpp"), ("wrmssgdec_ref", "mssg_dec.cpp")):
        out = os.path.join(outdir, exe)
        srcs = [os.path.join(mssg, main), os.path.join(mssg, "ctrl_aux.cpp"), wrap]
        if force or newer(out, srcs + [wav, rc, __file__]):
            cobjs = []
            for s in (wav, rc):
                o = os.path.join(outdir, "%s_cli.o" % os.path.basename(s)[:-2])
                run(["gcc", "-c"] + strict + ["-o", o, s])
                cobjs.append(o)
            run(["g++"] + strict + ["-D__STDC_LIMIT_MACROS", "-o", out] + srcs + cobjs)
            for o in cobjs:
                os.remove(o)
        outs[exe] = out
    return outs


def build_dropin(force=False):
    """reference front-ends + the product library (built before this is called; skipped when it is missing)"""
    src = os.path.join(REF, "src")
    lib = os.path.join(os.path.dirname(HERE), "waverange_b200", "libwaverange_b200.so")
    if not os.path.isdir(src) or not os.path.exists(lib):
        return None
    outdir = os.path.join(HERE, "_ref")
    os.makedirs(outdir, exist_ok=True)
    flags = ["-O2", "-ffp-contract=off", "-w", "-D__STDC_LIMIT_MACROS"]
    link = ["-L" + os.path.dirname(lib), "-lwaverange_b200", "-Wl,-rpath,$ORIGIN/../../waverange_b200"]
    jobs = {
        "wrenc_dropin": ["generic/gen_enc.cpp", "generic/gen_aux.cpp"],
        "wrdec_dropin": ["generic/gen_dec.cpp", "generic/gen_aux.cpp"],
        "wrmssgenc_dropin": ["mssg/mssg_enc.cpp", "mssg/ctrl_aux.cpp"],
        "wrmssgdec_dropin": ["mssg/mssg_dec.cpp", "mssg/ctrl_aux.cpp"],
    }
    outs = {}
    for exe, rel in jobs.items():
        out = os.path.join(outdir, exe)
        srcs = [os.path.join(src, r) for r in rel]
        if force or newer(out, srcs + [lib, __file__]):
            run(["g++"] + flags + ["-o", out] + srcs + link)
        outs[exe] = out
    return outs


def main():
    force = "--force" in sys.argv
    build_restatement(force)
    build_ref(force)
    build_dropin(force)


if __name__ == "__main__":
    main()
