#!/usr/bin/env python3
"""Build oracle/_ref from the REFERENCE's own sources (test infrastructure only).

The reference's closure library, src/libbsdl, compiles stand-alone: it needs
nothing but Imath's V2f/V3f/C3f, for which oracle/ref_shim holds a stand-in.  This script

  1. compiles the reference's LUT baker  src/libbsdl/src/genluts.cpp  -> oracle/_ref/genluts
  2. runs it                                                    -> oracle/_ref/include/BSDL/{MTX,SPI}/*_luts.h
  3. compiles oracle/ref_bsdl.cpp (a C-ABI window onto the reference's lobe classes) against the
     reference's headers + those LUTs                          -> oracle/_ref/libref_bsdl.so

Sources are compiled where they lie under /root/reference; only outputs land in oracle/_ref/
(git-ignored; it travels to the GPU box with the snapshot).  The rest of the reference (liboslexec,
testshade, testrender) needs LLVM, OpenImageIO, Imath, pugixml, flex and bison and is not buildable
in this image.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("OSL_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "libref_bsdl.so")


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return r.stdout


def build(force=False):
    bsdl = os.path.join(REF, "src", "libbsdl")
    if not os.path.isdir(bsdl):
        return SO if os.path.exists(SO) else None      # GPU box: use the prebuilt library
    shim = os.path.join(HERE, "ref_shim")
    inc = os.path.join(OUT, "include")
    src = os.path.join(HERE, "ref_bsdl.cpp")
    deps = [src, __file__, os.path.join(shim, "Imath", "ImathVec.h"), os.path.join(shim, "Imath", "ImathColor.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in deps):
        return SO
    for d in ("MTX", "SPI"):
        os.makedirs(os.path.join(inc, "BSDL", d), exist_ok=True)
    genluts = os.path.join(OUT, "genluts")
    luts_done = os.path.join(inc, "BSDL", "MTX", "bsdf_dielectric_bothback_luts.h")
    if force or not os.path.exists(luts_done):
        run(["g++", "-std=c++17", "-O2", "-I", shim, "-I", os.path.join(bsdl, "include"),
             os.path.join(bsdl, "src", "genluts.cpp"), "-lpthread", "-o", genluts])
        run([genluts, os.path.join(inc, "BSDL")])
    run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", HERE, "-I", shim, "-I", inc,
         "-I", os.path.join(bsdl, "include"), src, "-o", SO + ".tmp"])
    os.replace(SO + ".tmp", SO)
    return SO


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
