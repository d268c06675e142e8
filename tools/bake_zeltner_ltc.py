#!/usr/bin/env python
"""Write openshadinglanguage_b200/data/zeltner_ltc.bin: the 32 x 32 x (A, B, R) table of fitted
linearly-transformed-cosine coefficients of the Zeltner-Burley sheen (Zeltner, Burley, Chiang,
"Practical Multiple-Scattering Sheen Using Linearly Transformed Cosines", SIGGRAPH 2022; the
published table of github.com/tizian/ltc-sheen as carried by libbsdl in
BSDL/MTX/bsdf_zeltnersheen_param.h).  It is fitted data, not an algorithm, so unlike the energy
tables (tools/bake_bsdl_luts.cpp) it cannot be regenerated: this tool reads it out of the
reference's own libbsdl build (oracle/_ref, needs /root/reference) once and stores the 3072
floats, the way the texture fixtures are stored.  Rows = roughness, columns = cos(theta_o)."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402


def main():
    so = build_ref.build(force=True)
    if not so:
        raise SystemExit("oracle/_ref is not available (no /root/reference): nothing baked")
    lib = ctypes.CDLL(so)
    lib.ref_bsdl_lut.restype = ctypes.POINTER(ctypes.c_float)
    n = ctypes.c_int(0)
    p = lib.ref_bsdl_lut(7, ctypes.byref(n))
    assert n.value == 32 * 32 * 3
    a = np.ctypeslib.as_array(p, shape=(n.value,)).astype(np.float32).copy()
    out = os.path.join(ROOT, "openshadinglanguage_b200", "data", "zeltner_ltc.bin")
    a.tofile(out)
    print("wrote", out, a.shape, float(a.min()), float(a.max()))


if __name__ == "__main__":
    main()
