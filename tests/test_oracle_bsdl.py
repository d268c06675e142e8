"""The restated libbsdl lobes (oracle/osl_oracle_lobes.h, osl_oracle_mxlobes.h) against the
REFERENCE'S OWN classes, compiled from /root/reference/src/libbsdl where they lie into
oracle/_ref/libref_bsdl.so (oracle/build_ref.py; testrender's configuration: BSDL_WRAP globals,
three RGB channels, OIIO fast_* math):

  * eval / sample / albedo / filter_o of mtx::ConductorLobe, DielectricLobe, SchlickLobe,
    TranslucentLobe, SheenLobe (Conty-Kulla), OrenNayarDiffuseLobe, BurleyDiffuseLobe on random
    parameters, directions, sides and path roughness: BIT-EXACT;
  * the energy tables shipped in openshadinglanguage_b200/data/bsdl_luts.bin (baked by
    tools/bake_bsdl_luts.cpp, a restatement of the reference's genluts) equal the tables the
    reference's genluts produces, entry by entry.

oracle/_ref is built here when /root/reference exists and travels to the GPU box prebuilt.
"""
import ctypes
import os

import numpy as np
import pytest

import helpers
from oracle import build_ref, oracle

SO = build_ref.build()
pytestmark = pytest.mark.skipif(SO is None, reason="oracle/_ref not built and /root/reference absent")

_ARGS = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_void_p,
         ctypes.c_void_p]


@pytest.fixture(scope="module")
def libs():
    ref = ctypes.CDLL(SO)
    orc = ctypes.CDLL(oracle.build_bsdl_check())
    ref.ref_bsdl.argtypes = _ARGS
    orc.oracle_bsdl.argtypes = _ARGS
    ref.ref_bsdl_lut.restype = ctypes.POINTER(ctypes.c_float)
    ref.ref_bsdl_lut.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    luts = oracle.bsdl_luts()
    orc.oracle_set_bsdl_luts.argtypes = [ctypes.c_void_p]
    orc.oracle_set_bsdl_luts(luts.ctypes.data)
    return ref, orc, luts


def test_baked_energy_tables_equal_the_reference_genluts_output(libs):
    ref, _, luts = libs
    tabs = []
    # MiniMicrofacetGGX, DielectricReflFront, BothFront, BothBack (baked by tools/bake_bsdl_luts.cpp)
    # + the Zeltner-Burley sheen LTC coefficients (published fit, stored by tools/bake_zeltner_ltc.py)
    # + spi::Thinlayer (same baker, openshadinglanguage_b200/data/thinlayer_lut.bin)
    for t in (0, 1, 2, 3, 7, 6):
        n = ctypes.c_int()
        p = ref.ref_bsdl_lut(t, ctypes.byref(n))
        tabs.append(np.ctypeslib.as_array(p, (n.value,)).copy())
    want = np.concatenate(tabs).astype(np.float32)
    assert luts.size == want.size == 256 + 3 * 8192 + 32 * 32 * 3 + 8192
    assert np.array_equal(luts.view(np.uint32), want.view(np.uint32))
    energy = np.concatenate([luts[:256 + 3 * 8192], luts[-8192:]])
    assert 0.0 <= energy.min() and energy.max() <= 1.0


def _unit(v):
    return v / np.linalg.norm(v)


def _params(rng, lobe):
    N, U = _unit(rng.normal(size=3)), _unit(rng.normal(size=3))
    c = lambda: rng.uniform(0, 1, 3)
    r2 = [rng.uniform(0, 0.8), rng.uniform(0, 0.8)]
    if lobe == 0:       # conductor: N U rx ry ior extinction
        p = np.concatenate([N, U, r2, rng.uniform(0.1, 3, 3), rng.uniform(0.5, 5, 3)])
    elif lobe == 1:     # dielectric: N U refl refr rx ry ior | thinfilm(2) absorption dispersion
        refr = c() if rng.random() < 0.6 else np.zeros(3)
        ab = c() * 2 if rng.random() < 0.5 else np.zeros(3)
        p = np.concatenate([N, U, c(), refr, r2, [rng.uniform(1.0, 3.0)], [0, 0], ab, [0]])
    elif lobe == 2:     # generalized schlick: N U refl refr rx ry F0 F90 exponent
        refr = c() if rng.random() < 0.5 else np.zeros(3)
        p = np.concatenate([N, U, c(), refr, r2, c() * 0.5, c(), [rng.uniform(1, 8)]])
    elif lobe == 3:     # translucent: N albedo
        p = np.concatenate([N, c()])
    elif lobe == 4:     # sheen: N albedo roughness mode (0 Conty-Kulla, 1 Zeltner-Burley LTC, other -> Conty-Kulla)
        p = np.concatenate([N, c(), [rng.uniform(0, 1), float(rng.choice([0, 1, 1, 2]))]])
    elif lobe == 5:     # oren-nayar diffuse: N albedo roughness energy_compensation
        p = np.concatenate([N, c(), [rng.uniform(0, 1), float(rng.integers(0, 2))]])
    elif lobe == 7:     # spi thinlayer: N T IOR roughness anisotropy thickness refl_tint refr_tint sigma_t
        sig = c() * rng.choice([0.0, 0.3, 3.0]) if rng.random() < 0.7 else np.zeros(3)
        p = np.concatenate([N, U, [rng.uniform(1.0, 3.0), rng.uniform(0, 1), rng.uniform(0, 0.95),
                                   rng.choice([0.0, 1.0, rng.uniform(0, 4)])], c(), c(), sig])
    else:               # burley diffuse: N albedo roughness
        p = np.concatenate([N, c(), [rng.uniform(0, 1)]])
    return N, p.astype(np.float32)


@pytest.mark.parametrize("lobe,name", [(0, "conductor"), (1, "dielectric"), (2, "generalized_schlick"),
                                       (3, "translucent"), (4, "sheen"), (5, "oren_nayar_diffuse"),
                                       (6, "burley_diffuse"), (7, "spi_thinlayer")])
def test_restated_lobe_is_bit_exact_against_the_reference_class(libs, lobe, name):
    ref, orc, _ = libs
    rng = np.random.default_rng(1000 + lobe)
    checked = 0
    for _ in range(1500):
        N, p = _params(rng, lobe)
        while True:                       # wo on the visible side (the integrator's -I)
            wo = _unit(rng.normal(size=3))
            if wo @ N > 0.02:
                break
        wo = wo.astype(np.float32)
        backfacing = int(rng.random() < 0.3)
        path_roughness = np.float32(rng.uniform(0, 0.5) if rng.random() < 0.5 else 0)
        for mode in (0, 1, 2, 3):         # eval, sample, albedo, filter_o
            arg = (_unit(rng.normal(size=3)) if mode == 0 else rng.uniform(0, 1, 3)).astype(np.float32)
            a, b = np.zeros(8, np.float32), np.zeros(8, np.float32)
            ra = ref.ref_bsdl(lobe, p.ctypes.data, wo.ctypes.data, backfacing, path_roughness, mode, arg.ctypes.data,
                              a.ctypes.data)
            rb = orc.oracle_bsdl(lobe, p.ctypes.data, wo.ctypes.data, backfacing, path_roughness, mode,
                                 arg.ctypes.data, b.ctypes.data)
            assert (ra == 0) == (rb == 0), (name, mode, ra, rb)
            if ra:
                continue                  # this lobe has no filter_o
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (name, mode, p, wo, arg, a, b)
            checked += 1
    assert checked >= 4500
