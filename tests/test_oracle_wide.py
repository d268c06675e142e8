"""The 16-wide batched CPU restatement (oracle/oso2cpp.py WideGen: blocks of 16 lanes, op-at-a-time
lane loops under an execution mask, like the reference's BatchedExecutor<16>) against the scalar
oracle: same groups, same inputs.  It is built with FMA contraction allowed (testshade --batched
turns llvm_jit_fma on, testshade.cpp:294-298), so floats agree to the fast-mode tolerance the GPU
path is held to (2e-6 abs); ragged batch ends (n % 16 != 0) and lazy layers under partial masks
are covered.  bench.py times it as the batched CPU baseline."""
import numpy as np
import pytest

import helpers
from oracle import oracle


def _first(var, full, n):
    """the first n points of SoA globals made for `full` points"""
    return {k: (np.asarray(v).reshape(-1, full)[:, :n].copy() if np.asarray(v).size != full
                else np.asarray(v)[:n].copy()) for k, v in var.items()}


@pytest.mark.parametrize("n", [16 * 40, 16 * 40 + 7, 5])
def test_wide_noise_group_matches_scalar(n):
    layers, outputs, _ = helpers.image_case_group("noise")
    res = 64
    var, uni = oracle.testshade_globals(res, res)
    var = _first(var, res * res, n)
    a = np.zeros((res * res, 3), np.float32)
    b = np.zeros((res * res, 3), np.float32)
    oracle.OracleGroup(layers, outputs=outputs).run(n, var, uni, a)
    oracle.OracleGroupWide(layers, outputs=outputs).run(n, var, uni, b, nthreads=3)
    assert np.abs(a - b).max() <= 2e-6
    assert not b[n:].any()                      # lanes beyond the batch end are masked off


def test_wide_layered_group_with_lazy_layers_matches_scalar():
    layers, conns, outputs = helpers.layers_group(derivs=True)
    res = 48
    n = res * res - 3
    gl = dict(vary_udxdy=True, vary_vdxdy=True, vary_pdxdy=True)
    var, uni = oracle.testshade_globals(res, res, **gl)
    var = _first(var, res * res, n)
    a = np.zeros((res * res, 12), np.float32)
    b = np.zeros((res * res, 12), np.float32)
    oracle.OracleGroup(layers, conns, outputs).run(n, var, uni, a)
    oracle.OracleGroupWide(layers, conns, outputs).run(n, var, uni, b, nthreads=2)
    assert np.abs(a - b).max() <= 2e-6


@pytest.mark.parametrize("case", ["noise-perlin", "pnoise", "cellnoise", "noise-simplex"])
def test_wide_image_cases_match_scalar_after_quantisation(case):
    """Control flow under masks (ifs, functions with return) on the reference's noise tests:
    identical 8-bit images except where FMA moves a value across a rounding boundary."""
    layers, outputs, res = helpers.image_case_group(case)
    res = 128
    var, uni = oracle.testshade_globals(res, res)
    a = np.zeros((res * res, 3), np.float32)
    b = np.zeros((res * res, 3), np.float32)
    oracle.OracleGroup(layers, outputs=outputs).run(res * res, var, uni, a)
    oracle.OracleGroupWide(layers, outputs=outputs).run(res * res, var, uni, b, nthreads=4)
    assert np.abs(a - b).max() <= 4e-6
    assert (helpers.quantize_u8(a) != helpers.quantize_u8(b)).mean() < 1e-3
