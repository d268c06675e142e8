"""The C++ mirror of the reference API (include/OSL/oslexec_b200.h) and the
testshade-style harness built on it."""
import os
import subprocess

import numpy as np
import pytest

import helpers

EXE = os.path.join(helpers.ROOT, "examples", "testshade_b200")


@pytest.fixture(scope="module")
def harness(b200lib):
    src = os.path.join(helpers.ROOT, "examples", "testshade_b200.cpp")
    libdir = os.path.dirname(b200lib.library_path())
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(
            os.path.getmtime(src), os.path.getmtime(b200lib.library_path())):
        r = subprocess.run(["g++", "-std=c++17", "-O2", src, "-o", EXE, "-L" + libdir, "-losl_b200",
                            "-Wl,-rpath," + libdir], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    return EXE


def test_cpp_harness_jit_without_gpu(harness):
    sp = os.path.join(helpers.GOLDEN, "oso")
    r = subprocess.run([harness, "-g", "8", "8", "--searchpath", sp, "--jitonly", "-o", "Cout", "/dev/null",
                        "noise_test"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cubin" in r.stdout


def test_cpp_harness_reports_missing_shader(harness):
    r = subprocess.run([harness, "--jitonly", "no_such_shader"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not find shader" in r.stderr


@pytest.mark.gpu
def test_cpp_harness_matches_oracle(harness, tmp_path):
    from oracle import oracle
    sp = os.path.join(helpers.GOLDEN, "oso")
    out = str(tmp_path / "cout.f32")
    res = 128
    r = subprocess.run([harness, "-g", str(res), str(res), "--searchpath", sp, "--fma", "0", "-o", "Cout", out,
                        "noise_test"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, np.float32).reshape(res * res, 3)
    layers, outputs, _ = helpers.image_case_group("noise")
    g = oracle.OracleGroup(layers, outputs=outputs)
    var, uni = oracle.testshade_globals(res, res)
    want = np.zeros((res * res, 3), np.float32)
    g.run(res * res, var, uni, want)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
