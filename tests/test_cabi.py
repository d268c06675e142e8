"""CPU-side checks of the product boundary: the C-ABI library builds, loads,
exports every symbol include/osl_b200.h declares, generates + NVRTC-compiles
sm_100a code for the fixture groups without a GPU, reports errors through
status codes, and fails loudly (no fallback) when no device is present."""
import ctypes
import os
import re

import pytest

import helpers


def _declared_symbols():
    hdr = open(os.path.join(helpers.ROOT, "include", "osl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(b200lib):
    L = ctypes.CDLL(b200lib.library_path())
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), "libosl_b200.so does not export %s" % s
    assert L.b200_abi_version() == 3


@pytest.mark.parametrize("case", ["noise", "pnoise", "cellnoise", "noise-perlin"])
def test_groups_compile_to_sm100a_cubin(b200lib, case):
    layers, outputs, _ = helpers.image_case_group(case)
    g = b200lib.ShaderGroup(layers, outputs=outputs, options="fma=0")
    src = g.cuda_source
    assert "osl_b200_group_kernel" in src and "layer_0" in src
    cubin = g.cubin
    assert cubin[:4] == b"\x7fELF" and len(cubin) > 1000
    # only the globals the shader reads are loaded
    assert g.reads_global("u") and g.reads_global("v")
    assert not g.reads_global("P") and not g.reads_global("N")


def test_layers_group_analysis(b200lib):
    layers, conns, outputs = helpers.layers_group()
    g = b200lib.ShaderGroup(layers, conns, outputs)
    src = g.cuda_source
    # blayer feeds only an unread param: never lazily pulled, so never called
    assert "layer_1(sg, gd, L);" not in src
    # alayer owns renderer outputs => not lazy: the entry runs it unconditionally
    assert "layer_0(sg, gd, L);" in src
    # outputs with derivs => u,v derivatives are read
    assert g.reads_global("dudx") and g.reads_global("dvdy")
    assert any("printf" in w for w in g.warnings)


def test_lazy_layer_is_guarded(b200lib):
    layers, conns, _ = helpers.layers_group(with_outputs=False)
    # give the entry layer something observable so the group is not empty
    g = b200lib.ShaderGroup(layers, conns, [])
    assert "if (!(gd.ran & 1u)) layer_0(sg, gd, L);" in g.cuda_source


def test_error_reporting(b200lib):
    with pytest.raises(b200lib.B200Error, match="not an OSO file"):
        b200lib.ShaderGroup([dict(oso="garbage", name="x")])
    layers, outputs, _ = helpers.image_case_group("noise")
    with pytest.raises(b200lib.B200Error, match="no parameter"):
        b200lib.ShaderGroup([dict(oso=layers[0]["oso"], name="l", params=dict(nosuch=1.0))])
    with pytest.raises(b200lib.B200Error, match="not found"):
        b200lib.ShaderGroup(layers, outputs=[dict(name="nosuch", offset=0, stride=4)])
    with pytest.raises(b200lib.B200Error, match="unknown layer"):
        b200lib.ShaderGroup(layers, connections=[("a", "b", "c", "d")])


def test_unsupported_op_is_an_error_not_a_fallback(b200lib):
    oso = helpers.oso("noise_test").replace("\tnoise\t", "\tgabornoisezz\t", 1)
    with pytest.raises(b200lib.B200Error, match="not implemented"):
        b200lib.ShaderGroup([dict(oso=oso, name="l")], outputs=[dict(name="Cout", offset=0, stride=12)])


def test_execute_without_gpu_fails_loudly(b200lib):
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    layers, outputs, _ = helpers.image_case_group("noise")
    g = b200lib.ShaderGroup(layers, outputs=outputs)
    var, uni = b200lib.grid_globals(4, 4)
    out = np.zeros((16, 3), np.float32)
    with pytest.raises(b200lib.B200Error):
        g.execute_host(16, var, uni, out)
