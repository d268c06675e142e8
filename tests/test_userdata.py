"""Userdata / interpolated parameters ([[ int lockgeom = 0 ]]): reference osl_bind_interpolated_param
(src/liboslexec/llvm_instance.cpp:805-970), RendererServices::get_userdata as testshade's
SimpleRenderer implements it (src/testshade/simplerend.cpp:517-590), testsuite/userdata,
userdata-partial, userdata-passthrough, userdata-defaults.

The renderer hands the values over as dense per-name arrays in a UserData arena (b200_userdata:
offset / stride like SymLocationDesc, plus a validity plane for values only some points have);
a parameter binds by name and type, points without the value run the default / init ops."""
import numpy as np
import pytest

import helpers
from oracle import oracle


def _partial_case():
    """testsuite/userdata-partial: `-g 11 12 --center -od uint8 -o Cout`"""
    xres, yres = 11, 12
    layers = [dict(oso=helpers.oso("userdata_partial_test"), name="l0")]
    outputs = [dict(name="Cout", offset=0, stride=12)]
    return layers, outputs, xres, yres


def test_oracle_userdata_partial_matches_reference_image():
    layers, outputs, xres, yres = _partial_case()
    var, uni = oracle.testshade_globals(xres, yres, center=True)
    n = xres * yres
    uni["userdata"] = helpers.testshade_userdata(n, var, uni)
    out = np.zeros((n, 3), np.float32)
    oracle.OracleGroup(layers, outputs=outputs).run(n, var, uni, out)
    img = helpers.quantize_u8(out.reshape(yres, xres, 3))
    pixels, step, shape = helpers.golden_image("userdata-partial")
    assert shape == (yres, xres)
    assert np.array_equal(img[::step, ::step], pixels)
    # every branch is hit: points with and without each of red / green / blue
    assert len({tuple(p) for p in img.reshape(-1, 3)}) > 20
    assert (out[:, 0] == 0.5).any() and (out[:, 1] == 0.5).any() and (out[:, 2] == 0.5).any()


def test_oracle_userdata_passthrough():
    """testsuite/userdata-passthrough: `--userdata:type=vector Cd 1,1,1 -o Cd --print` prints
    "Cd : 1 1 1": an OUTPUT parameter bound to userdata, then scaled by the shader."""
    layers = [dict(oso=helpers.oso("userdata_passthrough_test"), name="l0", params=dict(scale=2.0))]
    outputs = [dict(name="Cd", offset=0, stride=12)]
    var, uni = oracle.testshade_globals(1, 1)
    uni["userdata"] = helpers.testshade_userdata(1, var, uni, extra=[("Cd", np.array([1, 1, 1], np.float32))])
    out = np.zeros((1, 3), np.float32)
    oracle.OracleGroup(layers, outputs=outputs).run(1, var, uni, out)
    assert out.tolist() == [[2.0, 2.0, 2.0]]
    uni.pop("userdata")                      # not supplied: the default 0 is scaled instead
    oracle.OracleGroup(layers, outputs=outputs).run(1, var, uni, out)
    assert out.tolist() == [[0.0, 0.0, 0.0]]


def test_userdata_binding_is_generated_only_for_interpolated_params(b200lib):
    arena, descs = b200lib.pack_userdata([dict(name="red", data=np.zeros((4, 3), np.float32), derivs=True,
                                               valid=np.ones(4, np.int32)),
                                          dict(name="scale", data=np.zeros((4, 1), np.float32))])
    layers, outputs, _, _ = _partial_case()
    src = b200lib.ShaderGroup(layers, (), outputs, userdata=descs).cuda_source
    assert "L.userdata_base" in src and src.count("got_") >= 3       # red is bound, with its validity plane
    layers = [dict(oso=helpers.oso("userdata_passthrough_test"), name="l0")]
    src = b200lib.ShaderGroup(layers, (), [dict(name="Cd", offset=0, stride=12)], userdata=descs).cuda_source
    assert "got_" not in src          # `scale` is lockgeom=1: userdata of that name is ignored; Cd has no entry


@pytest.mark.gpu
@pytest.mark.parametrize("host", [False, True])
def test_gpu_userdata_partial_bit_exact_vs_oracle(b200lib, cuda_device, host):
    import torch
    layers, outputs, xres, yres = _partial_case()
    xres, yres = 97, 61                      # ragged against the CTA tile
    n = xres * yres
    var, uni = oracle.testshade_globals(xres, yres, center=True)
    ents = helpers.testshade_userdata(n, var, uni)
    want = np.zeros((n, 3), np.float32)
    oracle.OracleGroup(layers, outputs=outputs).run(n, var, dict(uni, userdata=ents), want)
    arena, descs = b200lib.pack_userdata(ents)
    g = b200lib.ShaderGroup(layers, (), outputs, options="fma=0", userdata=descs)
    gvar, guni = b200lib.grid_globals(xres, yres, center=True)
    if host:
        got = np.zeros((n, 3), np.float32)
        g.execute_host(n, gvar, guni, got, userdata=arena)
    else:
        dvar = {k: torch.from_numpy(v).to(cuda_device) for k, v in gvar.items()}
        out = torch.zeros((n, 3), dtype=torch.float32, device=cuda_device)
        g.execute(n, dvar, guni, out, userdata=torch.from_numpy(arena).to(cuda_device))
        got = out.cpu().numpy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # without an arena every point runs the init ops (0.5 grey)
    if not host:
        out.zero_()
        g.execute(n, dvar, guni, out)
        assert (out.cpu().numpy() == 0.5).all()


@pytest.mark.gpu
def test_gpu_userdata_text_goldens(b200lib, cuda_device):
    """testsuite/userdata (`-g 2 2`: s, t <- u, v through userdata, printed) and userdata-defaults
    through the device journal: the reference's golden text, character for character."""
    import torch
    for d in ("userdata", "userdata-defaults"):
        gx, gy, center = helpers.TESTSUITE_TEXT[d]
        n = gx * gy
        gvar, guni = b200lib.grid_globals(gx, gy, center=bool(center))
        arena, descs = b200lib.pack_userdata(helpers.testshade_userdata(n, gvar, guni))
        g = b200lib.ShaderGroup([dict(oso=helpers.oso("ts_" + d), name="l0")], (), (), options="fma=0,journal=1",
                                userdata=descs)
        dvar = {k: torch.from_numpy(v).to(cuda_device) for k, v in gvar.items()}
        out = torch.zeros(16, dtype=torch.float32, device=cuda_device)
        g.execute(n, dvar, guni, out, userdata=torch.from_numpy(arena).to(cuda_device))
        assert g.journal().rstrip("\n") == helpers.testsuite_text_want(d).rstrip("\n")


@pytest.mark.gpu
def test_gpu_host_path_leaves_bytes_between_sparse_output_fields_alone(b200lib, cuda_device):
    """A renderer output placed in a record the group does not fully own (a color at offset 4 of
    a 32-byte record): the host path must write the symbol's 12 bytes per point and nothing else
    (SymLocationDesc contract), exactly like the device-pointer path."""
    layers, _, res = helpers.image_case_group("noise")
    outputs = [dict(name="Cout", offset=4, stride=32)]
    n = 64 * 64
    gvar, guni = b200lib.grid_globals(64, 64)
    g = b200lib.ShaderGroup(layers, (), outputs, options="fma=0")
    rec = np.full((n, 8), 7.25, np.float32)
    g.execute_host(n, gvar, guni, rec)
    dense = np.zeros((n, 3), np.float32)
    b200lib.ShaderGroup(layers, (), [dict(name="Cout", offset=0, stride=12)], options="fma=0").execute_host(
        n, gvar, guni, dense)
    assert np.array_equal(rec[:, 1:4], dense)
    assert (rec[:, 0] == 7.25).all() and (rec[:, 4:] == 7.25).all()
