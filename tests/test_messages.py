"""error() / warning() through the journal and getattribute() of renderer-supplied values
(SURVEY 8 row f.4): reference llvm_gen_printf's error / warning flavours and the error handler's
duplicate suppression (testsuite/error-dupes), llvm_gen_getattribute + SimpleRenderer's
get_attribute -> get_userdata fallback (testsuite/userdata-custom)."""
import numpy as np
import pytest

import helpers
from oracle import oracle


def _error_dupes_want():
    txt = helpers.golden_text("error-dupes")
    first = txt.split("Without repeated errors:\n")[1].split("With repeated errors:\n")[0]
    second = txt.split("With repeated errors:\n")[1]
    return first, second


def test_oracle_error_and_warning_match_reference_text():
    first, _ = _error_dupes_want()
    g = oracle.OracleGroup([dict(oso=helpers.oso("error_dupes_test"), name="l0")])
    var, uni = oracle.testshade_globals(2, 2)
    assert g.run_capture(4, var, uni).rstrip("\n") == first.rstrip("\n")


def test_oracle_getattribute_reads_userdata():
    g = oracle.OracleGroup([dict(oso=helpers.oso("userdata_custom_test"), name="l0")])
    var, uni = oracle.testshade_globals(1, 1)
    uni["userdata"] = helpers.testshade_userdata(1, var, uni, extra=[("testFloat", np.array([42], np.float32)),
                                                                    ("testColor", np.array([1, 2, 3], np.float32))])
    want = "\n".join(l for l in helpers.golden_text("userdata-custom").split("\n") if not l.startswith("Compiled"))
    assert g.run_capture(1, var, uni).rstrip("\n") == want.rstrip("\n")


def _message_group():
    layers = [dict(oso=helpers.oso("message_a"), name="a", params=dict(Kd=0.8)),
              dict(oso=helpers.oso("message_b"), name="b")]
    conns = [("a", "f_out", "b", "f_in")]
    names = [("Cfoo", 3), ("Fscalar", 1), ("Fmissing", 1), ("Fwrongtype", 1), ("Fsometimes", 1), ("Results", 3)]
    outputs, off = [], 0
    for n, c in names:
        outputs.append(dict(name="b." + n, offset=off, stride=40))
        off += 4 * c
    return layers, conns, outputs


def test_oracle_messages_between_layers():
    """setmessage in an upstream layer, getmessage downstream (opmessage.cpp): found + same type
    -> 1 and the value; never set or set with another type -> 0 and the destination untouched;
    a message set under a condition exists only for those points."""
    layers, conns, outputs = _message_group()
    res = 8
    var, uni = oracle.testshade_globals(res, res)
    out = np.zeros((res * res, 10), np.float32)
    oracle.OracleGroup(layers, conns, outputs).run(res * res, var, uni, out)
    u, v = var["u"], var["v"]
    f_in = np.float32(0.8) * u
    assert np.allclose(out[:, 0], np.float32(0.4) * f_in) and np.allclose(out[:, 1], u * f_in)
    assert np.allclose(out[:, 3], f_in * 3)
    assert (out[:, 4] == -2).all() and (out[:, 5] == -3).all()
    assert np.array_equal(out[:, 6], np.where(u > 0.5, v, np.float32(-4)))
    assert (out[:, 7] == 3).all()                      # r1 + 2 r2, no r3 / r4
    assert np.array_equal(out[:, 8], (u > 0.5).astype(np.float32)) and (out[:, 9] == 7).all()


@pytest.mark.gpu
def test_gpu_messages_bit_exact_vs_oracle(b200lib, cuda_device):
    import torch
    layers, conns, outputs = _message_group()
    res = 40
    n = res * res
    var, uni = oracle.testshade_globals(res, res)
    want = np.zeros((n, 10), np.float32)
    oracle.OracleGroup(layers, conns, outputs).run(n, var, uni, want)
    g = b200lib.ShaderGroup(layers, conns, outputs, options="fma=0")
    gvar, guni = b200lib.grid_globals(res, res)
    dvar = {k: torch.from_numpy(v).to(cuda_device) for k, v in gvar.items()}
    out = torch.zeros((n, 10), dtype=torch.float32, device=cuda_device)
    g.execute(n, dvar, guni, out)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_printf_format_is_checked_against_its_arguments(b200lib):
    """A hand-written .oso whose format does not fit its arguments must be a compile error, not a
    host crash when the journal is formatted (reference: "Mismatch between format string and
    arguments")."""
    oso = helpers.oso("ts_arithmetic")
    for bad in ('"%s"', '"%n"', '"%*d"'):
        lines = []
        done = False
        for l in oso.split("\n"):
            if not done and l.startswith("const\tstring") and "%" in l:
                l = l.split("\t")
                l[3] = bad
                l = "\t".join(l)
                done = True
            lines.append(l)
        assert done
        with pytest.raises(b200lib.B200Error, match="format"):
            b200lib.ShaderGroup([dict(oso="\n".join(lines), name="l0")], (), (), options="journal=1")


@pytest.mark.gpu
def test_gpu_error_dupes_text(b200lib, cuda_device):
    import torch
    first, second = _error_dupes_want()
    var, uni = b200lib.grid_globals(2, 2)
    dvar = {k: torch.from_numpy(v).to(cuda_device) for k, v in var.items()}
    out = torch.zeros(16, dtype=torch.float32, device=cuda_device)
    layers = [dict(oso=helpers.oso("error_dupes_test"), name="l0")]
    g = b200lib.ShaderGroup(layers, (), (), options="fma=0,journal=1")
    g.execute(4, dvar, uni, out)
    assert g.journal().rstrip("\n") == first.rstrip("\n")
    g = b200lib.ShaderGroup(layers, (), (), options="fma=0,journal=1,error_repeats=1")
    g.execute(4, dvar, uni, out)
    assert g.journal().rstrip("\n") == second.rstrip("\n")


@pytest.mark.gpu
def test_gpu_getattribute_reads_userdata(b200lib, cuda_device):
    import torch
    var, uni = b200lib.grid_globals(1, 1)
    arena, descs = b200lib.pack_userdata(helpers.testshade_userdata(
        1, var, uni, extra=[("testFloat", np.array([42], np.float32)), ("testColor", np.array([1, 2, 3], np.float32))]))
    g = b200lib.ShaderGroup([dict(oso=helpers.oso("userdata_custom_test"), name="l0")], (), (),
                            options="fma=0,journal=1", userdata=descs)
    dvar = {k: torch.from_numpy(v).to(cuda_device) for k, v in var.items()}
    out = torch.zeros(16, dtype=torch.float32, device=cuda_device)
    g.execute(1, dvar, uni, out, userdata=torch.from_numpy(arena).to(cuda_device))
    want = "\n".join(l for l in helpers.golden_text("userdata-custom").split("\n") if not l.startswith("Compiled"))
    assert g.journal().rstrip("\n") == want.rstrip("\n")
