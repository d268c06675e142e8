"""SURVEY 8 row f.2, the runtime-optimizer subset the back end needs.  The reference folds
constants in its own optimizer (src/liboslexec/constfold.cpp, 3069 lines) before it JITs; this
back end hands instance values to the generator as C++ literals in fully inlined code and lets
NVRTC fold them.  The test proves that the folding really happens: a group whose math depends on
instance parameters only compiles to a kernel without the Perlin lattice hash or any special-function
instruction and less than half the size, while the same shader with one per-point input keeps it all."""
import re
import subprocess

import numpy as np
import pytest

import helpers


def _sass(b200lib, params, tmp_path, tag):
    g = b200lib.ShaderGroup([dict(oso=helpers.oso("constfold_ops"), name="l0", params=params)], (),
                            [dict(name="Cout", offset=0, stride=12)], options="fma=1,stage=0")
    p = tmp_path / (tag + ".cubin")
    p.write_bytes(g.cubin)
    r = subprocess.run(["cuobjdump", "-sass", str(p)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    body = r.stdout.split("Function : osl_b200_group_kernel")[1]
    return [m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)], g


def test_instance_constants_fold_away(b200lib, tmp_path):
    const_ops, g = _sass(b200lib, dict(a=0.3, b=2.5, vary=0.0), tmp_path, "const")
    vary_ops, _ = _sass(b200lib, dict(a=0.3, b=2.5, vary=1.0), tmp_path, "vary")
    hashing = lambda ops: sum(o.startswith(("LOP3", "SHF")) for o in ops)        # the Perlin lattice hash
    assert not any(o.startswith("MUFU") for o in const_ops)      # sqrt / pow / exp / log / division: all folded
    assert hashing(const_ops) * 6 < hashing(vary_ops), (hashing(const_ops), hashing(vary_ops))
    assert len(const_ops) * 2 < len(vary_ops), (len(const_ops), len(vary_ops))  # what is left is the tile loop + stores


@pytest.mark.gpu
def test_folded_kernel_matches_oracle(b200lib, cuda_device):
    import torch
    from oracle import oracle
    layers = [dict(oso=helpers.oso("constfold_ops"), name="l0", params=dict(a=0.3, b=2.5, vary=0.0))]
    outputs = [dict(name="Cout", offset=0, stride=12)]
    n = 64
    var, uni = oracle.testshade_globals(8, 8)
    want = np.zeros((n, 3), np.float32)
    oracle.OracleGroup(layers, outputs=outputs).run(n, var, uni, want)
    g = b200lib.ShaderGroup(layers, (), outputs, options="fma=0")
    gvar, guni = b200lib.grid_globals(8, 8)
    dvar = {k: torch.from_numpy(v).to(cuda_device) for k, v in gvar.items()}
    out = torch.zeros((n, 3), dtype=torch.float32, device=cuda_device)
    g.execute(n, dvar, guni, out)
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
