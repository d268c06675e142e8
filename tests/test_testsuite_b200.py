"""The reference's BATCHED testsuite, replayed (src/cmake/testing.cmake:214-235
registers every testsuite directory holding a BATCHED marker a second time for the batched back
end - this is that list run through this library).

tools/testsuite_b200.py (run in the build container, where /root/reference exists) executes each
directory's run.py with stubbed command builders, compiles its shaders with tools/mini_oslc.py,
builds every `testshade` command's group through the product's generator + NVRTC, replays it on
the CPU oracle and compares with the reference's ref/out.txt and ref images.  Directories that
pass leave a bundle (commands, .oso, expected text, reference images) in
tests/golden/testsuite_b200/, and the GPU test below replays each bundle through the device.
manifest.json keeps the status of all 171 directories, including why the others do not pass."""
import collections
import json
import os
import re

import numpy as np
import pytest

import helpers

DIR = os.path.join(helpers.GOLDEN, "testsuite_b200")
MANIFEST = json.load(open(os.path.join(DIR, "manifest.json")))
PASSING = sorted(d for d, e in MANIFEST.items() if e["status"] == "pass")


def _bundle(d):
    b = json.load(open(os.path.join(DIR, d + ".json")))
    npz = os.path.join(DIR, d + ".npz")
    imgs = {}
    if os.path.exists(npz):
        z = np.load(npz)
        for fn, meta in b["images"].items():
            a = z[re.sub(r"\W", "_", fn)].astype(np.float32)
            imgs[fn] = (a / 255.0 if meta["kind"] == "uint8" else a, meta["kind"])
    return b, imgs


def test_manifest_covers_the_batched_testsuite():
    """171 directories carry a BATCHED marker in the reference; every one has a status and a reason
    when it does not pass; every passing one has its bundle."""
    assert len(MANIFEST) == 171
    c = collections.Counter(e["status"] for e in MANIFEST.values())
    assert set(c) <= {"pass", "harness", "oslc", "codegen", "oracle", "mismatch"}
    assert c["pass"] >= 102, c
    for d, e in MANIFEST.items():
        assert (e["status"] == "pass") == os.path.exists(os.path.join(DIR, d + ".json")), d
        assert e["status"] == "pass" or e["reason"], d


@pytest.mark.parametrize("d", PASSING)
def test_bundle_commands_parse_and_compile(b200lib, d):
    """Every passing directory's commands parse completely and their groups go through the
    product's generator and NVRTC to an sm_100a cubin (no GPU needed)."""
    from openshadinglanguage_b200 import testshade as tsh
    b, _ = _bundle(d)
    for cmd in [c for c in b["commands"] if not c.startswith("\x00echo ")][:2]:
        spec = tsh.parse_command(cmd)
        assert not spec["unsupported"], spec["unsupported"]
        layers = [dict(oso=b["oso"][l["shader"]], name=l["name"], params=l["params"]) for l in spec["layers"]]
        g = b200lib.ShaderGroup(layers, spec["connections"], (), options="fma=0,journal=1")
        assert g.cubin[:4] == b"\x7fELF"


@pytest.mark.gpu
@pytest.mark.parametrize("d", PASSING)
def test_reference_testsuite_directory_through_the_device(b200lib, cuda_device, d):
    """The directory's commands through the GPU path (strict mode, printf journal): the text equals
    the reference's ref/out.txt character for character and every output image is inside the
    test's own idiff thresholds."""
    import torch
    from openshadinglanguage_b200 import testshade as tsh
    b, ref_images = _bundle(d)

    class DeviceRunner:
        def __init__(self, layers, conns, outs, spec):
            self.spec = spec
            n = spec["xres"] * spec["yres"]
            var, uni = b200lib.grid_globals(spec["xres"], spec["yres"])
            self.arena, descs = b200lib.pack_userdata(helpers.testshade_userdata(n, var, uni, extra=spec["userdata"]))
            self.g = b200lib.ShaderGroup(layers, conns, outs, options="fma=0,journal=1" + (
                ",error_repeats=1" if "error_repeats=1" in spec["options"] else "") + (
                ",colorspace=" + spec["colorspace"] if spec.get("colorspace") else ""), userdata=descs,
                                  name=spec.get("groupname") or "group",
                                  attributes=tsh.harness_attributes(spec["xres"], spec["yres"]))

        def run(self, n, var, uni, arena):
            ud, _ = b200lib.pack_userdata(helpers.testshade_userdata(n, var, uni, extra=self.spec["userdata"]))
            dvar = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in var.items()}
            out = torch.from_numpy(arena).to(cuda_device)
            self.g.execute(n, dvar, uni, out, userdata=torch.from_numpy(ud).to(cuda_device))
            arena[:] = out.cpu().numpy()
            return self.g.journal()

    texts, images = [], {}
    for cmd in b["commands"]:
        if cmd.startswith("\x00echo "):         # an `echo TEXT >> out.txt` line of run.py between the commands
            texts.append(cmd[6:])
            continue
        r = tsh.run_command(tsh.parse_command(cmd), lambda name: b["oso"][name], DeviceRunner, b200lib.grid_globals)
        if r["text"]:
            texts.append(r["text"])
        for var, (fn, img) in r["images"].items():
            if fn != "null":
                images[fn] = img
    if "out.txt" in b["settings"]["outputs"] and b["want"]:
        assert "\n".join(texts).rstrip("\n") == b["want"].rstrip("\n")
    st = b["settings"]
    for fn, (ref, kind) in ref_images.items():
        assert fn in images, fn
        assert tsh.compare_image(images[fn], ref, kind, st["failthresh"], st["failpercent"], st["hardfail"]) is None, fn
