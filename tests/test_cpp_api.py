"""The C++ side of the boundary: include/OSL/oslexec.h mirrors the reference's API in namespace
OSL (ShadingSystem, ShaderGroup, RendererServices, BatchedExecutor<W>::execute with
Wide<const int,W> / BatchedShaderGlobals<W>, add_symlocs, contexts ...), examples/testshade_b200.cpp
is testshade written against it with the reference's call sites, examples/testrender_b200.cpp
drives the path tracer through the C ABI."""
import os
import subprocess

import numpy as np
import pytest

import helpers

EX = os.path.join(helpers.ROOT, "examples")


def _build(b200lib, name):
    src, exe = os.path.join(EX, name + ".cpp"), os.path.join(EX, name)
    libdir = os.path.dirname(b200lib.library_path())
    deps = [src, b200lib.library_path(), os.path.join(helpers.ROOT, "include", "OSL", "oslexec.h"),
            os.path.join(helpers.ROOT, "include", "osl_b200.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", src, "-o", exe, "-L" + libdir, "-losl_b200",
                            "-Wl,-rpath," + libdir], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    return exe


@pytest.fixture(scope="module")
def harness(b200lib):
    return _build(b200lib, "testshade_b200")


@pytest.fixture(scope="module")
def render_harness(b200lib):
    return _build(b200lib, "testrender_b200")


SP = os.path.join(helpers.GOLDEN, "oso")


def test_cpp_harness_jit_without_gpu(harness):
    r = subprocess.run([harness, "-g", "8", "8", "--searchpath", SP, "--jitonly", "-o", "Cout", "/dev/null",
                        "noise_test"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "cubin" in r.stdout


def test_cpp_harness_reports_missing_shader(harness):
    r = subprocess.run([harness, "--jitonly", "no_such_shader"], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not find shader" in r.stderr


def test_reference_style_translation_unit_compiles(b200lib, tmp_path):
    """A translation unit that only uses reference spellings (serialized group spec, ReParameter,
    register_closure, SymLocationDesc in the UserData arena, contexts, batched<8>) compiles and
    links against the header + library, and JITs without a GPU."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <OSL/oslexec.h>
using namespace OSL;
struct MyParams { Vec3 N; };
int main() {
    RendererServices rs;
    ShadingSystem ss(&rs, nullptr, nullptr);
    ss.attribute("searchpath:shader", "%s");
    ss.attribute("options", "llvm_jit_fma=0,b200_block=128");
    ClosureParam params[] = { { TypeNormal, 0, nullptr, (int)sizeof(Vec3) }, CLOSURE_FINISH_PARAM(MyParams) };
    ss.register_closure("diffuse", 3, params, nullptr, nullptr);
    const char* nm = "diffuse"; int id = -1;
    if (!ss.query_closure(&nm, &id, nullptr) || id != 3) return 3;
    ShaderGroupRef g = ss.ShaderGroupBegin("grp", "surface",
        "param float Kd 0.25; shader layers_lazy_a alayer; shader layers_lazy_b blayer; shader layers_lazy_c clayer;"
        "connect alayer.f_out clayer.f_in; connect alayer.c_out clayer.c_in, connect blayer.out clayer.unused;");
    if (!g) { fprintf(stderr, "%%s\n", ss.geterror().c_str()); return 4; }
    ss.ShaderGroupEnd(*g);
    SymLocationDesc locs[] = { SymLocationDesc("alayer.f_out", TypeFloat, true, SymArena::Outputs, 0, 48),
                               SymLocationDesc("alayer.c_out", TypeColor, true, SymArena::Outputs, 12, 48) };
    ss.add_symlocs(g.get(), locs, 2);
    if (!ss.find_symloc(g.get(), ustring("alayer.c_out"), SymArena::Outputs)) return 5;
    PerThreadInfo* ti = ss.create_thread_info();
    ShadingContext* ctx = ss.get_context(ti);
    auto ex = ss.batched<8>();
    ex.jit_group(g.get(), ctx);
    if (ss.has_error()) { fprintf(stderr, "%%s\n", ss.geterror().c_str()); return 6; }
    float kd = 0.5f;
    if (!ss.ReParameter(*g, "alayer", "Kd", TypeFloat, &kd) || g->handle) return 7;
    if (!ss.optimize_group(g.get())) return 8;
    std::string cu;
    if (!ss.getattribute(g.get(), "b200_cuda_source", cu) || cu.find("layer_0") == std::string::npos) return 9;
    ss.release_context(ctx);
    ss.destroy_thread_info(ti);
    return ss.raytype_bit(ustring("shadow")) == 2 ? 0 : 10;
}
''' % SP)
    exe = tmp_path / "t"
    libdir = os.path.dirname(b200lib.library_path())
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(helpers.ROOT, "include"), str(src), "-o",
                        str(exe), "-L" + libdir, "-losl_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)


def _oracle(layers, conns, outputs, res, nf, center=False, userdata=False, **gl):
    from oracle import oracle
    g = oracle.OracleGroup(layers, conns, outputs)
    var, uni = oracle.testshade_globals(res, res, center=center, **gl)
    if userdata:
        uni["userdata"] = helpers.testshade_userdata(res * res, var, uni)
    want = np.zeros((res * res, nf), np.float32)
    g.run(res * res, var, uni, want)
    return want


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["tile", "batched16"])
def test_cpp_harness_matches_oracle(harness, tmp_path, mode):
    """testshade through the reference API: one whole-grid call and the reference's own loop of
    BatchedExecutor<16>::execute calls give the oracle's bytes (res 50: ragged last batch)."""
    out = str(tmp_path / "cout.f32")
    res = 50
    cmd = [harness, "-g", str(res), str(res), "--searchpath", SP, "--fma", "0", "-o", "Cout", out]
    r = subprocess.run(cmd + (["--batched16"] if mode == "batched16" else []) + ["noise_test"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, np.float32).reshape(res * res, 3)
    layers, outputs, _ = helpers.image_case_group("noise")
    want = _oracle(layers, (), outputs, res, 3)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_cpp_layered_group_with_connections_and_derivs(harness, tmp_path):
    """The 3-layer layers-lazy group (BASELINE config 2's group) declared through Parameter /
    Shader / ConnectShaders, outputs with derivatives placed by add_symlocs."""
    res = 40
    fo, co = str(tmp_path / "f.f32"), str(tmp_path / "c.f32")
    r = subprocess.run([harness, "-g", str(res), str(res), "--searchpath", SP, "--fma", "0",
                        "--layer", "alayer", "layers_lazy_a", "--layer", "blayer", "layers_lazy_b",
                        "--layer", "clayer", "layers_lazy_c",
                        "--connect", "alayer", "f_out", "clayer", "f_in", "--connect", "alayer", "c_out", "clayer", "c_in",
                        "--connect", "blayer", "out", "clayer", "unused", "-od3", "alayer.c_out", co],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(co, np.float32).reshape(res * res, 9)
    layers, conns, outputs = helpers.layers_group(derivs=True)
    want = _oracle(layers, conns, outputs, res, 12)
    assert np.array_equal(got.view(np.uint32), want[:, 3:].view(np.uint32))
    assert np.abs(got[:, 3:]).max() > 0          # the derivatives are really there


@pytest.mark.gpu
def test_cpp_userdata_through_renderer_services(harness, tmp_path):
    """Interpolated parameters fed by RendererServices::get_userdata (the harness's SimpleRenderer
    supplies red / green / blue for some points only): testsuite/userdata-partial through the
    reference's lane interface equals the oracle."""
    res = 33
    out = str(tmp_path / "c.f32")
    r = subprocess.run([harness, "-g", str(res), str(res), "--center", "--searchpath", SP, "--fma", "0", "--batched16",
                        "-o", "Cout", out, "userdata_partial_test"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, np.float32).reshape(res * res, 3)
    layers = [dict(oso=helpers.oso("userdata_partial_test"), name="l0")]
    want = _oracle(layers, (), [dict(name="Cout", offset=0, stride=12)], res, 3, center=True, userdata=True)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.gpu
def test_cpp_render_through_the_c_abi(b200lib, render_harness, tmp_path):
    """render-cornell driven from C++ (b200_render_create + two interleaved tile work sets) gives
    the pixels of the Python binding, which the render tests pin to the oracle bit for bit."""
    from openshadinglanguage_b200 import api
    from openshadinglanguage_b200.render import scene as sc
    S = sc.load_scene(os.path.join(helpers.GOLDEN, "scenes", "cornell.xml"))
    A = S.prepare()
    res, aa = 96, 3
    blob, out = str(tmp_path / "cornell.b200scene"), str(tmp_path / "img.f32")
    api.write_scene_blob(blob, S, A, helpers.oso, res, res, aa)
    r = subprocess.run([render_harness, blob, out, "fma=0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, np.float32).reshape(res, res, 3)
    want = api.Renderer(S, A, helpers.oso, res, res, aa, options="fma=0").render()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
