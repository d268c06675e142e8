"""Pin the CPU oracle against the reference's own golden data (CPU only).

Every check here compares oracle output with data published by the reference:
known-answer vectors of src/liboslnoise/oslnoise_test.cpp, the exact integer
hashes of testsuite/hash/ref/out.txt, the golden images of the noise tests
(reference thresholds: failthresh 0.004 on [0,1] = 1 LSB of uint8, failpercent
0.05 %; testsuite/noise/run.py) and the lazy-evaluation text goldens.
"""
import os

import numpy as np
import pytest

import helpers
from oracle import oracle

EPS = 1e-3  # oslnoise_test.cpp:27


def _grid_points(N=4):
    return np.array([i / N for i in range(2 * N + 1)], np.float32)


def test_perlin_known_answers():
    v = helpers.noise_vectors()["test_perlin"]
    x = _grid_points()
    for dim in (1, 2, 3, 4):
        inp = np.stack([x] * dim)
        s = oracle.noise("snoise", 1, inp)[0]
        np.testing.assert_allclose(s, v["results_%dd" % dim], atol=EPS)
        u = oracle.noise("noise", 1, inp)[0]
        np.testing.assert_allclose(u, 0.5 + 0.5 * s, atol=EPS)
        vs = oracle.noise("snoise", 3, inp).T
        np.testing.assert_allclose(vs, np.array(v["vresults_%dd" % dim]), atol=EPS)


@pytest.mark.parametrize("kind,table", [("cellnoise", "test_cell"), ("hashnoise", "test_hash")])
def test_cell_hash_known_answers(kind, table):
    v = helpers.noise_vectors()[table]
    x = np.array([0.5, 1.5], np.float32)
    for dim in (1, 2, 3, 4):
        inp = np.stack([x] * dim)
        np.testing.assert_allclose(oracle.noise(kind, 1, inp)[0], v["results_%dd" % dim], atol=EPS)
        np.testing.assert_allclose(oracle.noise(kind, 3, inp).T, np.array(v["vresults_%dd" % dim]),
                                   atol=EPS)


def test_cell_boundaries():
    # oslnoise_test.cpp:253-263: cellnoise is constant on [i, i+1) incl. negatives
    def c(x):
        return oracle.noise("cellnoise", 1, np.array([[x]], np.float32))[0, 0]
    assert c(1.0001) == c(1.0)
    assert c(0.9999) != c(1.0)
    assert c(0.9999) == c(0.0)
    assert c(-0.0001) == c(-1.0)
    assert c(-0.9999) == c(-1.0)
    assert c(-1.0001) == c(-2.0)


def test_perlin_continuity():
    # oslnoise_test.cpp:176-181
    def n(x):
        return oracle.noise("noise", 1, np.array([[x]], np.float32))[0, 0]
    for a, b in [(0.9999, 1.0), (1.0001, 1.0), (-0.0001, 0.0), (-1.0001, -1.0), (-0.9999, -1.0)]:
        assert abs(n(a) - n(b)) < 1e-3


def test_hash_text_golden_exact():
    """testsuite/hash: `testshade -g 2 2 -center test` prints exact int hashes."""
    g = oracle.OracleGroup([dict(oso=helpers.oso("hash_test"), name="l0")])
    var, uni = oracle.testshade_globals(2, 2, center=True)
    txt = g.run_capture(4, var, uni)
    want = helpers.golden_text("hash").split("\n", 1)[1]  # drop "Compiled ..." line
    assert txt.rstrip("\n") == want.rstrip("\n")


def test_layers_lazy_text_golden():
    """testsuite/layers-lazy: C runs first, pulls A lazily; B never runs."""
    layers, conns, _ = helpers.layers_group(with_outputs=False)
    g = oracle.OracleGroup(layers, conns)
    var, uni = oracle.testshade_globals(2, 2)
    txt = g.run_capture(4, var, uni)
    want = "\n".join(l for l in helpers.golden_text("layers-lazy").split("\n")
                     if not l.startswith(("Compiled", "Connect")))
    assert txt.rstrip("\n") == want.rstrip("\n")
    assert "Running layer B" not in txt


def test_layers_text_golden():
    layers = [dict(oso=helpers.oso("layers_a"), name="alayer"),
              dict(oso=helpers.oso("layers_b"), name="blayer")]
    conns = [("alayer", "f_out", "blayer", "f_in"), ("alayer", "c_out", "blayer", "c_in")]
    g = oracle.OracleGroup(layers, conns)
    var, uni = oracle.testshade_globals(1, 1)
    txt = g.run_capture(1, var, uni)
    want = "\n".join(l for l in helpers.golden_text("layers").split("\n")
                     if not l.startswith(("Compiled", "Connect")))
    assert txt.rstrip("\n") == want.rstrip("\n")


@pytest.mark.parametrize("case", ["color", "transformc"])
def test_color_text_goldens(case):
    """testsuite/color and testsuite/transformc (`testshade test`, one point):
    colour constructors with space names, luminance, transformc between
    rgb/hsv/hsl/YIQ/XYZ/xyY/sRGB incl. derivatives, printed at %g / %0.3f."""
    g = oracle.OracleGroup([dict(oso=helpers.oso(case + "_test"), name="l0")])
    var, uni = oracle.testshade_globals(1, 1)
    txt = g.run_capture(1, var, uni)
    want = "\n".join(l for l in helpers.golden_text(case).split("\n") if not l.startswith("Compiled"))
    assert txt.rstrip("\n") == want.rstrip("\n")


@pytest.mark.parametrize("case", ["matrix", "transform"])
def test_matrix_text_goldens(case):
    """testsuite/matrix and testsuite/transform (`testshade -g 2 2 test`): matrix
    constructors (with space names, from-to), getmatrix incl. an unknown space,
    element access, * / unary -, transpose, determinant, ==, and transform /
    transformv / transformn of points, vectors, normals with derivatives between
    "common", "shader", "object" and the renderer-named "myspace"."""
    g = oracle.OracleGroup([dict(oso=helpers.oso(case + "_test"), name="l0")])
    var, uni = oracle.testshade_globals(2, 2)
    txt = g.run_capture(4, var, uni)
    want = "\n".join(l for l in helpers.golden_text(case).split("\n") if not l.startswith("Compiled"))
    assert txt.rstrip("\n") == want.rstrip("\n")


@pytest.mark.parametrize("d", sorted(helpers.TESTSUITE_TEXT))
def test_testsuite_text_goldens(d):
    """40 more `testshade` text tests of the reference (arithmetic, trig, exponential, hyperb,
    miscmath, geomath, blendmath, logic, loops, functions, arrays, derivs, vector / matrix
    constructors with spaces, splineinverse with derivatives, ...): the oracle's output equals
    the reference's golden text character for character."""
    gx, gy, center = helpers.TESTSUITE_TEXT[d]
    g = oracle.OracleGroup([dict(oso=helpers.oso("ts_" + d), name="l0")])
    var, uni = oracle.testshade_globals(gx, gy, center=bool(center))
    uni["userdata"] = helpers.testshade_userdata(gx * gy, var, uni)    # SimpleRenderer::get_userdata
    txt = g.run_capture(gx * gy, var, uni)
    assert txt.rstrip("\n") == helpers.testsuite_text_want(d).rstrip("\n")


@pytest.mark.parametrize("case,xres,yres", [("blackbody", 1000, 64), ("wavelength_color", 1000, 64)])
def test_color_exr_goldens(case, xres, yres):
    """testsuite/blackbody (`-g 1000 64 -od half`) and testsuite/wavelength_color
    (`-od float`): image thresholds of the tests (failthresh 0.004, 0.05 % of pixels)."""
    g = oracle.OracleGroup([dict(oso=helpers.oso(case + "_test"), name="l0")],
                           outputs=[dict(name="Cout", offset=0, stride=12)])
    var, uni = oracle.testshade_globals(xres, yres)
    out = np.zeros((xres * yres, 3), np.float32)
    g.run(xres * yres, var, uni, out, nthreads=4)
    ref = np.load(os.path.join(helpers.GOLDEN, "images", case + ".npz"))["pixels"]
    img = out.reshape(yres, xres, 3)
    assert ref.shape == img.shape
    if case == "blackbody":     # the test writes half pixels: compare what the file would hold
        img = img.astype(np.float16).astype(np.float32)
    d = np.abs(img - ref).max(axis=2)
    # one half ulp at the image's magnitude (values reach 8728) is the tightest meaningful bar
    ulp = np.maximum(np.abs(ref).max(axis=2), 1e-4) * (2.0 ** -10 if case == "blackbody" else 2.0 ** -22)
    assert (d > np.maximum(0.004, ulp)).mean() <= 0.0005, d.max()
    assert np.all(d <= 2 * ulp + 1e-7), (d / ulp).max()


@pytest.mark.parametrize("case", sorted(helpers.IMAGE_CASES))
def test_golden_images(case):
    layers, outputs, res = helpers.image_case_group(case)
    g = oracle.OracleGroup(layers, outputs=outputs)
    var, uni = oracle.testshade_globals(res, res)
    out = np.zeros((res * res, 3), np.float32)
    g.run(res * res, var, uni, out, nthreads=4)
    img = helpers.quantize_u8(out).reshape(res, res, 3)
    ref, step, shape = helpers.golden_image(case)
    assert shape == (res, res)
    d = np.abs(img[::step, ::step].astype(int) - ref.astype(int))
    # reference thresholds: <= 1 LSB everywhere except 0.05 % of pixels
    assert (d > 1).mean() <= 0.0005, "pixels beyond 1 LSB: %g" % (d > 1).mean()
    # in practice the restatement reproduces the goldens exactly
    assert d.max() == 0


def test_spline_golden_images():
    """testsuite/spline: five outputs (values and Dx derivatives, five bases chosen by a
    run-time string, knot arrays with derivatives) against the reference images."""
    layers, outputs, nfloats = helpers.spline_case()
    g = oracle.OracleGroup(layers, outputs=outputs)
    var, uni = oracle.testshade_globals(256, 256)
    arena = np.zeros(nfloats, np.float32)
    g.run(256 * 256, var, uni, arena, nthreads=4)
    helpers.check_spline_images(arena)


def test_oracle_globals_match_product_harness():
    """The product's grid harness and the oracle's restatement of testshade's
    setup_shaderglobals agree bit-for-bit."""
    from openshadinglanguage_b200.testshade import grid_globals
    for kw in (dict(), dict(center=True), dict(vary_udxdy=True, vary_vdxdy=True, vary_pdxdy=True)):
        va, ua = oracle.testshade_globals(7, 5, **kw)
        vb, ub = grid_globals(7, 5, **kw)
        assert set(va) == set(vb)
        for k in va:
            assert np.array_equal(np.asarray(va[k]).ravel(), np.asarray(vb[k]).ravel()), k
        for k in set(ua) | set(ub):
            if k == "transforms":
                assert set(ua[k]) == set(ub[k])
                for name in ua[k]:
                    assert np.array_equal(np.float32(ua[k][name]), np.float32(ub[k][name])), name
                continue
            assert list(map(float, ua[k])) == list(map(float, ub[k])), k
