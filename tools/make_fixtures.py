#!/usr/bin/env python3
"""Generate the committed test fixtures from the reference tree.

Run in the build container (needs /root/reference; the GPU box does not have
it).  Produces, under tests/golden/:

  oso/<name>.oso        testsuite shaders compiled by tools/mini_oslc.py
  images/<test>.npz     the reference's golden images (uint8, subsampled by
                        STEP in x and y to keep the repository small; the
                        sampling grid is stored in the file)
  text/<test>.txt       the reference's golden text outputs
  noise_vectors.json    known-answer vectors parsed out of
                        src/liboslnoise/oslnoise_test.cpp:86-334

Nothing here copies reference *sources*: the .oso files are compiler output,
the rest is the reference's published golden test data.
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import mini_oslc  # noqa: E402

REF = os.environ.get("OSL_REFERENCE", "/root/reference")
TS = os.path.join(REF, "testsuite")
OUT = os.path.join(ROOT, "tests", "golden")
STEP = 2

SHADERS = {
    # fixture name: path under testsuite/
    "noise_test": "noise/test.osl",
    "cellnoise_test": "cellnoise/test.osl",
    "hashnoise_test": "hashnoise/test.osl",
    "pnoise_test": "pnoise/test.osl",
    "testnoise": "common/shaders/testnoise.osl",
    "testpnoise": "common/shaders/testpnoise.osl",
    "hash_test": "hash/test.osl",
    "layers_lazy_a": "layers-lazy/a.osl",
    "layers_lazy_b": "layers-lazy/b.osl",
    "layers_lazy_c": "layers-lazy/c.osl",
    "spline_test": "spline/test.osl",
    "gabor2d_filter_test": "noise-gabor2d-filter/test.osl",
    "gabor3d_filter_test": "noise-gabor3d-filter/test.osl",
    "matrix_test": "matrix/test.osl",
    "transform_test": "transform/test.osl",
    "color_test": "color/test.osl",
    "transformc_test": "transformc/test.osl",
    "blackbody_test": "blackbody/test.osl",
    "wavelength_color_test": "wavelength_color/test.osl",
    "layers_a": "layers/a.osl",
    "layers_b": "layers/b.osl",
    # testrender materials (same sources in render-cornell and render-bunny)
    "matte": "render-cornell/matte.osl",
    "metal": "render-cornell/metal.osl",
    "emitter": "render-cornell/emitter.osl",
    "checkerboard": "render-veachmis/checkerboard.osl",
    "phong": "render-veachmis/phong.osl",
    "ward": "render-ward/ward.osl",
    "glossy_glass": "render-microfacet/glossy_glass.osl",
    "furnace": "render-furnace-diffuse/furnace.osl",
    "rough_matte": "render-oren-nayar/rough_matte.osl",
    # MaterialX furnace tests: several tests reuse the file names matte.osl / envmap.osl for
    # different shaders, so the fixtures (and the shader names in the copied scenes) are prefixed
    "mxon_matte": "render-mx-furnace-oren-nayar/matte.osl",
    "mxburley_matte": "render-mx-furnace-burley-diffuse/matte.osl",
    "mx_envmap": "render-mx-furnace-oren-nayar/envmap.osl",
    "mxlayer_layer": "render-mx-layer/layer.osl",
    "mxlayer_envmap": "render-mx-layer/envmap.osl",
    "mf_envmap": "render-microfacet/envmap.osl",      # texture() of the HDR probe
    # MaterialX microfacet closures (libbsdl conductor / dielectric / generalized Schlick lobes)
    "mxspec_envmap": "render-mx-conductor/envmap.osl",
    "mxcond_glossy": "render-mx-conductor/glossy.osl",
    "mxdiel_glossy": "render-mx-dielectric/glossy.osl",
    "mxgs_glossy": "render-mx-generalized-schlick/glossy.osl",
    "mxsheen_sheen": "render-mx-sheen/sheen.osl",     # sheen_bsdf in both modes (Conty-Kulla, Zeltner-Burley LTC)
    "mxfsheen_sheen": "render-mx-furnace-sheen/sheen.osl",   # ... layered over diffuse in a furnace
    # refractive MaterialX lobes and participating media (MediumStack)
    "mxdielglass_glossy": "render-mx-dielectric-glass/glossy.osl",
    "mxgsglass_glossy": "render-mx-generalized-schlick-glass/glossy.osl",
    "mxmed_envmap": "render-mx-medium-vdf/envmap.osl",
    "mxmed_medium": "render-mx-medium-vdf/medium.osl",
    "mxmedglass_glossy": "render-mx-medium-vdf-glass/glossy.osl",
    "mxaniso_anisotropic": "render-mx-anisotropic-vdf/anisotropic.osl",
    "spithin_glossy_glass": "render-spi-thinlayer/glossy_glass.osl",   # thinlayer closure (spi::ThinLayerLobe)
    # bump mapping through Dx / Dy / calculatenormal of displaced P, glass and metal around it
    "bumptest": "render-bumptest/bumptest.osl",
    "bump_glass": "render-bumptest/glass.osl",
    "bump_metal": "render-bumptest/metal.osl",
    "disp": "render-displacement/disp.osl",            # displacement group: P += fBm(P) * N, run over the vertices
    # raytype() queries per bounce kind, <Background /> without a resolution
    "raytype_envmap": "render-raytypes/raytype-envmap.osl",
    "raytypes_glossy": "render-raytypes/glossy.osl",
    "raytypes_magic": "render-raytypes/magic.osl",
    # this repo's own test shaders (path relative to the repo root)
    "glossy_mix": "repo:tests/shaders/glossy_mix.osl",
    "color_ops": "repo:tests/shaders/color_ops.osl",
    "matrix_ops": "repo:tests/shaders/matrix_ops.osl",
    "texture_ops": "repo:tests/shaders/texture_ops.osl",
    "constfold_ops": "repo:tests/shaders/constfold_ops.osl",
    "message_a": "repo:tests/shaders/message_a.osl",
    "message_b": "repo:tests/shaders/message_b.osl",
    "closure_base": "repo:tests/shaders/closure_base.osl",    # closure-typed output parameter
    "closure_mix": "repo:tests/shaders/closure_mix.osl",      # closure-typed input parameters
    "error_dupes_test": "error-dupes/test.osl",
    "userdata_custom_test": "userdata-custom/test.osl",
    "noise_generic_test": "noise-generic/test.osl",
    "pnoise_generic_test": "pnoise-generic/test.osl",
    "userdata_partial_test": "userdata-partial/test.osl",
    "userdata_passthrough_test": "userdata-passthrough/test.osl",
}
# scene descriptions + model data of the testrender configs (test input data)
SCENES = {
    "cornell.xml": "render-cornell/cornell.xml",
    "bunny.xml": "render-bunny/bunny.xml",
    "bunny.obj": "render-bunny/bunny.obj",
    "veach.xml": "render-veachmis/veach.xml",
    "ward.xml": "render-ward/scene.xml",
    "furnace.xml": "render-furnace-diffuse/scene.xml",
    "oren_nayar.xml": "render-oren-nayar/scene.xml",
    # (path, {shader name in the scene: fixture name})
    "mx_furnace_oren_nayar.xml": ("render-mx-furnace-oren-nayar/scene.xml",
                                  {"matte": "mxon_matte", "envmap": "mx_envmap"}),
    "mx_furnace_burley.xml": ("render-mx-furnace-burley-diffuse/scene.xml",
                              {"matte": "mxburley_matte", "envmap": "mx_envmap"}),
    "mx_layer.xml": ("render-mx-layer/scene.xml", {"layer": "mxlayer_layer", "envmap": "mxlayer_envmap"}),
    # (microfacet.xml is this repo's own emitter-lit variant; this is the reference's scene, whose
    #  environment shader reads ../common/textures/kitchen_probe.hdr -> ../textures/ here)
    "render_microfacet.xml": ("render-microfacet/scene.xml", {"envmap": "mf_envmap"}),
    "mx_conductor.xml": ("render-mx-conductor/scene.xml", {"glossy": "mxcond_glossy", "envmap": "mxspec_envmap"}),
    "mx_dielectric.xml": ("render-mx-dielectric/scene.xml", {"glossy": "mxdiel_glossy", "envmap": "mxspec_envmap"}),
    "mx_generalized_schlick.xml": ("render-mx-generalized-schlick/scene.xml",
                                   {"glossy": "mxgs_glossy", "envmap": "mxspec_envmap"}),
    "mx_sheen.xml": ("render-mx-sheen/scene.xml", {"sheen": "mxsheen_sheen", "envmap": "mxspec_envmap"}),
    "mx_furnace_sheen.xml": ("render-mx-furnace-sheen/scene.xml", {"sheen": "mxfsheen_sheen", "envmap": "mx_envmap"}),
    "mx_burley_diffuse.xml": ("render-mx-burley-diffuse/scene.xml", {"matte": "mxburley_matte", "envmap": "mxspec_envmap"}),
    "mx_dielectric_glass.xml": ("render-mx-dielectric-glass/scene.xml",
                                {"glossy": "mxdielglass_glossy", "envmap": "mxspec_envmap"}),
    "mx_generalized_schlick_glass.xml": ("render-mx-generalized-schlick-glass/scene.xml",
                                         {"glossy": "mxgsglass_glossy", "envmap": "mxspec_envmap"}),
    "mx_medium_vdf.xml": ("render-mx-medium-vdf/scene.xml", {"medium": "mxmed_medium", "envmap": "mxmed_envmap"}),
    "mx_medium_vdf_glass.xml": ("render-mx-medium-vdf-glass/scene.xml", {"glossy": "mxmedglass_glossy"}),
    "mx_anisotropic_vdf.xml": ("render-mx-anisotropic-vdf/scene.xml",
                               {"anisotropic": "mxaniso_anisotropic", "envmap": "mxmed_envmap"}),
    "bumptest.xml": ("render-bumptest/bumptest.xml", {"glass": "bump_glass", "metal": "bump_metal"}),
    "displacement.xml": "render-displacement/scene.xml",
    "raytypes.xml": ("render-raytypes/scene.xml", {"glossy": "raytypes_glossy", "magic": "raytypes_magic"}),
    "spi_thinlayer.xml": ("render-spi-thinlayer/scene.xml", {"glossy_glass": "spithin_glossy_glass", "envmap": "mf_envmap"}),
}
# input images read by texture() (test input data, copied byte for byte)
TEXTURES = {"kitchen_probe.hdr": "common/textures/kitchen_probe.hdr"}
# golden renders (half-float EXR in the reference; stored as float16 npz)
RENDERS = {
    "render-cornell": "render-cornell/ref/out.exr",
    "render-bunny": "render-bunny/ref/out.exr",
    "render-veachmis": "render-veachmis/ref/out.exr",
    "render-ward": "render-ward/ref/out.exr",
    "render-furnace-diffuse": "render-furnace-diffuse/ref/out.exr",
    "render-oren-nayar": "render-oren-nayar/ref/out.exr",
    "render-mx-furnace-oren-nayar": "render-mx-furnace-oren-nayar/ref/out.exr",
    "render-mx-furnace-burley-diffuse": "render-mx-furnace-burley-diffuse/ref/out.exr",
    "render-mx-layer": "render-mx-layer/ref/out.exr",
    "render-microfacet": "render-microfacet/ref/out.exr",
    "render-mx-conductor": "render-mx-conductor/ref/out.exr",
    "render-mx-dielectric": "render-mx-dielectric/ref/out.exr",
    "render-mx-generalized-schlick": "render-mx-generalized-schlick/ref/out.exr",
    "render-mx-sheen": "render-mx-sheen/ref/out.exr",
    "render-mx-furnace-sheen": "render-mx-furnace-sheen/ref/out.exr",
    "render-mx-burley-diffuse": "render-mx-burley-diffuse/ref/out.exr",
    "render-mx-dielectric-glass": "render-mx-dielectric-glass/ref/out.exr",
    "render-mx-generalized-schlick-glass": "render-mx-generalized-schlick-glass/ref/out.exr",
    "render-mx-medium-vdf": "render-mx-medium-vdf/ref/out.exr",
    "render-mx-medium-vdf-glass": "render-mx-medium-vdf-glass/ref/out.exr",
    "render-mx-anisotropic-vdf": "render-mx-anisotropic-vdf/ref/out.exr",
    "render-spi-thinlayer": "render-spi-thinlayer/ref/out.exr",
    "render-bumptest": "render-bumptest/ref/out.exr",
    "render-raytypes": "render-raytypes/ref/out.exr",
    "render-displacement": "render-displacement/ref/out.exr",
}
IMAGES = {
    # golden name: testsuite-relative image
    "noise": "noise/ref/out.tif",
    "cellnoise": "cellnoise/ref/out.tif",
    "hashnoise": "hashnoise/ref/out.tif",
    "noise-cell": "noise-cell/ref/out.tif",
    "noise-perlin": "noise-perlin/ref/out.tif",
    "noise-simplex": "noise-simplex/ref/out.tif",
    "noise-gabor": "noise-gabor/ref/out.tif",
    "noise-gabor2d-filter": "noise-gabor2d-filter/ref/out.tif",
    "noise-gabor3d-filter": "noise-gabor3d-filter/ref/out.tif",
    "pnoise-gabor": "pnoise-gabor/ref/out.tif",
    "spline-color": "spline/ref/color.tif",
    "spline-dcolor": "spline/ref/dcolor.tif",
    "spline-float": "spline/ref/float.tif",
    "spline-dfloat": "spline/ref/dfloat.tif",
    "spline-numknots": "spline/ref/numknots.tif",
    "pnoise": "pnoise/ref/out.tif",
    "pnoise-cell": "pnoise-cell/ref/out.tif",
    "pnoise-perlin": "pnoise-perlin/ref/out.tif",
    "userdata-partial": "userdata-partial/ref/out.lazy_userdata_ON.tif",
    "noise-generic": "noise-generic/ref/out.tif",
    "pnoise-generic": "pnoise-generic/ref/out.tif",
}
TEXTS = {
    "hash": "hash/ref/out.txt",
    "layers-lazy": "layers-lazy/ref/out.txt",
    "layers": "layers/ref/out.txt",
    "color": "color/ref/out.txt",
    "matrix": "matrix/ref/out.txt",
    "transform": "transform/ref/out.txt",
    "transformc": "transformc/ref/out.txt",
    "error-dupes": "error-dupes/ref/out.txt",
    "userdata-custom": "userdata-custom/ref/out.txt",
}
# testsuite directories whose run.py is a single `testshade [-g X Y] [-center] test` with a text
# golden: (grid x, grid y, center).  Fixtures: oso/ts_<dir>.oso, text/ts_<dir>.txt.
# tests/helpers.py TESTSUITE_TEXT mirrors this table.
TESTSUITE_TEXT = {
    "arithmetic": (1, 1, 0), "array-derivs": (2, 2, 0), "blendmath": (1, 1, 0), "breakcont": (1, 1, 0),
    "bug-locallifetime": (1, 1, 0), "bug-peep": (2, 2, 0), "comparison": (1, 1, 0),
    "const-array-fill": (1, 1, 0), "const-array-params": (1, 1, 0), "derivs": (2, 2, 0),
    "derivs-muldiv-clobber": (1, 1, 0), "exit": (1, 1, 0), "exponential": (1, 1, 0),
    "function-earlyreturn": (2, 2, 0), "function-outputelem": (2, 2, 0), "function-simple": (1, 1, 0),
    "geomath": (2, 2, 0), "hex": (1, 1, 0), "hyperb": (1, 1, 0), "ieee_fp": (1, 1, 0), "incdec": (1, 1, 0),
    "intbits": (1, 1, 0), "logic": (1, 1, 0), "loop": (2, 2, 0), "miscmath": (2, 2, 0),
    "named-components": (1, 1, 0), "oslc-literalfold": (1, 1, 0), "pragma-nowarn": (1, 1, 0),
    "printf-whole-array": (1, 1, 0), "select": (2, 2, 0), "shortcircuit": (2, 2, 0),
    "spline-boundarybug": (1, 1, 0), "splineinverse": (3, 1, 1), "ternary": (2, 2, 0),
    "transitive-assign": (1, 1, 0), "trig": (1, 1, 0), "typecast": (2, 2, 0), "userdata-defaults": (1, 1, 0),
    "userdata": (2, 2, 0),
    "vecctr": (1, 1, 0), "vector": (1, 1, 0),
}
# float / half EXR goldens of testshade image tests (stored as float32 npz, full size)
EXR_IMAGES = {
    "blackbody": "blackbody/ref/out.exr",
    "wavelength_color": "wavelength_color/ref/out.exr",
}


def parse_noise_vectors():
    """Pull the results_Nd / vresults_Nd tables of test_perlin/test_cell/test_hash."""
    src = open(os.path.join(REF, "src/liboslnoise/oslnoise_test.cpp")).read()
    out = {}
    for fn in ("test_perlin", "test_cell", "test_hash"):
        a = src.index("\n" + fn + "()")
        b = src.index("for (int i", a)
        body = src[a:b]
        tables = {}
        for m in re.finditer(r"static\s+(float|Vec3)\s+(\w+)\[[^\]]*\]\s*=\s*\{(.*?)\};", body, re.S):
            kind, name, vals = m.group(1), m.group(2), m.group(3)
            if kind == "float":
                tables[name] = [float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", vals)]
            else:
                tables[name] = [[float(x) for x in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?", v)]
                                for v in re.findall(r"Vec3\(([^)]*)\)", vals)]
        out[fn] = tables
    return out


def main():
    from PIL import Image
    os.makedirs(os.path.join(OUT, "oso"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "images"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "text"), exist_ok=True)
    inc = [os.path.join(REF, "src/shaders")]
    for name, rel in SHADERS.items():
        src = os.path.join(ROOT, rel[5:]) if rel.startswith("repo:") else os.path.join(TS, rel)
        oso = mini_oslc.compile_osl(src, inc)
        with open(os.path.join(OUT, "oso", name + ".oso"), "w") as f:
            f.write(oso)
    for name, rel in IMAGES.items():
        img = np.array(Image.open(os.path.join(TS, rel)))
        if img.ndim == 2:
            img = img[..., None]
        np.savez_compressed(os.path.join(OUT, "images", name + ".npz"),
                            pixels=img[::STEP, ::STEP, :3].copy(), step=STEP,
                            shape=np.array(img.shape[:2]))
    for d in TESTSUITE_TEXT:
        oso = mini_oslc.compile_osl(os.path.join(TS, d, "test.osl"), inc)
        with open(os.path.join(OUT, "oso", "ts_" + d + ".oso"), "w") as f:
            f.write(oso)
        with open(os.path.join(TS, d, "ref", "out.txt")) as f, open(os.path.join(OUT, "text", "ts_" + d + ".txt"), "w") as o:
            o.write(f.read())
    for name, rel in TEXTS.items():
        with open(os.path.join(TS, rel)) as f, open(os.path.join(OUT, "text", name + ".txt"), "w") as o:
            o.write(f.read())
    os.makedirs(os.path.join(OUT, "scenes"), exist_ok=True)
    for name, rel in SCENES.items():
        rename = {}
        if isinstance(rel, tuple):
            rel, rename = rel
        with open(os.path.join(TS, rel), "rb") as f, open(os.path.join(OUT, "scenes", name), "wb") as o:
            data = f.read()
            for old, new in rename.items():
                data = re.sub((r"\bshader\s+%s\b" % re.escape(old)).encode(), ("shader " + new).encode(), data)
            data = data.replace(b"../common/textures/", b"../textures/")
            o.write(data)
    os.makedirs(os.path.join(OUT, "textures"), exist_ok=True)
    for name, rel in TEXTURES.items():
        with open(os.path.join(TS, rel), "rb") as f, open(os.path.join(OUT, "textures", name), "wb") as o:
            o.write(f.read())
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    for name, rel in RENDERS.items():
        img = cv2.imread(os.path.join(TS, rel), cv2.IMREAD_UNCHANGED)[..., ::-1][..., :3]
        np.savez_compressed(os.path.join(OUT, "images", name + ".npz"), pixels=img.astype(np.float16))
    for name, rel in EXR_IMAGES.items():
        img = cv2.imread(os.path.join(TS, rel), cv2.IMREAD_UNCHANGED)[..., ::-1][..., :3]
        np.savez_compressed(os.path.join(OUT, "images", name + ".npz"), pixels=img.astype(np.float32))
    with open(os.path.join(OUT, "noise_vectors.json"), "w") as f:
        json.dump(parse_noise_vectors(), f, indent=1)
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
