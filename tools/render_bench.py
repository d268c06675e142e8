#!/usr/bin/env python3
"""Time the wavefront path tracer on a testrender scene (paths/s).

  python tools/render_bench.py cornell.xml --res 1024 --aa 8 [--options fma=1,sort=1] [--cpu]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--aa", type=int, default=8)
    ap.add_argument("--options", default="fma=1,sort=1")
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle on a bounded sample")
    a = ap.parse_args()
    import helpers
    from openshadinglanguage_b200 import api
    from openshadinglanguage_b200.render import scene as sc
    path = a.scene if os.path.exists(a.scene) else os.path.join(helpers.GOLDEN, "scenes", a.scene)
    S = sc.load_scene(path)
    A = S.prepare()
    R = api.Renderer(S, A, helpers.oso, a.res, a.res, a.aa, options=a.options)
    out = {"scene": os.path.basename(path), "res": a.res, "aa": a.aa, "options": a.options,
           "tris": int(len(A["triangles"]))}
    best = None
    for _ in range(a.repeat):
        t0 = time.perf_counter()
        img = R.render()
        wall = time.perf_counter() - t0
        st = R.stats
        if best is None or st["device_ms"] < best["device_ms"]:
            best = dict(st, wall_s=wall)
    out.update(best)
    out["paths_per_s_device"] = best["paths"] / (best["device_ms"] * 1e-3)
    out["paths_per_s_wall"] = best["paths"] / best["wall_s"]
    out["mean"] = float(img.mean())
    if a.cpu:
        from oracle import oracle
        rows = max(8, min(a.res, 64))
        O = oracle.OracleRender(S, A, helpers.oso)
        # bounded sample: a centred band of `rows` rows is not expressible in the oracle API,
        # so render a lower-resolution frame with the same spp (same work per path)
        t0 = time.perf_counter()
        O.render(rows * 2, rows * 2, a.aa)
        dt = time.perf_counter() - t0
        out["cpu_paths_per_s"] = (rows * 2) ** 2 * a.aa * a.aa / dt
        out["cpu_threads"] = os.cpu_count()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
