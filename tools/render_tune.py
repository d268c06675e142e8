#!/usr/bin/env python3
"""Sweep integrator options on one scene (device paths/s per option string).

  python tools/render_tune.py cornell.xml --res 1024 --aa 8 "slots=1048576" "slots=4194304,chunk=24" ...
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--aa", type=int, default=8)
    ap.add_argument("--repeat", type=int, default=3)
    a, extra = ap.parse_known_args()
    a.options = extra or [""]
    import helpers
    from openshadinglanguage_b200 import api
    from openshadinglanguage_b200.render import scene as sc
    S = sc.load_scene(os.path.join(helpers.GOLDEN, "scenes", a.scene))
    A = S.prepare()
    for o in a.options:
        R = api.Renderer(S, A, helpers.oso, a.res, a.res, a.aa, options="fma=1," + o)
        best = None
        for _ in range(a.repeat):
            img = R.render()
            if best is None or R.stats["device_ms"] < best["device_ms"]:
                best = dict(R.stats)
        print(json.dumps({"scene": a.scene, "options": o, "Mpaths_s": round(best["paths"] / best["device_ms"] / 1e3, 1),
                          "device_ms": round(best["device_ms"], 2), "tail_ms": round(best["tail_ms"], 2),
                          "steps": best["bounce_iterations"], "launches": best["launches"], "mean": float(img.mean())}),
              flush=True)
        del R


if __name__ == "__main__":
    main()
