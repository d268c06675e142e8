#!/usr/bin/env python3
"""Sweep group-kernel options on a bench workload (device-resident points/s per option string).

  python tools/group_tune.py noise-1024 "" "minblocks=5" "block=128,minblocks=10" ...
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import bench
    import openshadinglanguage_b200 as ob
    name = sys.argv[1]
    w = bench.workload(name)
    dev = torch.device("cuda", 0)
    res = w["res"]
    n = res * res
    var, uni = ob.grid_globals(res, res, **w["globals"])
    for o in sys.argv[2:] or [""]:
        g = ob.ShaderGroup(w["layers"], w["conns"], w["outputs"], options="fma=1," + o)
        dvar = {k: torch.from_numpy(v).to(dev) for k, v in var.items() if g.reads_global(k)}
        dout = torch.zeros((n, w["out_floats"]), dtype=torch.float32, device=dev)
        launch = g.bind(n, dvar, uni, dout)
        steps = 200 if n <= (1 << 21) else 30
        for _ in range(10):
            launch()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                launch()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        print(json.dumps({"workload": name, "options": o, "us_per_step": round(best * 1e3, 2),
                          "Gpts_s": round(n / best / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
