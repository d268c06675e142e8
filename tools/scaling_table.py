#!/usr/bin/env python3
"""Markdown scaling table from profiles/bench_r02_n{1,2,4,8}.json (one bench.py line each).

  python tools/scaling_table.py profiles/bench_r02_n{1,2,4,8}.json
"""
import json
import sys


def main():
    runs = {}
    for p in sys.argv[1:]:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        runs[d["n_gpus"]] = d
    ns = sorted(runs)
    base = runs[ns[0]]
    rows = [("layers-4096, device-resident (weak)", lambda d: d["value"], "Gpts/s", 1e9, True),
            ("layers-4096, host buffers e2e (weak)", lambda d: d["e2e"]["value"], "Gpts/s", 1e9, True)]
    for k in base["other_workloads"]:
        if k.startswith("render"):
            rows.append((k + " (strong)", lambda d, k=k: d["other_workloads"].get(k, {}).get("value"), "Mpaths/s", 1e6, False))
            rows.append((k + " e2e incl. gather + D2H", lambda d, k=k: d["other_workloads"].get(k, {}).get("e2e_value"),
                         "Mpaths/s", 1e6, False))
        elif k == "noise-4096":
            rows.append((k + " (weak)", lambda d, k=k: d["other_workloads"][k]["value"], "Gpts/s", 1e9, True))
    print("| workload | unit | " + " | ".join("N=%d" % n for n in ns) + " | speed-up at N=%d | efficiency |" % ns[-1])
    print("|---|---|" + "---|" * (len(ns) + 2))
    for name, get, unit, scale, weak in rows:
        vals = [get(runs[n]) for n in ns]
        if vals[0] is None or vals[-1] is None:
            continue
        sp = vals[-1] / vals[0]
        cells = " | ".join("%.4g" % (v / scale) if v else "-" for v in vals)
        print("| %s | %s | %s | %.2fx | %.2f |" % (name, unit, cells, sp, sp / (ns[-1] / ns[0])))


if __name__ == "__main__":
    main()
