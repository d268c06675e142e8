#!/usr/bin/env python3
"""Summarise ncu output brought back in gpurun_out/ (run here: no GPU needed).

  python tools/ncu_summary.py shares gpurun_out/launches.csv          # time share per kernel
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [...]     # key metrics of --set full captures
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("registers_per_thread", "launch__registers_per_thread"),
    ("dynamic_smem_kb_per_block", "launch__shared_mem_per_block_dynamic"),
    ("occupancy_limit_blocks_registers", "launch__occupancy_limit_registers"),
    ("occupancy_limit_blocks_smem", "launch__occupancy_limit_shared_mem"),
    ("theoretical_occupancy_pct", "sm__maximum_warps_per_active_cycle_pct"),
    ("achieved_occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("active_lanes_per_instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("issue_slots_busy_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("pipe_fma_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("pipe_alu_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
    ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_hit_rate_pct", "lts__t_sector_hit_rate.pct"),
    ("local_load_instructions", "smsp__sass_inst_executed_op_local_ld.sum"),
    ("local_store_instructions", "smsp__sass_inst_executed_op_local_st.sum"),
    ("warp_instructions", "smsp__inst_executed.sum"),
    ("branch_efficiency_pct", "smsp__sass_average_branch_targets_threads_uniform.pct"),
]


def shares(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        tot[r[ki]] += v
        cnt[r[ki]] += 1
    T = sum(tot.values())
    print("%s: %d launches, %.1f us in total (per-launch times are cold-cache and serialised: shares only)"
          % (path, sum(cnt.values()), T))
    for k, v in tot.most_common():
        print("  %-34s n=%5d  %11.1f us  %5.1f %%  (%.1f us each)" % (k[:34], cnt[k], v, 100 * v / T, v / cnt[k]))


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for name, key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                try:
                    v = float(vals[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if name.endswith("_MB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if name == "duration_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                d[name] = round(v, 3)
        out.append(d)
    return out


if __name__ == "__main__":
    if sys.argv[1] == "shares":
        for p in sys.argv[2:]:
            shares(p)
    else:
        res = {}
        for p in sys.argv[2:]:
            res[p] = full(p)
        print(json.dumps(res, indent=1))
