#!/usr/bin/env python3
"""bench.py — shaded points/s of the B200 back end on BASELINE.json's configs.

One "step" = one pass of the hot path (execute a compiled ShaderGroup over the
whole synthetic testshade grid).  Headline workload at N=1 is BASELINE.json
configs[1]: the 3-layer `layers` group on a 4096x4096 grid with varying
derivatives and renderer outputs with derivs.

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N ...            # reference CPU arm

`value`   : device-resident throughput (inputs already in HBM), CUDA events.
`e2e`     : same metric through the host-pointer C-ABI call
            (b200_group_execute_host) with pinned host buffers: H2D of the
            planes the group reads + kernel + D2H of the output arena, per step.
`roofline`: achieved HBM GB/s of the group kernel = algorithmic bytes
            (72 B/pt: 24 B read + 48 B written) / event-timed
            launch duration, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline` / --impl reference: the restated reference algorithm (oracle
            port, compiled C++, one execute per point like testshade) on the
            host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "shaded points/sec (testshade grid)"
UNIT = "points/s"


def workload(name):
    import helpers
    if name == "layers-4096":
        layers, conns, outputs = helpers.layers_group(derivs=True)
        return dict(layers=layers, conns=conns, outputs=outputs, res=4096, out_floats=12,
                    globals=dict(vary_udxdy=True, vary_vdxdy=True, vary_pdxdy=True),
                    bytes_per_point=72, desc="testsuite/layers-lazy a,b,c 3-layer group, 4096x4096, "
                    "varying derivs, outputs f_out,c_out with derivs")
    if name == "noise-1024":
        layers, outputs, _ = helpers.image_case_group("noise")
        return dict(layers=layers, conns=(), outputs=outputs, res=1024, out_floats=3, globals={},
                    bytes_per_point=20, desc="testsuite/noise/test.osl, 1024x1024, Cout")
    raise SystemExit("unknown workload " + name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                   timeout=5)
                if r.returncode == 0:
                    self.samples.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference(wl, seconds_target=12.0, threads=None):
    """Restated reference algorithm (oracle port) on host cores, bounded sample."""
    import numpy as np
    from oracle import oracle
    threads = threads or os.cpu_count() or 1
    g = oracle.OracleGroup(wl["layers"], wl["conns"], wl["outputs"])
    res = wl["res"]
    rows = max(threads, min(res, 512))          # bounded sample: `rows` grid rows of the same grid
    var, uni = oracle.testshade_globals(res, res, **wl["globals"])
    n = rows * res
    full = res * res
    var = {k: (np.asarray(v).reshape(-1, full)[:, :n].copy() if np.asarray(v).size != full
               else np.asarray(v)[:n].copy()) for k, v in var.items()}
    out = np.zeros((n, wl["out_floats"]), np.float32)
    g.run(n, var, uni, out, nthreads=threads)   # warm
    reps, t0 = 0, time.perf_counter()
    while True:
        g.run(n, var, uni, out, nthreads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt > seconds_target or reps >= 200:
            break
    return dict(value=n * reps / dt, unit=UNIT, cores=threads, kind="port",
                sample="%d rows x %d cols of the %dx%d grid, %d repeats, scalar C++ restatement "
                       "(-O2 -ffp-contract=off), one execute per point, %d host threads"
                       % (rows, res, res, res, reps, threads)), n, g, var, uni, out


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    base, n, g, var, uni, out = cpu_reference(wl, seconds_target=2.0, threads=threads)
    for _ in range(args.warmup):
        g.run(n, var, uni, out, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.run(n, var, uni, out, nthreads=threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"],
                                            "points_per_step": n},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="layers-4096", choices=["layers-4096", "noise-1024"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary noise-1024 numbers")
    ap.add_argument("--config5", action="store_true",
                    help="also render BASELINE config 5 (render-bunny 4096^2, 256 spp; ~25 s on one GPU)")
    args = ap.parse_args()
    wl = workload(args.workload)
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import numpy as np
    import torch
    import torch.distributed as dist
    import openshadinglanguage_b200 as ob

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        # One process per GPU: run (and first-touch the pinned host buffers) on the NUMA node the
        # GPU's PCIe root hangs off, so that host<->device copies do not cross the socket link.
        try:
            bus = torch.cuda.get_device_properties(local).pci_bus_id
            dom = torch.cuda.get_device_properties(local).pci_domain_id
            dev_id = torch.cuda.get_device_properties(local).pci_device_id
            sysdir = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, dev_id)
            node = int(open(sysdir + "/numa_node").read())
            if node >= 0:
                cpus = set()
                for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                cpus &= os.sched_getaffinity(0)
                if cpus:
                    os.sched_setaffinity(0, cpus)
                    numa = node
        except Exception:
            numa = None
    if world > 1:
        # stdout carries exactly one JSON line.  With NCCL_DEBUG >= VERSION in the environment
        # NCCL printf()s its banner to stdout at communicator creation, so fd 1 points at stderr
        # while the communicator comes up.
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def bench_workload(w, steps, warmup, with_e2e=True):
        g = ob.ShaderGroup(w["layers"], w["conns"], w["outputs"], options="fma=1")
        res = w["res"]
        n = res * res                      # per-rank shard: one full grid tile per GPU (weak scaling)
        var, uni = ob.grid_globals(res, res, **w["globals"])
        used = {k: v for k, v in var.items() if g.reads_global(k)}
        dvar = {k: torch.from_numpy(v).to(dev) for k, v in used.items()}
        dout = torch.zeros((n, w["out_floats"]), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev)
        for _ in range(max(warmup, 3)):
            g.execute(n, dvar, uni, dout, device=local)
        barrier()
        c0 = ob.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            g.execute(n, dvar, uni, dout, device=local)
        e1.record(stream)
        barrier()
        launches = ob.launch_count() - c0
        ms = max_over_ranks(e0.elapsed_time(e1))
        res_d = dict(n=n, ms=ms, launches=launches, h2d=0, d2h=0, e2e_ms=None)
        if with_e2e:
            hvar = {k: torch.from_numpy(v).pin_memory() for k, v in used.items()}
            hout = torch.zeros((n, w["out_floats"]), dtype=torch.float32).pin_memory()
            for _ in range(2):
                g.execute_host(n, hvar, uni, hout, device=local)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                g.execute_host(n, hvar, uni, hout, device=local)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                dist.barrier()
            res_d["e2e_ms"] = max_over_ranks(dt * 1e3)
            res_d["h2d"] = int(sum(v.numel() * 4 for v in hvar.values()))
            res_d["d2h"] = int(hout.numel() * 4)
            # sanity: host path result equals device path result
            assert torch.equal(hout, dout.cpu()), "host path and device path disagree"
        return res_d

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    r = bench_workload(wl, args.steps, args.warmup)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=3)
    extra = {}
    if not args.no_extra and args.workload == "layers-4096":
        w2 = workload("noise-1024")
        r2 = bench_workload(w2, max(args.steps, 50), args.warmup, with_e2e=True)
        extra["noise-1024"] = {
            "value": world * r2["n"] * max(args.steps, 50) / (r2["ms"] * 1e-3), "unit": UNIT,
            "ms_per_step": r2["ms"] / max(args.steps, 50),
            "e2e_value": world * r2["n"] * max(args.steps, 50) / (r2["e2e_ms"] * 1e-3),
            "desc": w2["desc"], "bound": "simt (integer hash + fp32 lerps), not HBM"}
    # BASELINE config 3 (render-cornell 1024^2, 64 spp) and the render-mx-layer half of config 4
    # (2048^2, -aa 6: layered MaterialX closures under a procedural sky with background
    # importance sampling): rows sharded over the GPUs, framebuffer strips gathered to rank 0
    # with NCCL (the only collective on this path)
    render_cfgs = [("render-cornell-1024-64spp", "cornell.xml", 1024, 8, 128),
                   ("render-mx-layer-2048-36spp", "mx_layer.xml", 2048, 6, 160),
                   # config 4's other half: microfacet glass under the HDR probe read by texture()
                   ("render-microfacet-2048-64spp", "render_microfacet.xml", 2048, 8, 128)]
    if args.config5:
        render_cfgs.append(("render-bunny-4096-256spp", "bunny.xml", 4096, 16, 96))
    for rname, rxml, res, aa, cpu_res in (render_cfgs if (not args.no_extra and args.workload == "layers-4096") else []):
        try:
            import helpers
            from openshadinglanguage_b200 import api
            from openshadinglanguage_b200.render import scene as rsc
            from openshadinglanguage_b200.sharding import gather_strips
            S = rsc.load_scene(os.path.join(helpers.GOLDEN, "scenes", rxml))
            A = S.prepare()
            R = api.Renderer(S, A, helpers.oso, res, res, aa, options="fma=1,sort=1")
            rows = [(res * k) // world for k in range(world + 1)]
            y0, y1 = rows[rank], rows[rank + 1]
            # warm-up: module load, scene upload and the path-state allocation (sized for the band,
            # GBs of HBM, kept for the renderer's life) - one whole frame, as a renderer's first frame
            R.render(y0, y1, device=local)
            if world > 1:                                     # and the gather's communicator channels
                gather_strips(torch.zeros((res * (y1 - y0), 3), device=dev), res * res, rank, world, align=res)
            barrier()
            t0 = time.perf_counter()
            img = R.render(y0, y1, device=local)
            dt = max_over_ranks(time.perf_counter() - t0)
            dev_ms = max_over_ranks(R.stats["device_ms"])
            gather_ms = 0.0
            if world > 1:
                strip = torch.from_numpy(img.reshape(-1, 3)).to(dev)
                barrier()
                g0 = time.perf_counter()
                gather_strips(strip, res * res, rank, world, align=res)   # bands = whole rows
                torch.cuda.synchronize()
                gather_ms = max_over_ranks((time.perf_counter() - g0) * 1e3)
            paths = res * res * aa * aa
            extra[rname] = {
                "metric": "paths/sec (testrender)", "value": paths / (dev_ms * 1e-3), "unit": "paths/s",
                "e2e_value": paths / dt, "device_ms": dev_ms, "wall_ms": dt * 1e3,
                "framebuffer_gather_ms": gather_ms, "partition": "contiguous row bands per GPU (strong scaling)",
                "bounce_iterations": R.stats["bounce_iterations"], "launches": R.stats["launches"]}
            if rank == 0 and not args.no_cpu_baseline:
                from oracle import oracle as _o
                ncpu = os.cpu_count() or 1
                orender = _o.OracleRender(S, A, helpers.oso)       # compile outside the timed region
                orender.render(16, 16, 1, nthreads=1)
                t0 = time.perf_counter()
                orender.render(cpu_res, cpu_res, aa, nthreads=ncpu)
                extra[rname]["cpu_baseline"] = {
                    "value": cpu_res * cpu_res * aa * aa / (time.perf_counter() - t0), "unit": "paths/s",
                    "cores": ncpu, "kind": "port",
                    "sample": "same scene and spp at %dx%d, scalar C++ restatement, scanline-parallel on %d threads"
                              % (cpu_res, cpu_res, ncpu)}
            del R
        except Exception as e:  # the headline line must still be printed
            extra[rname] = {"error": str(e)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    steps = args.steps
    value = world * r["n"] * steps / (r["ms"] * 1e-3)
    e2e_value = world * r["n"] * steps / (r["e2e_ms"] * 1e-3)
    peak, peak_src = measured_peak()
    kernel_ms = r["ms"] / steps
    achieved = wl["bytes_per_point"] * r["n"] / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(args.warmup, 3), "ms_per_step": kernel_ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "points_per_step_per_gpu": r["n"],
                   "l2_policy": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2"
                                % (wl["bytes_per_point"] * r["n"] / 1e6),
                   "partition": "one full grid tile per GPU, no data-path collective", "fma": 1},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": r["h2d"],
                "d2h_bytes_per_step": r["d2h"], "ms_per_step": r["e2e_ms"] / steps,
                "timer": "host wall clock around the synchronous C-ABI call, max over ranks",
                "host_numa_node_rank0": numa},
        "gpu_launches": int(r["launches"]),
        "clocks": sampler.summary() if sampler else None,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "kernel": "osl_b200_group_kernel",
                     "algorithmic_bytes_per_launch": wl["bytes_per_point"] * r["n"]},
        "other_workloads": extra,
    }
    tr = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tr):
        try:
            line["roofline"]["traffic"] = json.load(open(tr)).get(args.workload)
        except Exception:
            pass
    if not args.no_cpu_baseline:
        base, *_ = cpu_reference(wl, seconds_target=12.0)
        line["cpu_baseline"] = base
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
