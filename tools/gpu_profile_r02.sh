#!/bin/bash
# Round-2 GPU session: parity tests, default bench (both arms), ncu launch lists and full captures.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
TAG=${1:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
tail -3 $O/${TAG}_pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 1200 python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
tail -c 600 $O/${TAG}_bench_n1.err
# launch lists (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 1500 --csv --log-file $O/${TAG}_launches_cornell.csv \
    python tools/render_bench.py cornell.xml --res 512 --aa 4 --repeat 1 > $O/${TAG}_ncu_cornell.log 2>&1
# full captures: the group kernels and the three integrator kernels in mid-render
timeout 600 ncu --set full --clock-control none --import-source on -k regex:osl_b200_group_kernel -s 4 -c 1 -o $O/${TAG}_full_layers \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $O/${TAG}_ncu_full_layers.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:osl_b200_group_kernel -s 4 -c 1 -o $O/${TAG}_full_noise \
    python bench.py --workload noise-1024 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $O/${TAG}_ncu_full_noise.log 2>&1
for k in rt_trace rt_shade; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o $O/${TAG}_full_$k \
      python tools/render_bench.py cornell.xml --res 1024 --aa 8 --repeat 1 > $O/${TAG}_ncu_full_$k.log 2>&1
done
ls -la $O | grep ${TAG}
