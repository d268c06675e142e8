set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_render.py -m gpu -x -q -k "microfacet or furnace-diffuse or mx-layer or thinlayer" > $O/r02h_pytest.log 2>&1; tail -3 $O/r02h_pytest.log
timeout 300 python tools/render_tune.py render_microfacet.xml --res 1024 --aa 8 --repeat 1 "sort=1" > $O/r02h_tune.jsonl 2>$O/r02h_tune.err
timeout 300 python tools/render_tune.py mx_layer.xml --res 2048 --aa 6 "sort=1" >> $O/r02h_tune.jsonl 2>>$O/r02h_tune.err
cut -c1-160 $O/r02h_tune.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rt_tail -c 1 -o $O/r02h_full_rt_tail python tools/render_bench.py render_microfacet.xml --res 384 --aa 4 --repeat 1 > $O/r02h_ncu.log 2>&1
tail -2 $O/r02h_ncu.log | cut -c1-300
