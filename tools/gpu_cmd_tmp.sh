set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_render.py -m gpu -x -q -k "cornell or bunny or tiles or veachmis or ward or glass or own_scene or thinlayer" > $O/r02g_pytest.log 2>&1; tail -3 $O/r02g_pytest.log
timeout 600 python tools/render_tune.py cornell.xml --res 1024 --aa 8 "sort=1" "sort=1,refill=12" > $O/r02g_tune.jsonl 2>$O/r02g_tune.err
timeout 300 python tools/render_tune.py bunny.xml --res 1024 --aa 8 "sort=1" >> $O/r02g_tune.jsonl 2>>$O/r02g_tune.err
timeout 300 python tools/render_tune.py mx_layer.xml --res 2048 --aa 6 "sort=1" >> $O/r02g_tune.jsonl 2>>$O/r02g_tune.err
timeout 300 python tools/render_tune.py render_microfacet.xml --res 1024 --aa 8 --repeat 1 "sort=1" >> $O/r02g_tune.jsonl 2>>$O/r02g_tune.err
cut -c1-150 $O/r02g_tune.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rt_trace -s 12 -c 1 -o $O/r02g_full_rt_trace python tools/render_bench.py cornell.xml --res 1024 --aa 8 --repeat 1 > $O/r02g_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 600 --csv --log-file $O/r02g_launches_mx_layer.csv python tools/render_bench.py mx_layer.xml --res 1024 --aa 6 --repeat 1 > $O/r02g_ncu2.log 2>&1
