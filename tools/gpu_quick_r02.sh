#!/bin/bash
# quick GPU check of a traversal / lobe change: a subset of the render parity tests + render timings
set -u
O=gpurun_out
mkdir -p $O
TAG=${1:-r02q}
timeout 900 python -m pytest tests/test_render.py -m gpu -x -q -k "${2:-thinlayer or cornell or bunny or mx-layer or dielectric-glass or tiles}" > $O/${TAG}_pytest.log 2>&1
tail -5 $O/${TAG}_pytest.log
for sc in cornell.xml:1024:8 mx_layer.xml:2048:6 render_microfacet.xml:1024:8; do
  IFS=: read f res aa <<< "$sc"
  timeout 300 python tools/render_bench.py $f --res $res --aa $aa --repeat 3 >> $O/${TAG}_render.jsonl 2>> $O/${TAG}_render.err
done
cat $O/${TAG}_render.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['scene'], d['res'], d['aa'], '%.1f Mpaths/s' % (d['paths_per_s_device'] / 1e6), 'tail_ms', d.get('tail_ms'), 'steps', d.get('bounce_iterations'))
"
