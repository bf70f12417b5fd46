#!/bin/bash
# Developer sweep of one environment variable on ONE box: tools/ab_vals.sh VAR tag v1 v2 ...   (two rounds, alternating)
VAR=$1; TAG=$2; shift 2
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-c5 --no-trainer --no-render"
for i in 1 2; do
  for v in "$@"; do env $VAR=$v $B > gpurun_out/${TAG}_v${v}_$i.json 2> gpurun_out/${TAG}_v${v}_$i.err; done
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_v*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), {k[3:-7]:round(v['ms_per_step'],3) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'failed', e)
PY
