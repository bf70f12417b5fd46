#!/bin/bash
# Developer A/B of one build on ONE box: an environment switch off / on, alternating short bench runs.
# Usage (under gpurun): tools/ab_env.sh SNERF_FWD_DIRECT [tag]
VAR=$1
TAG=${2:-abenv}
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-c5 --no-trainer --no-render"
for i in 1 2 3; do
  env $VAR=0 $B > gpurun_out/${TAG}_off$i.json 2> gpurun_out/${TAG}_off$i.err
  env $VAR=1 $B > gpurun_out/${TAG}_on$i.json 2> gpurun_out/${TAG}_on$i.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_o*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), {k:round(v,3) for k,v in r['ms_per_step'].items()}, {k[3:-7]:round(v['ms_per_step'],3) for k,v in r['kernels'].items()}, d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'failed', e)
PY
