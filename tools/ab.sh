#!/bin/bash
# Developer A/B on ONE box (box-to-box spread of the step time is ~4 %): the committed build (tools/ab/base.so, made by hand from
# a stash of the working tree) against the working tree's build, alternating, short bench runs.  Usage (under gpurun): tools/ab.sh [tag]
TAG=${1:-ab}
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-c5 --no-trainer --no-render"
for i in 1 2 3; do
  SNERF_B200_LIB_AB=$PWD/tools/ab/base.so $B > gpurun_out/${TAG}_base$i.json 2> gpurun_out/${TAG}_base$i.err
  $B > gpurun_out/${TAG}_new$i.json 2> gpurun_out/${TAG}_new$i.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${TAG}_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); r=d['roofline']
        print(f, round(d['ms_per_step'],3), {k:round(v,3) for k,v in r['ms_per_step'].items()}, {k[3:-7]:round(v['ms_per_step'],3) for k,v in r['kernels'].items()})
    except Exception as e: print(f, 'failed', e)
PY
