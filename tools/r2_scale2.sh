#!/bin/bash
# N=2: gradient exchange after the backward (one coalesced launch) against the overlapped per-bucket form
for mode in after overlap after overlap; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --no-render --exchange $mode > gpurun_out/r2_n2_$mode.json 2> gpurun_out/r2_n2_$mode.err
  python - gpurun_out/r2_n2_$mode.json $mode <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d["value"]),round(d["ms_per_step"],3),{k:round(v,3) for k,v in d["roofline"]["ms_per_step"].items()},round(d["host_enqueue_ms_per_step"],2),d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
done
tail -n 5 gpurun_out/r2_n2_after.err
