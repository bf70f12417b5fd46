#!/bin/bash
# round 2: pipelined stash writers (two bulk stores in flight) -- parity, then timing
T="tests/test_parity_gpu.py tests/test_parity_sizes_gpu.py"
python -m pytest $T -m gpu -q -x > gpurun_out/r2_gputests_e.log 2>&1; tail -n 3 gpurun_out/r2_gputests_e.log
SNERF_BWD_RING=24 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "mlp_backward or reproducible or gradient_parity" > gpurun_out/r2_gputests_e2.log 2>&1; tail -n 2 gpurun_out/r2_gputests_e2.log
for i in 1 2; do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-render --no-c5 --no-trainer > gpurun_out/r2_bench_b$i.json 2> gpurun_out/r2_bench_b$i.err
  python - gpurun_out/r2_bench_b$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]),round(d["ms_per_step"],3),{k:round(v,3) for k,v in d["roofline"]["ms_per_step"].items()},{k:round(v["ms_per_step"],3) for k,v in d["roofline"]["kernels"].items()},d["clocks"]["sm_mhz"],d["clocks"]["reasons"])
PY
done
