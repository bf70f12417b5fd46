#!/bin/bash
# round 2: fused backward (dgrad pairs + wgrad CTAs in one launch, gradient ring in L2) -- correctness, then A/B timing
set -x
mkdir -p gpurun_out
T="tests/test_parity_gpu.py"
K="mlp_backward_vs_autograd or tensor_path_reproducible or training_gradient_parity or full_size_invariants or tile_edges"
timeout 600 python -m pytest $T -x -q -m gpu -k "$K" > gpurun_out/r2_t_fused.log 2>&1; echo "fused rc=$?" >> gpurun_out/r2_t_fused.log
SNERF_BWD_RING=12 timeout 600 python -m pytest $T -x -q -m gpu -k "$K" > gpurun_out/r2_t_ring12.log 2>&1; echo "ring12 rc=$?" >> gpurun_out/r2_t_ring12.log
SNERF_BWD_RING=0 timeout 600 python -m pytest $T -x -q -m gpu -k "$K" > gpurun_out/r2_t_split.log 2>&1; echo "split rc=$?" >> gpurun_out/r2_t_split.log
for cfg in "0 43" "48 43" "96 43" "24 43" "1000000 43" "48 40" "48 46" "48 37" "96 46"; do
  set -- $cfg
  SNERF_BWD_RING=$1 SNERF_BWD_DGRAD_PAIRS=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-render > gpurun_out/r2_b_$1_$2.json 2> gpurun_out/r2_b_$1_$2.err
done
tail -n 3 gpurun_out/r2_t_*.log
for f in gpurun_out/r2_b_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['roofline']['ms_per_step'], d['clocks'])
except Exception as e:
    print('ERR', e)
PY
done
