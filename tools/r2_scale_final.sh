#!/bin/bash
# N = 1, 2, 4, 8 on ONE 8-GPU box, back to back (run under `gpurun --gpus 8`): the driver's own launch form.
TAG=${1:-f4}
A="--steps 20 --warmup 5 --no-cpu --no-c5 --no-trainer"
python bench.py --gpus 1 $A > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N $A > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
done
python - $TAG <<'PY'
import json, sys
tag = sys.argv[1]
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open(f'gpurun_out/{tag}_n{n}.json') if l.startswith('{')][-1])
        base = base or d['value']
        r = d['roofline']['ms_per_step']
        print(n, round(d['value']), round(d['ms_per_step'], 3), round(d['value'] / (n * base), 3), {k: round(v, 3) for k, v in r.items()},
              round(d['e2e']['value']), round(d['host_enqueue_ms_per_step'], 2), round(d['render']['ms_per_frame'], 1), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(n, 'ERR', e)
PY
