"""Row X1 evidence: one launch group (65536 rays, vanilla coarse + fine evaluation) through the drop-in with the samples kept on
chip (snerf_render_forward) and through the two-kernel path (MLP kernel -> sigma / rgb in HBM -> compositing kernel).
Run plain for CUDA-event times, or under `ncu --set full -k "regex:tc_forward|composite_|sample_fine"` for the DRAM bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from simplenerf_b200 import synthetic
from simplenerf_b200.models import get_model

DEV = 'cuda:0'
n = 65536
batch = synthetic.make_ray_batch('llff', n, 0, frame=True, start=200 * 1008)
batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
for fused in (True, False):
    cfg = synthetic.make_configs('vanilla')
    cfg['model']['fused_composite'] = fused
    model = get_model(cfg, None)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synthetic.densify_state(synthetic.deterministic_state(shapes, 0)))
    model = model.to(DEV).eval()
    with torch.no_grad():
        model(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 1 if os.environ.get('X1_ONCE') else 5
        for _ in range(reps):
            out = model(batch)
        e1.record()
        torch.cuda.synchronize()
    print(f"fused={fused}: {e0.elapsed_time(e1) / reps:.3f} ms per 65536-ray launch group; keys {sorted(out)}")
