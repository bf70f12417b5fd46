"""bf16 tensor-path MLP forward vs the fp32 precise path on the same device (quick numerics probe)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from oracle import nerf_oracle as orc
from simplenerf_b200 import ops, synthetic
from simplenerf_b200._lib import FLAG_PRECISE, FLAG_SAVE_FOR_BWD
from simplenerf_b200.models.FusedSimpleNeRF01 import MlpBlock

DEV = 'cuda:0'
configs = synthetic.make_configs('simplenerf')
save = '--save' in sys.argv
for n_rays, s in ((3, 7), (100, 64), (2048, 64), (1500, 192)):
    b = synthetic.make_ray_batch('llff', n_rays, 3)
    o, d, vd = b['rays_o_ndc'].to(DEV), b['rays_d_ndc'].to(DEV), b['view_dirs'].to(DEV)
    z = torch.sort(torch.rand(n_rays, s, device=DEV), -1)[0].contiguous()
    noise = torch.randn(n_rays * s, device=DEV)
    for slot, cfg in orc.model_slots(configs).items():
        spec = orc.MlpSpec(cfg)
        state = orc.deterministic_state(spec.param_shapes(), 100 + len(slot))
        state = {k: (v * 2 if 'weight' in k else v) for k, v in state.items()}
        block = MlpBlock(cfg); block.load_state_dict(state); block.to(DEV)
        table = [None if p is None else p.detach() for p in block.param_table()]
        res = {}
        for prec in ('fp32', 'bf16'):
            flags = (FLAG_PRECISE if prec == 'fp32' else 0) | (FLAG_SAVE_FOR_BWD if save else 0)
            packed = None if prec == 'fp32' else block.packed(table)
            ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n_rays, s, flags), dtype=torch.uint8, device=DEV)
            torch.cuda.synchronize(); t0 = time.time()
            res[prec] = ops.mlp_forward(block.desc, table, packed, o, d, vd, z, noise, ws, flags)
            torch.cuda.synchronize(); dt = time.time() - t0
            res[prec + '_t'] = dt
        es = (res['fp32'][0] - res['bf16'][0]).abs().max().item()
        er = (res['fp32'][1] - res['bf16'][1]).abs().max().item()
        print(f'{n_rays}x{s} {slot}: sigma max {res["fp32"][0].max().item():.3f} err {es:.2e} | rgb err {er:.2e} | '
              f't fp32 {res["fp32_t"]*1e3:.2f} ms bf16 {res["bf16_t"]*1e3:.2f} ms', flush=True)
