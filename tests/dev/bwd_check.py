"""Per-parameter gradient error of the bf16 tensor path and the fp32 precise path vs CPU autograd."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from oracle import nerf_oracle as orc
from simplenerf_b200 import ops, synthetic
from simplenerf_b200._lib import FLAG_PRECISE, FLAG_SAVE_FOR_BWD
from simplenerf_b200.models.FusedSimpleNeRF01 import MlpBlock
DEV = 'cuda:0'
from oracle.bf16_emulation import mlp_forward_bf16


def mlp_bf16_emulated(spec, P, pts, vd, noise):
    out = mlp_forward_bf16(spec, P, pts, vd, noise)
    return out["sigma"], out["rgb"]


configs = synthetic.make_configs('simplenerf')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
gen = torch.Generator().manual_seed(2)
pts = (torch.rand((n, 3), generator=gen) - .5) * 2.4
vd = torch.nn.functional.normalize(torch.randn((n, 3), generator=gen), dim=-1)
noise = torch.randn((n, 1), generator=gen)
c_s, c_r = torch.randn((n, 1), generator=gen), torch.randn((n, 3), generator=gen)
for slot, cfg in orc.model_slots(configs).items():
    if slot == 'fine_model':
        continue
    spec = orc.MlpSpec(cfg)
    state = orc.deterministic_state(spec.param_shapes(), 100 + len(slot))
    params = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    out = orc.mlp_forward(spec, params, pts, vd, noise)
    ((out['sigma'] * c_s).sum() + (out['rgb'] * c_r).sum()).backward()
    params_e = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    se, re_ = mlp_bf16_emulated(spec, params_e, pts, vd, noise)
    ((se * c_s).sum() + (re_ * c_r).sum()).backward()
    block = MlpBlock(cfg); block.load_state_dict(state); block.to(DEV)
    table = [None if p is None else p.detach() for p in block.param_table()]
    lookup = {id(p): k for k, p in block.named_parameters()}
    res = {}
    for prec in ('fp32', 'bf16'):
        flags = (FLAG_PRECISE if prec == 'fp32' else 0) | FLAG_SAVE_FOR_BWD
        packed = None if prec == 'fp32' else block.packed(table)
        ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n, 1, flags), dtype=torch.uint8, device=DEV)
        z = torch.zeros((n, 1), device=DEV); o = pts.to(DEV); d0 = torch.zeros((n, 3), device=DEV)
        sigma, rgb = ops.mlp_forward(block.desc, table, packed, o, d0, vd.to(DEV), z, noise.reshape(-1).to(DEV), ws, flags)
        grads = [None if p is None else torch.zeros_like(p) for p in table]
        ops.mlp_backward(block.desc, table, packed, o, d0, vd.to(DEV), z, sigma, rgb, c_s.reshape(n, 1).to(DEV),
                         c_r.reshape(n, 1, 3).to(DEV), grads, ws, flags)
        torch.cuda.synchronize()
        res[prec] = grads
    print(f'== {slot} (n={n})')
    for i, p in enumerate(block.param_table()):
        if p is None:
            continue
        name = lookup[id(p)]
        want = params[name].grad
        e32 = float((res['fp32'][i].cpu() - want).norm() / (want.norm() + 1e-12))
        e16 = float((res['bf16'][i].cpu() - want).norm() / (want.norm() + 1e-12))
        wante = params_e[name].grad
        eem = float((res['bf16'][i].cpu() - wante).norm() / (wante.norm() + 1e-12))
        eref = float((wante - want).norm() / (want.norm() + 1e-12))
        print(f'  {name:28s} |g| {float(want.norm()):9.3e}  rel err fp32 {e32:.2e}  bf16 {e16:.2e} | bf16 vs emulated {eem:.2e} (emulated vs fp32 {eref:.2e})', flush=True)
