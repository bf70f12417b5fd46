"""Gradient parity of the drop-in (bf16 tensor path) against the fp32 CPU oracle under a coherent training loss."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import golden_util as gu
from oracle import nerf_oracle as orc
from oracle.bf16_emulation import mlp_forward_bf16
from simplenerf_b200 import synthetic
from simplenerf_b200.models import get_model
from simplenerf_b200.models.FusedSimpleNeRF01 import FixedRandoms

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dense = '--raw' not in sys.argv
configs = synthetic.make_configs('simplenerf')
state = gu.full_state(configs, 7, dense=dense)
batch = synthetic.make_ray_batch('llff', n, 1021)
g = torch.Generator().manual_seed(5)
table = {'t_rand': torch.rand((n, 64), generator=g), 'u': torch.rand((n, 128), generator=g)}
for slot in orc.model_slots(configs):
    table[f'noise_{slot}'] = torch.randn((n * (192 if 'fine' in slot else 64), 1), generator=g)
target = torch.rand((n, 3), generator=g)
tdepth = 1 + 4 * torch.rand((n,), generator=g)
KEYS = [('rgb_coarse', 'depth_coarse'), ('rgb_fine', 'depth_fine'), ('points_augmentation_rgb_coarse', 'points_augmentation_depth_coarse'),
        ('views_augmentation_rgb_coarse', 'views_augmentation_depth_coarse')]


def loss_of(out, dev):
    t, d = target.to(dev), tdepth.to(dev)
    return sum(((out[a] - t) ** 2).mean() + 0.1 * ((out[b] - d) ** 2).mean() for a, b in KEYS)


grads = {}
for tag, impl in (('fp32_oracle', orc.mlp_forward), ('bf16_emulated', mlp_forward_bf16)):
    o = orc.NerfOracle(configs); o.load_state_dict(state); o.randoms = orc.FixedRandoms(table); o.mlp_impl = impl; o.train()
    t0 = time.time(); out = o(batch); loss_of(out, 'cpu').backward()
    grads[tag] = {k: p.grad.clone() for k, p in o.named_parameters()}
    print(tag, f'{time.time()-t0:.1f}s', 'loss', float(loss_of(out, 'cpu')), flush=True)
    if tag == 'fp32_oracle':
        ref_out = {k: v.detach() for k, v in out.items()}
for prec in ('fp32', 'bf16'):
    cfg = dict(configs, model=dict(configs['model'], precision=prec))
    m = get_model(cfg, None); m.load_state_dict(state); m = m.to('cuda:0').train(); m.randoms = FixedRandoms(table)
    out = m({k: (v.to('cuda:0') if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
    loss_of(out, 'cuda:0').backward(); torch.cuda.synchronize()
    grads[prec] = {k: p.grad.cpu() for k, p in m.named_parameters()}
    for a, b in KEYS:
        print(f'  {prec} {a}: max abs err {float((out[a].detach().cpu()-ref_out[a]).abs().max()):.2e}  {b}: {float((out[b].detach().cpu()-ref_out[b]).abs().max()):.2e}')
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-20))
worst = {}
for k in grads['fp32_oracle']:
    r = grads['fp32_oracle'][k]
    row = (rel(grads['fp32'][k], r), rel(grads['bf16'][k], r), rel(grads['bf16'][k], grads['bf16_emulated'][k]), rel(grads['bf16_emulated'][k], r))
    print(f'{k:52s} |g| {float(r.norm()):.2e} fp32 {row[0]:.1e} bf16 {row[1]:.1e} bf16-vs-emul {row[2]:.1e} emul-vs-fp32 {row[3]:.1e}')
    for i, t in enumerate(('fp32', 'bf16', 'bf16_vs_emul', 'emul_vs_fp32')):
        worst[t] = max(worst.get(t, 0), row[i])
print('WORST', worst)
flat = lambda d: torch.cat([v.flatten() for v in d.values()])
print('whole-gradient rel err: fp32 %.2e bf16 %.2e emul %.2e' % (rel(flat(grads['fp32']), flat(grads['fp32_oracle'])), rel(flat(grads['bf16']), flat(grads['fp32_oracle'])), rel(flat(grads['bf16_emulated']), flat(grads['fp32_oracle']))))
