"""Next-row N3: the masked per-ray losses.  CPU: the oracle restatement against fixtures generated from the unmodified
reference loss modules (oracle/make_golden_losses.py), and the host logic that maps the reference's loss modules onto
kernel streams.  GPU: `snerf_ray_losses_*` through the C ABI against the same fixtures."""
import pytest
import torch

import golden_util as gu
from oracle import loss_oracle
from simplenerf_b200 import synthetic
from simplenerf_b200.loss_functions import FusedLossComputer, ray_losses, stream_plan
from simplenerf_b200.loss_functions.FusedLossComputer01 import get_loss_weight

LOSSES = [dict(name='MSE01', weight=1), dict(name='SparseDepthMSE01', weight=0.1), dict(name='MSE02', weight=1),
          dict(name='SparseDepthMSE02', weight=0.1), dict(name='MSE03', weight=1), dict(name='SparseDepthMSE03', weight=0.1)]
# value tolerance: the kernel sums squared errors in a different (fixed) order than torch.mean
RTOL = 2e-6


def _configs():
    return dict(synthetic.make_configs('simplenerf'), losses=[dict(lc) for lc in LOSSES])


def _case(g, tag, device='cpu'):
    inp = {k[len(tag) + 4:]: v.to(device) for k, v in g.items() if k.startswith(f'{tag}_in_')}
    out = {k[len(tag) + 5:]: v.to(device).clone().requires_grad_() for k, v in g.items() if k.startswith(f'{tag}_out_')}
    inp['iter_num'] = 100
    return inp, out


def _streams(configs, inp, out):
    streams = []
    for lc in configs['losses']:
        for pk, tk, mk in stream_plan(configs, lc['name'], inp, out):
            target = inp[tk][:, 0] if tk == 'sparse_depth_values' else inp[tk]
            streams.append((lc['name'], out[pk], target, inp[mk], get_loss_weight(lc, inp['iter_num'])))
    return streams


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_loss_oracle_matches_reference(tag):
    g = gu.load('losses.npz')
    configs = _configs()
    inp, out = _case(g, tag)
    streams = _streams(configs, inp, out)
    assert [s[0] for s in streams] == ['MSE01', 'MSE01', 'SparseDepthMSE01', 'MSE02', 'SparseDepthMSE02', 'MSE03', 'SparseDepthMSE03']
    per_loss = {}
    for name, p, t, m, w in streams:
        per_loss[name] = per_loss.get(name, 0) + loss_oracle.masked_mse(p, t, m)
    for name, v in per_loss.items():
        torch.testing.assert_close(v.detach(), g[f'{tag}_loss_{name}'], rtol=1e-6, atol=0)
    total = loss_oracle.total_loss([s[1:] for s in streams])
    torch.testing.assert_close(total.detach(), g[f'{tag}_loss_TotalLoss'], rtol=1e-6, atol=0)
    total.backward()
    for k, v in out.items():
        want = g[f'{tag}_grad_{k}']
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-9)


def test_stream_plan_follows_the_reference_modules():
    configs = _configs()
    out = {k: None for k in ('rgb_coarse', 'rgb_fine', 'points_augmentation_rgb_coarse', 'views_augmentation_rgb_coarse')}
    inp = {'indices_mask_sparse_depth': None}
    assert stream_plan(configs, 'MSE01', inp, out) == [('rgb_coarse', 'target_rgb', 'indices_mask_nerf'), ('rgb_fine', 'target_rgb', 'indices_mask_nerf')]
    assert stream_plan(configs, 'MSE02', inp, out) == [('points_augmentation_rgb_coarse', 'target_rgb', 'indices_mask_nerf')]
    assert stream_plan(configs, 'MSE02', inp, {}) == []                      # MSE02.py:32: key absent in eval
    assert stream_plan(configs, 'SparseDepthMSE01', inp, out) == [('depth_fine', 'sparse_depth_values', 'indices_mask_sparse_depth')]
    assert stream_plan(configs, 'SparseDepthMSE03', inp, out) == [('views_augmentation_depth_coarse', 'sparse_depth_values', 'indices_mask_sparse_depth')]
    assert stream_plan(configs, 'SparseDepthMSE01', {}, out) == []           # SparseDepthMSE01.py:31: full images carry no mask
    vanilla = dict(synthetic.make_configs('vanilla'), losses=LOSSES[:2])
    del vanilla['model']['fine_mlp']
    assert stream_plan(vanilla, 'SparseDepthMSE01', inp, out) == [('depth_coarse', 'sparse_depth_values', 'indices_mask_sparse_depth')]
    assert get_loss_weight({'iter_weights': {'0': 0, '10000': 0.1}}, 9999) == 0
    assert get_loss_weight({'iter_weights': {'0': 0, '10000': 0.1}}, 10000) == 0.1


def test_unknown_loss_raises_and_cpu_tensors_are_refused():
    configs = dict(_configs(), losses=LOSSES + [dict(name='PointsAugmentationDepthLoss02', iter_weights={'0': 0, '10000': 0.1})])
    g = gu.load('losses.npz')
    inp, out = _case(g, 'b')
    with pytest.raises(RuntimeError, match='CUDA'):
        FusedLossComputer(configs).compute_losses(inp, out)
    inp['iter_num'] = 20000
    with pytest.raises(RuntimeError, match='Unknown Loss Function'):
        FusedLossComputer(configs).compute_losses(inp, out)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_fused_losses_match_reference(tag):
    g = gu.load('losses.npz')
    configs = _configs()
    inp, out = _case(g, tag, 'cuda:0')
    res = FusedLossComputer(configs).compute_losses(inp, out)
    for lc in LOSSES:
        torch.testing.assert_close(res[lc['name']]['loss_value'].detach().cpu(), g[f"{tag}_loss_{lc['name']}"], rtol=RTOL, atol=1e-9)
    torch.testing.assert_close(res['TotalLoss'].detach().cpu(), g[f'{tag}_loss_TotalLoss'], rtol=RTOL, atol=0)
    res['TotalLoss'].backward()
    for k, v in out.items():
        got = v.grad.cpu() if v.grad is not None else torch.zeros(v.shape)
        torch.testing.assert_close(got, g[f'{tag}_grad_{k}'], rtol=RTOL, atol=1e-9)
    # a second call reuses the workspace (ticket counter back at zero) and reproduces the values bit for bit
    inp2, out2 = _case(g, tag, 'cuda:0')
    res2 = FusedLossComputer(configs).compute_losses(inp2, out2)
    assert torch.equal(res2['TotalLoss'], res['TotalLoss'])


@pytest.mark.gpu
def test_ray_losses_general_gradient_and_large_batch():
    gen = torch.Generator().manual_seed(3)
    n = 100003
    p = [torch.rand((n, 3), generator=gen), torch.rand((n,), generator=gen)]
    t = [torch.rand((n, 3), generator=gen), torch.rand((n,), generator=gen)]
    m = [torch.rand((n,), generator=gen) < 0.5, None]
    w = [0.7, 0.1]
    coef = torch.tensor([0.3, -2.0, 1.5])
    pc = [x.clone().requires_grad_() for x in p]
    want = torch.stack([loss_oracle.masked_mse(pc[0], t[0], m[0]), loss_oracle.masked_mse(pc[1], t[1], torch.ones(n, dtype=torch.bool))])
    want = torch.cat([want, (want * torch.tensor(w)).sum()[None]])
    (want * coef).sum().backward()
    pg = [x.cuda().requires_grad_() for x in p]
    got = ray_losses(pg, [x.cuda() for x in t], [m[0].cuda(), None], w)
    torch.testing.assert_close(got.detach().cpu(), want.detach(), rtol=1e-5, atol=0)
    (got * coef.cuda()).sum().backward()
    for a, b in zip(pg, pc):
        torch.testing.assert_close(a.grad.cpu(), b.grad, rtol=1e-5, atol=1e-12)
