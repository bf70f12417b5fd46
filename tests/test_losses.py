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
    configs = dict(_configs(), losses=LOSSES + [dict(name='SomeOtherLoss01', iter_weights={'0': 0, '10000': 0.1})])
    g = gu.load('losses.npz')
    inp, out = _case(g, 'b')
    with pytest.raises(RuntimeError, match='CUDA'):
        FusedLossComputer(configs).compute_losses(inp, out)
    inp['iter_num'] = 20000
    with pytest.raises(RuntimeError, match='Unknown Loss Function'):
        FusedLossComputer(configs).compute_losses(inp, out)


REPROJ = [dict(name=n, iter_weights={'0': 0, '10000': 0.1}, rmse_threshold=0.1, patch_size=[5, 5])
          for n in ('PointsAugmentationDepthLoss02', 'ViewsAugmentationDepthLoss02', 'CoarseFineConsistencyLoss02')]
REPROJ_OTHER = {'PointsAugmentationDepthLoss02': 'points_augmentation_depth_coarse',
                'ViewsAugmentationDepthLoss02': 'views_augmentation_depth_coarse', 'CoarseFineConsistencyLoss02': 'depth_fine'}


def _reproj_case(g, tag, device='cpu'):
    inp, out = _case(g, tag, device)
    inp['iter_num'] = 20000
    inp['common_data'] = {k[len(tag) + 8:]: v.to(device) for k, v in g.items() if k.startswith(f'{tag}_common_')}
    inp['common_data']['resolution'] = tuple(int(x) for x in inp['common_data']['resolution'])
    return inp, out


@pytest.mark.parametrize('tag', ['r', 's'])
def test_reprojection_oracle_matches_reference(tag):
    """The restatement reproduces the unmodified modules, including the one-sided gradient that their in-place masking
    on detach() aliases produces (only the main coarse depth receives a gradient)."""
    g = gu.load('losses.npz')
    inp, out = _reproj_case(g, tag)
    cd = inp['common_data']
    total = 0
    for lc in REPROJ:
        loss = loss_oracle.reprojection_depth_loss(out['depth_coarse'], out[REPROJ_OTHER[lc['name']]], inp['indices_mask_nerf'],
                                                   inp['rays_o'], inp['rays_d'], cd['poses'], cd['images'], inp['pixel_id'],
                                                   cd['intrinsics'], cd['resolution'])
        if lc['name'] == 'CoarseFineConsistencyLoss02':       # compute_loss_sd: the data_loader block has sparse_depth
            loss = loss + loss_oracle.masked_mse(out['depth_coarse'], out['depth_fine'].detach(), inp['indices_mask_sparse_depth'])
        torch.testing.assert_close(loss.detach(), g[f"{tag}_loss_{lc['name']}"], rtol=1e-6, atol=0)
        total = total + 0.1 * loss
    torch.testing.assert_close(total.detach(), g[f'{tag}_loss_TotalLoss'], rtol=1e-6, atol=0)
    total.backward()
    for k, v in out.items():
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        torch.testing.assert_close(got, g[f'{tag}_grad_{k}'], rtol=1e-5, atol=1e-9)
    assert int((g[f'{tag}_grad_depth_coarse'] != 0).sum()) > 0 and not bool(g[f'{tag}_grad_depth_fine'].any())


def _reproj_configs():
    configs = dict(synthetic.make_configs('simplenerf'), losses=[dict(lc) for lc in REPROJ])
    configs['data_loader'] = dict(configs['data_loader'], sparse_depth={})
    return configs


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['r', 's'])
def test_fused_reprojection_losses_match_reference(tag):
    g = gu.load('losses.npz')
    inp, out = _reproj_case(g, tag, 'cuda:0')
    res = FusedLossComputer(_reproj_configs()).compute_losses(inp, out)
    # a projection that lands within an ulp of a pixel boundary may round to the other pixel than on the host, which moves
    # one ray between the masks: values to 1e-3 relative, gradients equal except on a handful of rays
    for lc in REPROJ:
        torch.testing.assert_close(res[lc['name']]['loss_value'].detach().cpu(), g[f"{tag}_loss_{lc['name']}"], rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(res['TotalLoss'].detach().cpu(), g[f'{tag}_loss_TotalLoss'], rtol=1e-3, atol=1e-7)
    res['TotalLoss'].backward()
    for k, v in out.items():
        got = v.grad.cpu() if v.grad is not None else torch.zeros(v.shape)
        want = g[f'{tag}_grad_{k}']
        bad = ~torch.isclose(got, want, rtol=1e-4, atol=1e-8)
        assert int(bad.sum()) <= max(1, got.numel() // 500), (k, int(bad.sum()))


@pytest.mark.gpu
def test_reprojection_masks_and_symmetric_form():
    from simplenerf_b200.loss_functions import reprojection_losses
    g = gu.load('losses.npz')
    inp, out = _reproj_case(g, 'r', 'cuda:0')
    cpu_inp, cpu_out = _reproj_case(g, 'r')
    cd, ccd = inp['common_data'], cpu_inp['common_data']
    others = ['points_augmentation_depth_coarse', 'depth_fine']
    values, codes = reprojection_losses(out['depth_coarse'], [out[k] for k in others], [0.1, 0.1], inp['rays_o'], inp['rays_d'],
                                        inp['pixel_id'], inp['indices_mask_nerf'], cd['images'], cd['poses'], cd['intrinsics'],
                                        symmetric=True)
    want_total = 0
    for j, k in enumerate(others):
        m1, m2 = loss_oracle.reprojection_masks(cpu_out['depth_coarse'], cpu_out[k], cpu_inp['indices_mask_nerf'], cpu_inp['rays_o'],
                                                cpu_inp['rays_d'], ccd['poses'], ccd['images'], cpu_inp['pixel_id'], ccd['intrinsics'],
                                                ccd['resolution'])
        code = codes[j].cpu()[cpu_inp['indices_mask_nerf']]
        assert int(((code & 1).bool() != m1).sum()) + int(((code & 2).bool() != m2).sum()) <= 2
        assert not bool(codes[j].cpu()[~cpu_inp['indices_mask_nerf']].any())
        assert int(m1.sum()) > 10 and int(m2.sum()) > 10                 # both directions are exercised by the fixture
        want = loss_oracle.reprojection_depth_loss(cpu_out['depth_coarse'], cpu_out[k], cpu_inp['indices_mask_nerf'], cpu_inp['rays_o'],
                                                   cpu_inp['rays_d'], ccd['poses'], ccd['images'], cpu_inp['pixel_id'], ccd['intrinsics'],
                                                   ccd['resolution'], symmetric=True)
        torch.testing.assert_close(values[j].detach().cpu(), want.detach(), rtol=2e-3, atol=1e-6)
        want_total = want_total + 0.1 * want
    values[-1].backward()
    want_total.backward()
    for k in ['depth_coarse'] + others:
        bad = ~torch.isclose(out[k].grad.cpu(), cpu_out[k].grad, rtol=1e-4, atol=1e-8)
        assert int(bad.sum()) <= 3, (k, int(bad.sum()))
    assert bool(out['depth_fine'].grad.any())                            # the symmetric form does reach the other model


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


@pytest.mark.gpu
def test_train_step_driver_runs_the_whole_iteration():
    """RayShardedTrainStep = Trainer01.train_one_iter for one rank: batch assembly -> model -> fused losses -> backward ->
    one-launch Adam, in sub-batches; the total loss falls over a few iterations and every configured loss is reported."""
    from simplenerf_b200.models import get_model
    from simplenerf_b200.optim import FusedAdam
    from simplenerf_b200.trainer import RayShardedTrainStep
    g = gu.load('losses.npz')
    inp, _ = _reproj_case(g, 'r', 'cuda:0')
    n = inp['rays_o'].shape[0]
    configs = dict(synthetic.make_configs('simplenerf', ndc=False), losses=[dict(lc) for lc in LOSSES + REPROJ], sub_batch_size=512)
    configs['data_loader'] = dict(configs['data_loader'], sparse_depth={})
    gen = torch.Generator().manual_seed(5)
    d = inp['rays_d']
    batch = dict(inp, view_dirs=d / d.norm(dim=-1, keepdim=True), near=torch.full((n, 1), 2.0, device='cuda'),
                 far=torch.full((n, 1), 6.0, device='cuda'), num_frames=3,
                 target_rgb=torch.rand((n, 3), generator=gen).cuda(), sparse_depth_values=torch.full((n, 1), 4.0, device='cuda'))
    model = get_model(configs, None)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synthetic.densify_state(synthetic.deterministic_state(shapes, 2)))
    model = model.to('cuda:0').train()
    step = RayShardedTrainStep(configs, model, FusedLossComputer(configs, ray_sharded=True), FusedAdam(model.parameters(), lr=5e-4))
    totals = []
    for it in range(6):
        batch['iter_num'] = 20000 + it
        losses = step(batch)
        assert set(losses) == {lc['name'] for lc in LOSSES + REPROJ} | {'TotalLoss'}
        totals.append(float(losses['TotalLoss']))
    assert all(torch.isfinite(torch.tensor(totals))) and totals[-1] < totals[0], totals


PAIRS = [dict(name='PointsAugmentationDepthLoss01', weight=0.3), dict(name='ViewsAugmentationDepthLoss01', weight=0.2),
         dict(name='CoarseFineConsistencyLoss01', weight=0.5), dict(name='DenseDepthMSE01', weight=0.7)]
PAIR_KEYS = {'PointsAugmentationDepthLoss01': ('depth_coarse', 'points_augmentation_depth_coarse'),
             'ViewsAugmentationDepthLoss01': ('depth_coarse', 'views_augmentation_depth_coarse'),
             'CoarseFineConsistencyLoss01': ('depth_coarse', 'depth_fine')}


def test_two_sided_depth_losses_oracle_matches_reference():
    """PointsAugmentationDepthLoss01.py:59-74, CoarseFineConsistencyLoss01.py:37-41: mean((a - b)^2) over every ray with
    gradients to both sides; DenseDepthMSE01.py:56-68: the masked mean against the dense depth prior (coarse and fine)."""
    g = gu.load('losses.npz')
    inp, out = _case(g, 'p')
    total = 0
    for lc in PAIRS:
        if lc['name'] in PAIR_KEYS:
            a, b = (out[k] for k in PAIR_KEYS[lc['name']])
            loss = torch.mean(torch.square(a - b))
        else:
            loss = sum(loss_oracle.masked_mse(out[k], inp['dense_depth_values'][:, 0], inp['indices_mask_nerf']) for k in ('depth_coarse', 'depth_fine'))
        torch.testing.assert_close(loss.detach(), g[f"p_loss_{lc['name']}"], rtol=1e-6, atol=0)
        total = total + lc['weight'] * loss
    torch.testing.assert_close(total.detach(), g['p_loss_TotalLoss'], rtol=1e-6, atol=0)
    total.backward()
    for k, v in out.items():
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        torch.testing.assert_close(got, g[f'p_grad_{k}'], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
def test_fused_two_sided_depth_losses_match_reference():
    g = gu.load('losses.npz')
    configs = dict(synthetic.make_configs('simplenerf'), losses=[dict(lc) for lc in PAIRS])
    inp, out = _case(g, 'p', 'cuda:0')
    res = FusedLossComputer(configs).compute_losses(inp, out)
    for lc in PAIRS:
        torch.testing.assert_close(res[lc['name']]['loss_value'].detach().cpu(), g[f"p_loss_{lc['name']}"], rtol=RTOL, atol=1e-9)
    torch.testing.assert_close(res['TotalLoss'].detach().cpu(), g['p_loss_TotalLoss'], rtol=RTOL, atol=0)
    res['TotalLoss'].backward()
    for k, v in out.items():
        got = v.grad.cpu() if v.grad is not None else torch.zeros(v.shape)
        torch.testing.assert_close(got, g[f'p_grad_{k}'], rtol=1e-5, atol=1e-9)


VIS = [dict(name='VisibilityLoss01', weight=0.4), dict(name='VisibilityPriorLoss01', weight=0.25)]


def _vis_case(g, tag, device='cpu'):
    inp, out = _case(g, tag, device)
    inp['num_frames'] = 3
    for level in ('coarse', 'fine'):      # VisibilityPriorLoss01.py:29-31 only looks for the key
        out[f'raw_visibility2_{level}'] = torch.zeros(1, device=device)
    return inp, out


@pytest.mark.parametrize('tag', ['v', 'w'])
def test_visibility_losses_oracle_matches_reference(tag):
    """VisibilityLoss01.py:56-74 (two-sided MAE, each side detached in turn) and VisibilityPriorLoss01.py:64-80 (mean over the
    NeRF rays of sum_v prior_v (1 - visibility2_v); prior = ones when the batch carries none, :38-41)."""
    g = gu.load('losses.npz')
    inp, out = _vis_case(g, tag)
    prior = inp.get('visibility_prior_masks', torch.ones(inp['rays_o'].shape[0], 2))
    mae = sum((out[f'raw_visibility_{lv}'][..., 0] - out[f'visibility_{lv}'].detach()).abs().mean(1).mean() +
              (out[f'raw_visibility_{lv}'][..., 0].detach() - out[f'visibility_{lv}']).abs().mean(1).mean() for lv in ('coarse', 'fine'))
    m = inp['indices_mask_nerf']
    pri = sum((prior[m] * (1 - out[f'visibility2_{lv}'][m])).sum(1).mean() for lv in ('coarse', 'fine'))
    torch.testing.assert_close(mae.detach(), g[f'{tag}_loss_VisibilityLoss01'], rtol=1e-6, atol=0)
    torch.testing.assert_close(pri.detach(), g[f'{tag}_loss_VisibilityPriorLoss01'], rtol=1e-6, atol=0)
    (0.4 * mae + 0.25 * pri).backward()
    for k, v in out.items():
        if f'{tag}_grad_{k}' in g:
            torch.testing.assert_close(v.grad if v.grad is not None else torch.zeros_like(v), g[f'{tag}_grad_{k}'], rtol=1e-5, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['v', 'w'])
def test_fused_visibility_losses_match_reference(tag):
    g = gu.load('losses.npz')
    configs = dict(synthetic.make_configs('vanilla'), losses=[dict(lc) for lc in VIS])
    inp, out = _vis_case(g, tag, 'cuda:0')
    res = FusedLossComputer(configs).compute_losses(inp, out)
    for lc in VIS:
        torch.testing.assert_close(res[lc['name']]['loss_value'].detach().cpu(), g[f"{tag}_loss_{lc['name']}"], rtol=RTOL, atol=1e-9)
    torch.testing.assert_close(res['TotalLoss'].detach().cpu(), g[f'{tag}_loss_TotalLoss'], rtol=RTOL, atol=0)
    res['TotalLoss'].backward()
    for k, v in out.items():
        if f'{tag}_grad_{k}' in g:
            got = v.grad.cpu() if v.grad is not None else torch.zeros(v.shape)
            torch.testing.assert_close(got, g[f'{tag}_grad_{k}'], rtol=1e-5, atol=1e-9)
    # without the head's outputs the prior loss is skipped like the reference's `return None`
    del out['raw_visibility2_fine']
    assert 'VisibilityPriorLoss01' not in FusedLossComputer(configs).compute_losses(inp, out)


@pytest.mark.gpu
def test_fused_loss_maps_are_the_reference_modules_per_ray_errors():
    """return_loss_maps=True (validation, src/Trainer01.py:195-196): names `<Module>_<level>` (LossUtils01.py:7-10) and values
    `mean((pred[mask] - target[mask])^2, dim=1)` (MSE01.py:53-66); the sparse-depth modules return no maps (SparseDepthMSE01.py:67-70)."""
    g = gu.load('losses.npz')
    configs = _configs()
    inp, out = _case(g, 'a', 'cuda:0')
    res = FusedLossComputer(configs).compute_losses(inp, out, return_loss_maps=True)
    plain = FusedLossComputer(configs).compute_losses(inp, out)
    assert torch.equal(res['TotalLoss'], plain['TotalLoss'])
    m_nerf = inp['indices_mask_nerf']
    want = {
        'MSE01': {'MSE01_coarse': ((out['rgb_coarse'] - inp['target_rgb'])[m_nerf] ** 2).mean(1),
                  'MSE01_fine': ((out['rgb_fine'] - inp['target_rgb'])[m_nerf] ** 2).mean(1)},
        'MSE02': {'MSE02_coarse': ((out['points_augmentation_rgb_coarse'] - inp['target_rgb'])[m_nerf] ** 2).mean(1)},
    }
    assert res['SparseDepthMSE01']['loss_maps'] == {}               # "# No loss maps" (SparseDepthMSE01.py:67-70)
    for name, maps in want.items():
        assert set(maps) <= set(res[name]['loss_maps']), (name, set(res[name]['loss_maps']))
        for key, ref in maps.items():
            got = res[name]['loss_maps'][key]
            assert got.shape == ref.shape and not got.requires_grad
            torch.testing.assert_close(got, ref.detach(), rtol=1e-6, atol=1e-9)
    for lc in LOSSES:
        assert 'loss_maps' in res[lc['name']]
