"""CPU-side checks: the C-ABI library loads and exports what include/simplenerf_b200.h declares, argument
validation never reaches a kernel, and the host-side mirror of the reference interface behaves like it."""
import ctypes
import os
import re

import pytest
import torch

from simplenerf_b200 import _lib, build, synthetic
from simplenerf_b200.models import get_model
from simplenerf_b200.models.FusedSimpleNeRF01 import FusedSimpleNeRF, MlpBlock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, 'include', 'simplenerf_b200.h')).read()
    declared = set(re.findall(r'\b(snerf_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.snerf_abi_version() == 1


def test_flag_values_match_the_header():
    header = open(os.path.join(ROOT, 'include', 'simplenerf_b200.h')).read()
    declared = {name: int(val) for name, val in re.findall(r'#define\s+SNERF_(FLAG_[A-Z_]+)\s+(\d+)u', header)}
    assert declared and all(getattr(_lib, name) == val for name, val in declared.items()), declared
    assert len(set(declared.values())) == len(declared) and all(v & (v - 1) == 0 for v in declared.values())     # distinct single bits


def test_argument_validation_without_gpu(lib):
    assert lib.snerf_sample_coarse(None, None, None, None, None, 4, 64, 0, None) == 1
    assert b'null' in lib.snerf_last_error()
    assert lib.snerf_sample_fine(None, None, None, 0, None, None, None, None, None, 4, 64, 128, None) == 1
    assert lib.snerf_composite_forward(*([None] * 15), 4, 64, 0, None) == 1
    desc = _lib.MlpDesc(depth=8, width=128, skip_layer=4, pts_degree=10, trunk_degree=10, view_degree=4, view_width=128,
                        head_out=1)
    assert lib.snerf_mlp_workspace_bytes(ctypes.byref(desc), 16, 64, 0) == 0     # unsupported width
    assert b'width' in lib.snerf_last_error()
    good = MlpBlock(synthetic.make_configs()['model']['coarse_mlp']).desc
    assert lib.snerf_mlp_workspace_bytes(ctypes.byref(good), 16, 64, _lib.FLAG_PRECISE | _lib.FLAG_SAVE_FOR_BWD) > 0
    assert lib.snerf_mlp_forward(ctypes.byref(good), None, None, None, None, None, None, None, None, None, None, 0, 1, 1, 0,
                                 None) == 1


def test_model_factory_contract():
    configs = synthetic.make_configs('simplenerf')
    assert isinstance(get_model(configs, None), FusedSimpleNeRF)
    with pytest.raises(RuntimeError, match='Unknown model'):
        get_model(dict(configs, model=dict(configs['model'], name='Nope01')), None)


def test_state_dict_layout_matches_survey():
    model = get_model(synthetic.make_configs('simplenerf'), None)
    sd = model.state_dict()
    assert sum(v.numel() for v in sd.values()) == 2265488
    assert tuple(sd['coarse_model.pts_linears.0.weight'].shape) == (256, 63)
    assert tuple(sd['coarse_model.pts_linears.5.weight'].shape) == (256, 319)
    assert tuple(sd['pts_aug_coarse_model.pts_linears.0.weight'].shape) == (256, 21)
    assert tuple(sd['pts_aug_coarse_model.pts_linears.5.weight'].shape) == (256, 277)
    assert tuple(sd['pts_aug_coarse_model.views_linears.0.weight'].shape) == (128, 325)
    assert tuple(sd['views_aug_coarse_model.pts_output_linear.weight'].shape) == (4, 256)
    assert 'views_aug_coarse_model.feature_linear.weight' not in sd
    vanilla = get_model(synthetic.make_configs('vanilla'), None)
    assert set(k.split('.')[0] for k in vanilla.state_dict()) == {'coarse_model', 'fine_model'}


def test_cpu_batch_is_refused_loudly():
    model = get_model(synthetic.make_configs('vanilla'), None)
    with pytest.raises(RuntimeError, match='CUDA'):
        model(synthetic.make_ray_batch('llff', 4, 0))


def test_unbuilt_features_raise():
    cfg = synthetic.make_configs('vanilla')
    cfg['model']['coarse_mlp']['views_net_depth'] = 2
    with pytest.raises(NotImplementedError):
        get_model(cfg, None)
    cfg = synthetic.make_configs('vanilla')
    cfg['model']['coarse_mlp']['predict_visibility'] = True          # row a14 / N4: built on both paths
    model = get_model(cfg, None)
    assert tuple(model.coarse_model.views_output_linear.weight.shape) == (4, 128) and model.predict_visibility


def test_synthetic_rays_match_reference_geometry():
    b = synthetic.make_ray_batch('llff', 64, 3)
    assert torch.allclose(b['view_dirs'].norm(dim=-1), torch.ones(64), atol=1e-6)
    # NDC origins sit on the near plane z = -1 (o2 = 1 + 2 near / oz with oz = -near)
    assert torch.allclose(b['rays_o_ndc'][:, 2], -torch.ones(64), atol=1e-5)
    f = synthetic.make_ray_batch('llff', 1008 * 2, 0, frame=True)
    assert f['rays_o'].shape == (2016, 3)


# ---- N2 / N3 host logic that needs no GPU -----------------------------------------------------------------------
class _StubModel(torch.nn.Module):
    """rgb_fine = w * mean(rays_d) per ray: enough to see gradients accumulate over sub-batches."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.tensor(2.0))
        self.calls = []

    def forward(self, batch):
        self.calls.append((batch['rays_o'].shape[0], batch['iter_num'], sorted(batch['common_data'])))
        return {'rgb_fine': self.w * batch['rays_d'].mean(-1, keepdim=True).expand(-1, 3)}


class _StubLosses:
    def compute_losses(self, input_dict, output_dict):
        mse = ((output_dict['rgb_fine'] - input_dict['target_rgb']) ** 2).mean()
        return {'MSE01': {'loss_value': mse}, 'TotalLoss': 1.0 * mse}


def test_train_step_mirrors_train_one_iter_on_cpu():
    """RayShardedTrainStep (src/Trainer01.py:61-107): sub-batches of `sub_batch_size`, per-sub-batch backward with
    accumulating gradients, loss values summed over sub-batches (update_losses_dict_ with num_samples_=1), one
    optimizer step; tensors with one row per ray are sliced, everything else is passed through."""
    from simplenerf_b200.trainer import RayShardedTrainStep
    g = torch.Generator().manual_seed(0)
    n = 10
    batch = {'rays_o': torch.zeros(n, 3), 'rays_d': torch.rand((n, 3), generator=g), 'target_rgb': torch.rand((n, 3), generator=g),
             'iter_num': 7, 'num_frames': 3, 'common_data': {'poses': torch.eye(4)[None].repeat(n, 1, 1)}}   # poses: n rows, but not per ray
    model = _StubModel()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    step = RayShardedTrainStep({'sub_batch_size': 4}, model, _StubLosses(), opt)
    losses = step(batch)
    assert [c[0] for c in model.calls] == [4, 4, 2] and all(c[1] == 7 and c[2] == ['poses'] for c in model.calls)
    # the same arithmetic by hand
    w = torch.tensor(2.0, requires_grad=True)
    total = 0
    for lo in (0, 4, 8):
        sl = slice(lo, lo + 4)
        mse = ((w * batch['rays_d'][sl].mean(-1, keepdim=True).expand(-1, 3) - batch['target_rgb'][sl]) ** 2).mean()
        mse.backward()
        total = total + mse.detach()
    torch.testing.assert_close(losses['TotalLoss'], total)
    torch.testing.assert_close(losses['MSE01'], total)
    torch.testing.assert_close(model.w.detach(), torch.tensor(2.0) - 0.1 * w.grad)
    assert step(batch, shard=True)['TotalLoss'].shape == ()     # one rank: the shard is the whole batch


def test_fused_loss_computer_plumbing_without_gpu():
    """extra_losses are added with the reference's weight rule; a loss that is neither fused nor supplied raises only when its
    weight is not 0; CPU tensors are refused before any kernel is reached."""
    from simplenerf_b200.loss_functions import FusedLossComputer
    configs = dict(synthetic.make_configs('simplenerf'),
                   losses=[{'name': 'SomeOtherLoss01', 'iter_weights': {'0': 0, '100': 0.5}}])

    class Extra:
        def compute_loss(self, input_dict, output_dict, return_loss_maps=False):
            return {'loss_value': output_dict['x'] * 2}

    inp, out = {'iter_num': 50, 'rays_o': torch.zeros(2, 3)}, {'x': torch.tensor(3.0)}
    assert float(FusedLossComputer(configs).compute_losses(inp, out)['TotalLoss']) == 0             # weight 0: skipped
    got = FusedLossComputer(configs, extra_losses={'SomeOtherLoss01': Extra()}).compute_losses(dict(inp, iter_num=100), out)
    assert float(got['TotalLoss']) == 3.0 and float(got['SomeOtherLoss01']['loss_value']) == 6.0
    with pytest.raises(RuntimeError, match='Unknown Loss Function'):
        FusedLossComputer(configs).compute_losses(dict(inp, iter_num=100), out)
    # return_loss_maps (validation): every reported loss carries a `loss_maps` dict; extra losses keep theirs (none here)
    got = FusedLossComputer(configs, extra_losses={'SomeOtherLoss01': Extra()}).compute_losses(dict(inp, iter_num=100), out, return_loss_maps=True)
    assert got['SomeOtherLoss01']['loss_maps'] == {} and float(got['TotalLoss']) == 3.0


def test_argument_validation_of_the_next_row_entry_points(lib):
    """Loss, gather and visibility-head entry points reject bad arguments before any launch (no GPU needed)."""
    streams = (_lib.LossStream * 1)()
    assert lib.snerf_ray_losses_forward(streams, 1, 4, None, None, None, 0, None) == 1          # null pred / target
    assert b'null' in lib.snerf_last_error()
    assert lib.snerf_ray_losses_forward(streams, 9, 4, None, None, None, 0, None) == 1 and b'streams' in lib.snerf_last_error()
    assert lib.snerf_ray_losses_workspace_bytes() >= 148 * 8 * 8
    assert lib.snerf_ray_losses_backward(streams, 0, 4, None, None, None) == 1
    args = _lib.ReprojArgs()
    assert lib.snerf_reprojection_losses_forward(ctypes.byref(args), 4, None, None, None, None, 0, None) == 1
    assert b'other depths' in lib.snerf_last_error()
    tables = (_lib.GatherTable * 1)()
    assert lib.snerf_gather_rows(tables, 1, None, 0, None) == 0                                 # empty batch: nothing to do
    assert lib.snerf_gather_rows(tables, 1, None, 3, None) == 1
    assert lib.snerf_gather_rows(tables, 99, None, 3, None) == 1 and b'tables' in lib.snerf_last_error()
    good = MlpBlock(synthetic.make_configs()['model']['coarse_mlp']).desc
    assert lib.snerf_visibility_workspace_bytes(ctypes.byref(good), 16, 64, 2) > 0
    noview = MlpBlock(synthetic.make_configs('simplenerf')['model']['views_augmentation']['coarse_mlp']).desc
    assert lib.snerf_visibility_workspace_bytes(ctypes.byref(noview), 16, 64, 2) == 0           # no view branch, no head
    assert lib.snerf_visibility2_composite_forward(None, None, None, None, 4, 64, 9, None) == 1  # more than 8 other views
    assert lib.snerf_visibility2_composite_forward(None, None, None, None, 0, 64, 2, None) == 0


def test_fused_adam_state_has_the_torch_adam_layout_both_ways():
    """Trainer01.py:352-381 saves optimizer.state_dict() and resumes with load_state_dict(); a checkpoint
    written with torch.optim.Adam must load into FusedAdam and the other way round."""
    from simplenerf_b200.optim import from_adam_state, to_torch_adam_state
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in ((4, 3), (5,), (2, 2))]
    ref = torch.optim.Adam(params, lr=5e-4, betas=(0.9, 0.999))
    for _ in range(3):
        for p in params:
            p.grad = torch.randn_like(p)
        ref.step()
    sd = ref.state_dict()
    step, m, v, group = from_adam_state(sd, len(params))                        # torch -> ours
    assert step == 3 and group['lr'] == 5e-4
    for i, p in enumerate(params):
        torch.testing.assert_close(m[i], ref.state[p]['exp_avg'])
        torch.testing.assert_close(v[i], ref.state[p]['exp_avg_sq'])
    ours = to_torch_adam_state(step, m, v, {'lr': 1e-4, 'betas': (0.9, 0.999), 'eps': 1e-8})   # ours -> torch
    assert set(ours['param_groups'][0]) == set(sd['param_groups'][0])
    fresh = [torch.nn.Parameter(p.detach().clone()) for p in params]
    other = torch.optim.Adam(fresh, lr=5e-4)
    other.load_state_dict(ours)
    assert other.param_groups[0]['lr'] == 1e-4
    for p, q in zip(params, fresh):                                              # and the next update is the same update
        p.grad = torch.ones_like(p)
        q.grad = torch.ones_like(q)
    ref.param_groups[0]['lr'] = 1e-4
    ref.step()
    other.step()
    for p, q in zip(params, fresh):
        torch.testing.assert_close(p, q)
    # the flat layout round 1 wrote still loads
    step2, m2, _, _ = from_adam_state({'step': 7, 'exp_avg': m, 'exp_avg_sq': v, 'param_groups': [{'lr': 1.0}]}, len(params))
    assert step2 == 7 and m2[0] is m[0]
    # never-stepped optimizer: torch keeps no state
    assert to_torch_adam_state(0, m, v, {'lr': 1e-4, 'betas': (0.9, 0.999), 'eps': 1e-8})['state'] == {}
