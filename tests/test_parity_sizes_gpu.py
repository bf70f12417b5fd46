"""Parity that can fail, at the sizes BASELINE.json names (VERDICT r1 "weak" 1-3, "next round" 2-3).

* trained-scale MLP fixture from the UNMODIFIED reference (tests/golden/mlp_trained.npz, oracle/make_golden_trained.py): rgb spans
  (0, 1), sigma is O(1); the bf16 kernels are held to a fraction of the SIGNAL's spread, and a deliberately broken layer
  (one 64-wide K chunk of one weight matrix dropped) must miss that bound by a wide margin;
* C2 (LLFF, 4096 rays, 4 MLPs, fwd+bwd) and C4's per-GPU shard (RealEstate camera, 4096 rays) against the oracle at full
  size, with north_star's numbers as stated: composited rgb / depth within 1e-3 abs (depth in NDC units unscaled, metric depth
  relative to the scene's depth range), gradients within 1e-2 relative;
* a2: more rays than one launch group (C3's situation), eval and training with injected randoms, against the oracle;
* the near-empty random-init field: depth = sum(w z) / (acc + 1e-6) is a ratio of tiny numbers, so its two sums are compared
  separately instead of being skipped."""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import nerf_oracle as orc
from simplenerf_b200 import ops, synthetic
from simplenerf_b200.models import get_model
from simplenerf_b200.models.FusedSimpleNeRF01 import FixedRandoms, MlpBlock

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
STREAMS = [('rgb_coarse', 'depth_coarse'), ('rgb_fine', 'depth_fine'),
           ('points_augmentation_rgb_coarse', 'points_augmentation_depth_coarse'),
           ('views_augmentation_rgb_coarse', 'views_augmentation_depth_coarse')]


def cuda(t):
    return t.to(DEV).contiguous()


def _has_tc():
    from simplenerf_b200 import _lib
    return bool(_lib.load().snerf_has_tensor_path())


def _to_dev(batch):
    return {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def _build(configs, state, precision, **extra):
    configs = dict(configs, model=dict(configs['model'], precision=precision, **extra))
    model = get_model(configs, None)
    model.load_state_dict(state)
    return model.to(DEV)


# ------------------------------------------------------------------------------------------------
# 1. trained-scale weights: a tolerance relative to the signal, and proof that the test can fail
# ------------------------------------------------------------------------------------------------
RMS_FRAC, MAX_FRAC = 0.03, 0.12        # bf16 forward vs the fp32 reference: measured 0.006-0.018 / 0.025-0.074 of the signal's std


def _run_block(block, precision, pts, vd, noise):
    from simplenerf_b200._lib import FLAG_PRECISE
    n = pts.shape[0]
    flags = FLAG_PRECISE if precision == 'fp32' else 0
    table = [None if p is None else p.detach() for p in block.param_table()]
    packed = None if precision == 'fp32' else block.packed(table, force=True)
    ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n, 1, flags), dtype=torch.uint8, device=DEV)
    sigma, rgb = ops.mlp_forward(block.desc, table, packed, cuda(pts), torch.zeros((n, 3), device=DEV), cuda(vd),
                                 torch.zeros((n, 1), device=DEV), None if noise is None else cuda(noise.reshape(-1)), ws, flags)
    return sigma.cpu().reshape(-1, 1), rgb.cpu().reshape(-1, 3)


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_mlp_trained_scale_vs_reference_golden(precision):
    if precision == 'bf16' and not _has_tc():
        pytest.skip('tensor path not built')
    g = gu.load('mlp_trained.npz')
    configs = synthetic.make_configs('simplenerf')
    for slot, mlp_cfg in orc.model_slots(configs).items():
        if slot == 'fine_model':
            continue
        spec = orc.MlpSpec(mlp_cfg)
        state = synthetic.trained_scale_state(orc.deterministic_state(spec.param_shapes(), int(g[f'{slot}_seed'][0])))
        np.testing.assert_allclose(gu.checksum(state), g[f'{slot}_checksum'].numpy(), rtol=1e-12)
        block = MlpBlock(mlp_cfg)
        block.load_state_dict(state)
        block.to(DEV)

        def errors(training):
            sigma, rgb = _run_block(block, precision, g['pts'], g['view_dirs'], g['noise'] if training else None)
            tag = f"{slot}_{'train' if training else 'eval'}"
            out = {}
            for name, got in (('sigma', sigma), ('rgb', rgb)):
                want = g[f'{tag}_{name}']
                err, sd = got - want, float(want.std())
                out[name] = (float(err.pow(2).mean().sqrt()) / sd, float(err.abs().max()) / sd)
            return out

        for training in (False, True):
            for name, (rms, mx) in errors(training).items():
                if precision == 'fp32':
                    assert mx <= 2e-4, (slot, training, name, mx)
                else:
                    assert rms <= RMS_FRAC and mx <= MAX_FRAC, (slot, training, name, rms, mx)
        # the same check must FAIL on a broken layer: drop one 64-wide K chunk of the fourth trunk layer
        with torch.no_grad():
            kept = block.pts_linears[3].weight[:, 64:128].clone()
            block.pts_linears[3].weight[:, 64:128] = 0
        broken = errors(False)
        with torch.no_grad():
            block.pts_linears[3].weight[:, 64:128] = kept
        for name, (rms, mx) in broken.items():
            assert rms > 5 * RMS_FRAC, (slot, name, 'a dropped K chunk went unnoticed', rms)


# ------------------------------------------------------------------------------------------------
# 2. / 3. full-size training steps against the oracle
# ------------------------------------------------------------------------------------------------
def _randoms(configs, n, seed):
    gen = torch.Generator().manual_seed(seed)
    table = {'t_rand': torch.rand((n, 64), generator=gen), 'u': torch.rand((n, 128), generator=gen)}
    for slot in orc.model_slots(configs):
        table[f'noise_{slot}'] = torch.randn((n * (192 if 'fine' in slot else 64), 1), generator=gen)
    target = torch.rand((n, 3), generator=gen)
    tdepth = 1 + 4 * torch.rand((n,), generator=gen)
    return table, target, tdepth


def _slice_table(table, lo, hi, configs):
    out = {'t_rand': table['t_rand'][lo:hi], 'u': table['u'][lo:hi]}
    for slot in orc.model_slots(configs):
        s = 192 if 'fine' in slot else 64
        out[f'noise_{slot}'] = table[f'noise_{slot}'][lo * s:hi * s]
    return out


def _oracle_step(configs, state, batch, table, target, tdepth, streams, chunk=1024):
    """Forward + backward of the oracle over the whole batch in ray chunks (bounded host memory); the loss is a sum of per-ray
    terms divided by the total element count (the means of the GPU side's loss), so chunked accumulation equals one big backward."""
    n = batch['rays_o'].shape[0]
    oracle = orc.NerfOracle(configs)
    oracle.load_state_dict(state)
    oracle.train()
    outs = {}
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        sub = {k: (v[lo:hi] if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
        oracle.randoms = orc.FixedRandoms(_slice_table(table, lo, hi, configs))
        out = oracle(sub)
        loss = sum(((out[a] - target[lo:hi]) ** 2).sum() / (3 * n) + 0.1 * ((out[b] - tdepth[lo:hi]) ** 2).sum() / n for a, b in streams)
        loss.backward()
        for k, v in out.items():
            outs.setdefault(k, []).append(v.detach())
    return {k: torch.cat(v, 0) for k, v in outs.items()}, {k: p.grad.clone() for k, p in oracle.named_parameters()}


def _check_step(configs, state, batch, camera, kind, precisions):
    n = batch['rays_o'].shape[0]
    streams = STREAMS if kind == 'simplenerf' else STREAMS[:2]
    table, target, tdepth = _randoms(configs, n, 5)
    ref_out, ref_grad = _oracle_step(configs, state, batch, table, target, tdepth, streams)
    depth_range = float(synthetic.CAMERAS[camera]['far'])
    # tolerances as north_star states them; fp32 path far below
    for precision, out_tol, whole, per_mlp in precisions:
        model = _build(configs, state, precision).train()
        model.randoms = FixedRandoms(table)
        out = model(_to_dev(batch))
        t, d = target.to(DEV), tdepth.to(DEV)
        sum(((out[a] - t) ** 2).mean() + 0.1 * ((out[b] - d) ** 2).mean() for a, b in streams).backward()
        ndc = configs['data_loader']['ndc']
        for a, b in streams:
            assert float((out[a].detach().cpu() - ref_out[a]).abs().max()) <= out_tol, (precision, a)
            if 'fine' in b:
                continue          # the fine samples follow the coarse weights: checked with teacher forcing below
            if ndc:               # depth in NDC units, [0, 1]: 1e-3 abs as stated, unscaled
                k = b.replace('depth', 'depth_ndc')
                assert float((out[k].detach().cpu() - ref_out[k]).abs().max()) <= out_tol, (precision, k)
            assert float((out[b].detach().cpu() - ref_out[b]).abs().max()) <= out_tol * depth_range, (precision, b)
        # fine stream in isolation: the oracle's own z_vals_fine through the fine MLP + compositing
        with torch.no_grad():
            model.randoms = FixedRandoms({'noise_fine_model': table['noise_fine_model']})
            fine = {}
            model._stream(fine, 'fine_model', '', 'fine', cuda(ref_out['z_vals_fine']), _to_dev(batch), True)
        keys = ['rgb_fine', 'acc_fine'] + (['depth_ndc_fine'] if ndc else [])
        for k in keys:
            assert float((fine[k].cpu() - ref_out[k]).abs().max()) <= out_tol, (precision, 'teacher-forced', k)
        assert float((fine['depth_fine'].cpu() - ref_out['depth_fine']).abs().max()) <= out_tol * depth_range, precision
        got = {k: p.grad.cpu() for k, p in model.named_parameters()}
        flat = lambda dct, keys: torch.cat([dct[k].flatten() for k in keys])   # noqa: E731
        rel = float((flat(got, ref_grad) - flat(ref_grad, ref_grad)).norm() / flat(ref_grad, ref_grad).norm())
        assert rel <= whole, (precision, 'whole gradient', rel)
        for slot in orc.model_slots(configs):
            keys = [k for k in ref_grad if k.startswith(slot + '.')]
            rel = float((flat(got, keys) - flat(ref_grad, keys)).norm() / flat(ref_grad, keys).norm())
            assert rel <= per_mlp, (precision, slot, rel)
        del model, out
        torch.cuda.empty_cache()


def _precisions():
    p = [('fp32', 2e-5, 5e-5, 2e-4)]
    if _has_tc():
        p.append(('bf16', 1e-3, 1e-2, 1e-2))          # north_star: 1e-3 abs on composited rgb / depth, 1e-2 relative on gradients
    return p


def test_c2_llff_training_step_4096_rays_vs_oracle():
    """BASELINE.json config 2 at its full size: 4096 rays, coarse + fine + points-aug + views-aug MLPs, fwd + bwd."""
    configs = synthetic.make_configs('simplenerf')
    state = gu.full_state(configs, 7, dense=True)
    _check_step(configs, state, synthetic.make_ray_batch('llff', 4096, 1021), 'llff', 'simplenerf', _precisions())


def test_c4_realestate_shard_4096_rays_vs_oracle():
    """BASELINE.json config 4, one GPU's shard: RealEstate-10K camera (576x1024, far = 133), the shipped train0021 model
    (NDC, 4 MLPs), 4096 of the step's 32768 rays."""
    configs = synthetic.make_configs('simplenerf')
    state = gu.full_state(configs, 21, dense=True)
    _check_step(configs, state, synthetic.make_ray_batch('re10k', 4096, 21), 're10k', 'simplenerf', _precisions())


def test_non_ndc_training_step_2048_rays_vs_oracle():
    """The non-NDC branch of the path (metric depths, last interval 1e10: :433-441) at a sub-batch's size."""
    configs = synthetic.make_configs('vanilla', ndc=False)
    state = gu.full_state(configs, 33, dense=True)
    _check_step(configs, state, synthetic.make_ray_batch('llff', 2048, 33), 'llff', 'vanilla', _precisions())


# ------------------------------------------------------------------------------------------------
# a2: more rays than one launch group
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_several_launch_groups_equal_the_oracle(precision):
    """`batchify_rays` / `merge_mini_batch_data` (:81-106, :504-512): the drop-in cuts n > launch_rays rays into launch groups and
    concatenates every output key.  Eval (C3's situation: a frame is 12 groups) and training with injected randoms
    (FixedRandoms must hand every group its own rows), ragged last group, against the oracle on the whole batch."""
    if precision == 'bf16' and not _has_tc():
        pytest.skip('tensor path not built')
    tol = 2e-5 if precision == 'fp32' else 1e-3
    n = 2 * 1024 + 321
    # eval, vanilla, rows of a frame (Tester contract: retraw False)
    configs = synthetic.make_configs('vanilla')
    state = gu.full_state(configs, 3, dense=True)
    batch = synthetic.make_ray_batch('llff', n, 0, frame=True, start=300 * 1008 + 17)
    oracle = orc.NerfOracle(configs)
    oracle.load_state_dict(state)
    oracle.eval()
    with torch.no_grad():
        want = oracle(batch)
        model = _build(configs, state, precision, launch_rays=1024).eval()
        got = model(_to_dev(batch))
        one = _build(configs, state, precision).eval()(_to_dev(batch))
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert torch.equal(got[k], one[k]), k                                  # grouping does not change a single bit in eval
        if 'fine' in k or 'alpha' in k:
            continue
        scale = max(1.0, float(want[k].abs().max())) if 'depth' in k else 1.0
        assert float((got[k].cpu() - want[k]).abs().max()) <= tol * scale * (5 if 'depth_var' in k else 1), k
    # training, 4 MLPs, injected randoms: group g must read rows [g * launch_rays, ...) of every random table
    configs = synthetic.make_configs('simplenerf')
    state = gu.full_state(configs, 9, dense=True)
    batch = synthetic.make_ray_batch('llff', n, 77)
    table, _, _ = _randoms(configs, n, 13)
    oracle = orc.NerfOracle(configs)
    oracle.load_state_dict(state)
    oracle.randoms = orc.FixedRandoms(table)
    oracle.train()
    with torch.no_grad():
        want = oracle(batch)
        model = _build(configs, state, precision, launch_rays=1024).train()
        model.randoms = FixedRandoms(table)
        got = model(_to_dev(batch))
    assert set(got) == set(want)
    assert torch.equal(got['z_vals_coarse'].cpu(), want['z_vals_coarse'])       # stratified depths are bit exact, so offsets are right
    for a, b in STREAMS:
        assert float((got[a].cpu() - want[a]).abs().max()) <= tol, (a,)
        if 'fine' not in b:
            k = b.replace('depth', 'depth_ndc')
            assert float((got[k].cpu() - want[k]).abs().max()) <= tol, (k,)
    for slot in ('coarse', ):
        assert float((got[f'raw_sigma_{slot}'].cpu() - want[f'raw_sigma_{slot}']).abs().max()) <= (1e-4 if precision == 'fp32' else 0.5)


# ------------------------------------------------------------------------------------------------
# the near-empty random-init field: compare the two sums of the depth ratio
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_random_init_field_depth_sums_vs_reference_golden(precision):
    """tests/golden/render_llff_simplenerf.npz (unmodified reference, raw random-init weights, acc ~ 4e-3): depth is
    sum(w z) / (acc + 1e-6).  Both sums are held to the 1e-3 abs of north_star on every stream, fine pass included; the ratio
    itself is checked where it is well conditioned (relative to the reference's own sensitivity to a 1e-3 change of acc)."""
    if precision == 'bf16' and not _has_tc():
        pytest.skip('tensor path not built')
    configs, state, batch, table, g = gu.render_case('render_llff_simplenerf.npz')
    model = _build(configs, state, precision).train()
    model.randoms = FixedRandoms(dict(table))
    with torch.no_grad():
        out = model(_to_dev(batch))
    tol = 2e-5 if precision == 'fp32' else 1e-3
    for prefix, level in (('', 'coarse'), ('', 'fine'), ('points_augmentation_', 'coarse'), ('views_augmentation_', 'coarse')):
        acc, acc_ref = out[f'{prefix}acc_{level}'].cpu(), g[f'train__{prefix}acc_{level}']
        assert float((acc - acc_ref).abs().max()) <= tol, (prefix, level, 'acc')
        for dk in ('depth', 'depth_ndc'):
            d, d_ref = out[f'{prefix}{dk}_{level}'].cpu(), g[f'train__{prefix}{dk}_{level}']
            swz, swz_ref = d * (acc + 1e-6), d_ref * (acc_ref + 1e-6)
            scale = max(1.0, float(d_ref.abs().max()))
            # fine pass: a resampled depth may fall on the other side of a bin edge (the coarse weights differ in the last bits);
            # metric depth: the NDC far samples convert to z ~ 1e3 (:495-501), so sum(w z) of a translucent ray carries the fp32
            # rounding of terms that large -- 1e-3 relative to the depth range there, the stated bound on the NDC depth
            lim = (max(tol, 1e-3) if (level == 'fine' or dk == 'depth') else tol) * scale
            assert float((swz - swz_ref).abs().max()) <= lim, (prefix, level, dk, 'sum w z')


# ------------------------------------------------------------------------------------------------
# row X1: evaluation with the samples kept on chip (snerf_render_forward) against the two-kernel path and the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('camera', ['llff', 're10k'])
def test_fused_evaluation_equals_the_two_kernel_path(camera):
    """Tester contract (retraw False, no_grad): the drop-in composites inside the MLP kernel's head epilogue (one record per 32
    samples + a per-ray fold).  Same key set; per-ray maps and alpha of the coarse pass equal the MLP-kernel + compositing-kernel
    path to fp32 summation-order noise (the sigma / rgb values are the same bits: only the order of the sums differs); the fine
    pass inherits ~1e-6 differences of the coarse weights through sample_pdf, so it gets the parity tolerance; the oracle is
    met within north_star's 1e-3 as before."""
    if not _has_tc():
        pytest.skip('tensor path not built')
    n = 1024 + 96
    configs = synthetic.make_configs('vanilla', ndc=(camera == 'llff'))
    state = gu.full_state(configs, 5, dense=True)
    batch = synthetic.make_ray_batch(camera, n, 3)
    with torch.no_grad():
        fused = _build(configs, state, 'bf16').eval()
        assert fused.fused_composite
        got = fused(_to_dev(batch))
        two = _build(configs, state, 'bf16', fused_composite=False).eval()(_to_dev(batch))
        oracle = orc.NerfOracle(configs)
        oracle.load_state_dict(state)
        oracle.eval()
        want = oracle(batch)
    assert set(got) == set(two) == set(want)
    for k in two:
        assert got[k].shape == two[k].shape, k
        a, b = got[k].float().cpu(), two[k].float().cpu()
        scale = max(1.0, float(b.abs().max()))
        tol = 2e-6 if 'coarse' in k else 1e-3
        if 'depth_var' in k:
            tol *= 50          # sum of w (z - depth)^2: the centred per-run form is the better conditioned of the two
        assert float((a - b).abs().max()) <= tol * scale, (k, float((a - b).abs().max()), scale)
    for k in ('rgb_coarse', 'acc_coarse', 'depth_coarse'):
        scale = max(1.0, float(want[k].abs().max())) if 'depth' in k else 1.0
        assert float((got[k].cpu() - want[k]).abs().max()) <= 1e-3 * scale, k


@pytest.mark.parametrize('s,ndc,white', [(32, False, True), (96, True, False), (64, True, True), (192, False, False)])
def test_render_forward_entry_point_shapes_and_flags(s, ndc, white):
    """snerf_render_forward against snerf_mlp_forward + snerf_composite_forward on the same rays: run lengths other than the
    model's (one run per ray at 32 samples, three at 96), white background, metric and NDC depths, a ray count that fills neither
    a fold block nor a tile; weights on request equal the compositing kernel's."""
    if not _has_tc():
        pytest.skip('tensor path not built')
    n = 1001
    cfg = synthetic.make_configs('vanilla')['model']['coarse_mlp']
    block = MlpBlock(cfg).to(DEV)
    with torch.no_grad():
        block.pts_output_linear.weight.mul_(30.0)          # a dense field: acc ~ 1, depth well conditioned
        block.pts_output_linear.bias.add_(5.0)
    b = synthetic.make_ray_batch('llff', n, 9)
    g = torch.Generator().manual_seed(s)
    z = torch.sort(torch.rand((n, s), generator=g), -1)[0] if ndc else 1 + 5 * torch.sort(torch.rand((n, s), generator=g), -1)[0]
    z = cuda(z)
    rays_o, rays_d, vd = cuda(b['rays_o']), cuda(b['rays_d']), cuda(b['view_dirs'])
    pts_o, pts_d = (cuda(b['rays_o_ndc']), cuda(b['rays_d_ndc'])) if ndc else (rays_o, rays_d)
    table = [None if p is None else p.detach() for p in block.param_table()]
    packed = block.packed(table, force=True)
    ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n, s, 0), dtype=torch.uint8, device=DEV)
    sigma, rgb = ops.mlp_forward(block.desc, table, packed, pts_o, pts_d, vd, z, None, ws, 0)
    want = ops.composite_forward(sigma, rgb, z, rays_o, rays_d, pts_d if ndc else None, ndc, white)
    got = ops.render_forward(block.desc, table, packed, pts_o, pts_d, vd, z, rays_o, rays_d, ndc, white, want_weights=True)
    assert set(got) == set(want) - {'visibility'}
    for k, v in got.items():
        scale = max(1.0, float(want[k].abs().max()))
        tol = 1e-4 if 'depth_var' in k else 3e-6
        assert float((v - want[k]).abs().max()) <= tol * scale, (k, float((v - want[k]).abs().max()), scale)
