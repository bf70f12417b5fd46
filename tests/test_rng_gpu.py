"""In-kernel random numbers (reference draw sites src/models/SimpleNeRF01.py:299, :341, :670).

The kernels that consume a random number draw it themselves from a counter-based generator; `snerf_fill_random` writes the very
same numbers to memory.  That identity is the test: the tensor-input entry points (pinned bit-exactly against the oracle
elsewhere) fed with the filled tensors must reproduce the *_rng entry points bit for bit."""
import pytest
import torch

from simplenerf_b200 import ops, synthetic
from simplenerf_b200.models import get_model
from simplenerf_b200.models.FusedSimpleNeRF01 import MlpBlock

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_fill_random_distributions_and_streams():
    n = 1 << 20
    u = ops.RngDraw(1234, 1).materialize((n,), DEV, normal=False)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 2e-3 and abs(float(u.var()) - 1 / 12) < 1e-3
    hist = torch.histc(u, bins=64, min=0, max=1)
    assert float((hist - n / 64).abs().max()) < 6 * (n / 64) ** 0.5                    # flat to counting noise
    g = ops.RngDraw(1234, 2, scale=3.0).materialize((n,), DEV, normal=True)
    assert abs(float(g.mean())) < 1.5e-2 and abs(float(g.std()) - 3.0) < 1.5e-2 and float(g.abs().max()) < 3.0 * 6.5
    assert abs(float(((g / 3.0) ** 4).mean()) - 3.0) < 0.1                               # kurtosis of a normal
    # a draw is a function of (seed, offset, element): repeatable, and every change of the key gives other numbers
    assert torch.equal(u, ops.RngDraw(1234, 1).materialize((n,), DEV, normal=False))
    assert torch.equal(u[:1001], ops.RngDraw(1234, 1).materialize((1001,), DEV, normal=False))   # ragged length: same prefix
    for other in (ops.RngDraw(1234, 3), ops.RngDraw(1235, 1), ops.RngDraw(1234 + (1 << 40), 1), ops.RngDraw(1234, 1 + (1 << 33))):
        v = other.materialize((n,), DEV, normal=False)
        assert float((v == u).float().mean()) < 1e-3
        assert abs(float(((u - 0.5) * (v - 0.5)).mean())) < 1e-3                         # uncorrelated


@pytest.mark.parametrize('n,s,lindisp', [(1000, 64, False), (333, 64, True), (77, 48, False), (50, 33, False)])
def test_sample_coarse_draws_in_kernel_what_fill_random_writes(n, s, lindisp):
    gen = torch.Generator().manual_seed(n)
    near = (0.5 + torch.rand(n, generator=gen)).to(DEV)
    far = near + 1 + 4 * torch.rand(n, generator=gen).to(DEV)
    t = torch.linspace(0., 1., s).to(DEV)
    draw = ops.RngDraw(99, 5)
    want = ops.sample_coarse(near, far, t, draw.materialize((n, s), DEV, normal=False), lindisp)
    got = ops.sample_coarse(near, far, t, draw, lindisp)
    assert torch.equal(got, want)
    assert not torch.equal(got, ops.sample_coarse(near, far, t, ops.RngDraw(99, 6), lindisp))


@pytest.mark.parametrize('n,sc,n_new', [(1001, 64, 128), (500, 64, 64), (300, 64, 256), (97, 64, 96), (64, 40, 128)])
def test_sample_fine_draws_in_kernel_what_fill_random_writes(n, sc, n_new):
    gen = torch.Generator().manual_seed(n + n_new)
    z = torch.sort(torch.rand((n, sc), generator=gen), -1)[0].to(DEV).contiguous()
    w = (torch.rand((n, sc), generator=gen) ** 3).to(DEV)
    draw = ops.RngDraw(7, 11)
    want = ops.sample_fine(z, w, draw.materialize((n, n_new), DEV, normal=False))
    got = ops.sample_fine(z, w, (draw, n_new))
    assert torch.equal(got, want)
    assert bool((got[:, 1:] >= got[:, :-1]).all())


def test_sigma_noise_drawn_in_the_head_epilogue_equals_the_filled_tensor():
    from simplenerf_b200 import _lib
    if not _lib.load().snerf_has_tensor_path():
        pytest.skip('tensor path not built')
    model_cfg = synthetic.make_configs('simplenerf')['model']
    for cfg, n_rays, s in ((model_cfg['coarse_mlp'], 700, 64), (model_cfg['views_augmentation']['coarse_mlp'], 301, 192)):
        torch.manual_seed(0)
        block = MlpBlock(cfg).to(DEV)
        table = [None if p is None else p.detach() for p in block.param_table()]
        packed = block.packed(table)
        b = synthetic.make_ray_batch('llff', n_rays, 3)
        o, d, vd = b['rays_o_ndc'].to(DEV), b['rays_d_ndc'].to(DEV), b['view_dirs'].to(DEV)
        z = torch.sort(torch.rand(n_rays, s, device=DEV), -1)[0].contiguous()
        ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n_rays, s, 0), dtype=torch.uint8, device=DEV)
        draw = ops.RngDraw(2024, 17, scale=0.7)
        noise = draw.materialize((n_rays * s,), DEV, normal=True)
        want_s, want_r = ops.mlp_forward(block.desc, table, packed, o, d, vd, z, noise, ws, 0)
        got_s, got_r = ops.mlp_forward(block.desc, table, packed, o, d, vd, z, draw, ws, 0)
        assert torch.equal(got_s, want_s) and torch.equal(got_r, want_r)
        plain, _ = ops.mlp_forward(block.desc, table, packed, o, d, vd, z, None, ws, 0)
        assert float((plain - got_s).abs().max()) > 0.1                                   # the noise is really there


def test_dropin_trains_on_in_kernel_draws_repeatably():
    """configs['model']['rng'] = 'device' (the default): no random tensor is materialised; torch.manual_seed makes a run repeatable
    and consecutive steps see different numbers."""
    configs = synthetic.make_configs('simplenerf')
    assert configs['model'].get('rng', 'device') == 'device'
    batch = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in synthetic.make_ray_batch('llff', 512, 4).items()}

    def run():
        torch.manual_seed(5)
        model = get_model(configs, None)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(synthetic.densify_state(synthetic.deterministic_state(shapes, 2)))
        model = model.to(DEV).train()
        launches = ops.LAUNCHES['count']
        outs = [model(batch), model(batch)]
        return outs, ops.LAUNCHES['count'] - launches

    (a1, a2), launches = run()
    (b1, b2), _ = run()
    for k in ('z_vals_coarse', 'z_vals_fine', 'raw_sigma_coarse', 'rgb_fine', 'views_augmentation_rgb_coarse'):
        assert torch.equal(a1[k], b1[k]) and torch.equal(a2[k], b2[k]), k                 # same seed, same run
        assert not torch.equal(a1[k], a2[k]), k                                           # a new draw every step
    z = a1['z_vals_coarse']
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and float(z.min()) >= 0 and float(z.max()) <= 1
    assert launches <= 2 * 20, launches     # per forward: sampler, 3 view-bias + 4 packs + 4 chain kernels, 4 composites, resampler
