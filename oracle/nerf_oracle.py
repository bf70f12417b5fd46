"""CPU oracle for the SimpleNeRF volumetric-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``simplenerf_b200/`` may import this module; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` as the *checker* and as the timed CPU arm, never as the product.

It is a functional restatement (plain torch fp32 on CPU) of the algorithm in the reference file
``src/models/SimpleNeRF01.py``; each function cites the reference lines it follows.  The torch
op *order* is kept identical to the reference wherever floating-point association matters, so
that on the same host the outputs are bit-identical to the reference's CPU outputs.

Parity pinning: the reference ships no tests / golden vectors for this path, so
the oracle is pinned against outputs of the unmodified reference module itself, generated in the
build container by ``oracle/make_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` replays them).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# a3  stratified coarse sampling        (reference src/models/SimpleNeRF01.py:272-302)
# --------------------------------------------------------------------------------------------


def stratified_z(near: torch.Tensor, far: torch.Tensor, num_samples: int, lindisp: bool = False,
                 t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """near/far [N,1] -> z [N,num_samples].  ``t_rand`` [N,num_samples] in [0,1) enables the
    perturbed (training) branch (:293-301); ``None`` is the deterministic branch."""
    steps = torch.linspace(0., 1., steps=num_samples)
    if lindisp:
        z = 1. / (1. / near * (1. - steps) + 1. / far * steps)          # :289
    else:
        z = near * (1. - steps) + far * steps                           # :287
    z = z.expand([near.shape[0], num_samples])                          # :291
    if t_rand is not None:
        centre = .5 * (z[..., 1:] + z[..., :-1])                        # :295
        hi = torch.cat([centre, z[..., -1:]], -1)                       # :296
        lo = torch.cat([z[..., :1], centre], -1)                        # :297
        z = lo + (hi - lo) * t_rand                                     # :301
    return z


# --------------------------------------------------------------------------------------------
# a6  positional encoding               (reference :525-557, configured at :612-624)
# --------------------------------------------------------------------------------------------


def positional_encoding(x: torch.Tensor, degree: int) -> torch.Tensor:
    """[P,3] -> [P, 3*(1+2*degree)]: x, then per band k<degree: sin(x*2^k), cos(x*2^k)."""
    bands = 2. ** torch.linspace(0., degree - 1, steps=degree)          # :544 (log sampling)
    feats = [x]
    for k in range(degree):
        feats.append(torch.sin(x * bands[k]))                           # :550 (sin first)
        feats.append(torch.cos(x * bands[k]))
    return torch.cat(feats, -1)                                         # :557


# --------------------------------------------------------------------------------------------
# a7/a8  the NeRF MLP and its variants  (reference :560-715)
# --------------------------------------------------------------------------------------------


class MlpSpec:
    """Shape bookkeeping of one reference ``MLP`` (ctor :561-609)."""

    def __init__(self, mlp_cfg: dict):
        self.depth = mlp_cfg['points_net_depth']
        self.width = mlp_cfg['points_net_width']
        self.view_depth = mlp_cfg['views_net_depth']
        self.view_width = mlp_cfg['views_net_width']
        self.pts_degree = mlp_cfg['points_positional_encoding_degree']
        self.use_view_dirs = mlp_cfg['use_view_dirs']
        self.view_dep_rgb = mlp_cfg['view_dependent_rgb']
        self.predict_visibility = mlp_cfg['predict_visibility']
        self.pts_enc_dim = (2 * self.pts_degree + 1) * 3
        self.trunk_in = self.pts_enc_dim
        self.view_in = 0
        self.view_degree = 0
        if self.use_view_dirs:
            self.view_degree = mlp_cfg['views_positional_encoding_degree']
            self.view_in = (2 * self.view_degree + 1) * 3
        if 'points_sigma_positional_encoding_degree' in mlp_cfg:         # :576-578
            self.trunk_in = (2 * mlp_cfg['points_sigma_positional_encoding_degree'] + 1) * 3
            self.view_in += self.pts_enc_dim - self.trunk_in
        self.skips = (4,)                                                # :580
        self.has_view_branch = self.view_dep_rgb or self.predict_visibility
        self.head_out = 1 + (0 if self.view_dep_rgb else 3)              # :596-601
        self.view_head_out = (3 if self.view_dep_rgb else 0) + (1 if self.predict_visibility else 0)

    def param_shapes(self) -> Dict[str, tuple]:
        """state_dict names/shapes, identical to the reference module's (:586-608)."""
        shapes = {}
        for i in range(self.depth):
            fan_in = self.trunk_in if i == 0 else self.width + (self.trunk_in if (i - 1) in self.skips else 0)
            shapes[f'pts_linears.{i}.weight'] = (self.width, fan_in)
            shapes[f'pts_linears.{i}.bias'] = (self.width,)
        if self.has_view_branch:
            for i in range(self.view_depth):
                fan_in = self.view_in + self.width if i == 0 else self.view_width
                shapes[f'views_linears.{i}.weight'] = (self.view_width, fan_in)
                shapes[f'views_linears.{i}.bias'] = (self.view_width,)
        shapes['pts_output_linear.weight'] = (self.head_out, self.width)
        shapes['pts_output_linear.bias'] = (self.head_out,)
        if self.has_view_branch:
            shapes['feature_linear.weight'] = (self.width, self.width)
            shapes['feature_linear.bias'] = (self.width,)
            shapes['views_output_linear.weight'] = (self.view_head_out, self.view_width)
            shapes['views_output_linear.bias'] = (self.view_head_out,)
        return shapes


def mlp_forward(spec: MlpSpec, params: Dict[str, torch.Tensor], pts: torch.Tensor,
                view_dirs: Optional[torch.Tensor], sigma_noise: Optional[torch.Tensor] = None,
                view_dirs2: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """pts [P,3], view_dirs [P,3] (already expanded per point, :375) -> sigma [P,1], rgb [P,3]; with
    ``predict_visibility`` also visibility [P,1] and, given view_dirs2 [P,nf-1,3] (the directions from the other
    views' cameras to the point, :646-649), visibility2 [P,nf-1,1] (row a14 / N4).

    ``sigma_noise`` [P,1] is ``randn * raw_noise_std`` (:670), ``None`` in eval.
    Follows MLP.forward :626-654, trunk :656-685, view branch :687-715.
    """
    enc = positional_encoding(pts, spec.pts_degree)                       # :630
    trunk_in = enc[:, :spec.trunk_in]                                     # :631
    h = trunk_in
    for i in range(spec.depth):                                           # :659-663
        h = F.linear(h, params[f'pts_linears.{i}.weight'], params[f'pts_linears.{i}.bias'])
        h = F.relu(h)
        if i in spec.skips:
            h = torch.cat([trunk_in, h], -1)
    head = F.linear(h, params['pts_output_linear.weight'], params['pts_output_linear.bias'])  # :665
    sigma = head[..., 0:1]
    if sigma_noise is not None:                                           # :669-671
        sigma = sigma + sigma_noise
    sigma = F.relu(sigma)                                                 # :672
    out = {'sigma': sigma}
    if not spec.view_dep_rgb:
        out['rgb_view_independent'] = torch.sigmoid(head[..., 1:4])       # :676-680
        out['rgb'] = out['rgb_view_independent']
    if spec.has_view_branch:
        feat = F.linear(h, params['feature_linear.weight'], params['feature_linear.bias'])     # :683
        feat = torch.cat([feat, enc[:, spec.trunk_in:]], dim=1)           # :633
        venc = positional_encoding(view_dirs, spec.view_degree)           # :640
        hv = torch.cat([feat, venc], -1)                                  # :695
        for i in range(spec.view_depth):                                  # :697-699
            hv = F.relu(F.linear(hv, params[f'views_linears.{i}.weight'], params[f'views_linears.{i}.bias']))
        vout = F.linear(hv, params['views_output_linear.weight'], params['views_output_linear.bias'])  # :701
        ch = 0
        if spec.view_dep_rgb:
            out['rgb_view_dependent'] = torch.sigmoid(vout[..., 0:3])     # :704-707
            out['rgb'] = out['rgb_view_dependent']
            ch = 3
        if spec.predict_visibility:
            out['visibility'] = torch.sigmoid(vout[..., ch:ch + 1])       # :710-713
            if view_dirs2 is not None:                                    # :646-649: the view branch again, per other view
                venc2 = positional_encoding(view_dirs2, spec.view_degree)
                hv2 = torch.cat([feat[:, None, :].repeat([1, view_dirs2.shape[1], 1]), venc2], -1)     # :691-695
                for i in range(spec.view_depth):
                    hv2 = F.relu(F.linear(hv2, params[f'views_linears.{i}.weight'], params[f'views_linears.{i}.bias']))
                vout2 = F.linear(hv2, params['views_output_linear.weight'], params['views_output_linear.bias'])
                out['visibility2'] = torch.sigmoid(vout2[..., ch:ch + 1])
    return out


def other_view_dirs(z: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, rays_o2: torch.Tensor, ndc: bool) -> torch.Tensor:
    """compute_other_view_dirs (:317-325): unit directions from the other views' camera centres rays_o2 [N,nf-1,3] to the
    sample points; z [N,S] is NDC depth when ``ndc`` (converted with near hard-coded to 1 and a 1e-6 guard, :319-321)."""
    if ndc:
        tn = -(1 + rays_o[..., 2]) / rays_d[..., 2]
        z = (((rays_o[..., None, 2] + tn[..., None] * rays_d[..., None, 2]) / (1 - z + 1e-6)) - rays_o[..., None, 2]) / rays_d[..., None, 2]
    pts = rays_o[..., None, :] + z[..., None] * rays_d[..., None, :]
    dirs = pts[:, :, None] - rays_o2[..., None, :, :]                     # (N, S, nf-1, 3)
    return dirs / torch.norm(dirs, dim=-1, keepdim=True)


def other_view_origins(poses: torch.Tensor, pixel_id: torch.Tensor, num_frames: int) -> torch.Tensor:
    """render_rays :120-133 when the batch carries no rays_o2: camera centres of the num_frames-1 other views, in view
    order with the ray's own view skipped."""
    image_id = pixel_id[:, 0].long()
    cols = []
    for i in range(num_frames - 1):
        other = i + (i >= image_id).long()
        cols.append(poses[other][:, :3, 3])
    return torch.stack(cols, dim=1)


# --------------------------------------------------------------------------------------------
# a10  NDC depth -> metric depth        (reference :485-502, static)
# --------------------------------------------------------------------------------------------


def ndc_to_metric_depth(z_ndc: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor) -> torch.Tensor:
    oz = rays_o[..., 2:3]
    dz = rays_d[..., 2:3]
    t_near = -(1 + oz) / dz                                               # :498 (near hard-coded 1)
    guard = torch.where(z_ndc == 1., 1e-3, 0.)                            # :499
    return (oz + t_near * dz) / dz * (1 / (1 - z_ndc + guard) - 1) + t_near   # :501


# --------------------------------------------------------------------------------------------
# a9  alpha compositing                 (reference :430-483)
# --------------------------------------------------------------------------------------------


def composite(sigma: torch.Tensor, rgb: torch.Tensor, z: torch.Tensor, ndc: bool,
              rays_o: torch.Tensor, rays_d: torch.Tensor, rays_d_ndc: Optional[torch.Tensor] = None,
              white_bkgd: bool = False, visibility2: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """sigma [N,S], rgb [N,S,3], z [N,S] (z is NDC depth when ``ndc``) -> per-ray maps."""
    n = z.shape[0]
    if ndc:
        tail = torch.Tensor([1]).expand(z[..., :1].shape)                 # :438-439
        scale = torch.norm(rays_d_ndc[..., None, :], dim=-1)              # :441
    else:
        tail = torch.Tensor([1e10]).expand(z[..., :1].shape)              # :433-434
        scale = torch.norm(rays_d[..., None, :], dim=-1)                  # :436
    z1 = torch.cat([z, tail], -1)
    delta = (z1[..., 1:] - z1[..., :-1]) * scale

    alpha = 1. - torch.exp(-sigma * delta)                                # :446
    trans = torch.cumprod(torch.cat([torch.ones((n, 1)), 1. - alpha + 1e-10], -1), -1)[:, :-1]  # :447
    weights = alpha * trans                                               # :448
    rgb_map = torch.sum(weights[..., None] * rgb, dim=-2)                 # :449
    acc = torch.sum(weights, dim=-1)                                      # :451
    out = {}
    if ndc:
        depth_ndc = torch.sum(weights * z, dim=-1) / (acc + 1e-6)         # :456
        out['depth_ndc'] = depth_ndc
        out['depth_var_ndc'] = torch.sum(weights * torch.square(z - depth_ndc[..., None]), dim=-1)  # :457
        z = ndc_to_metric_depth(z, rays_o, rays_d)                        # :458
    depth = torch.sum(weights * z, dim=-1) / (acc + 1e-6)                 # :453/:459
    depth_var = torch.sum(weights * torch.square(z - depth[..., None]), dim=-1)   # :454/:460
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc[..., None])                         # :463
    out.update({'rgb': rgb_map, 'acc': acc, 'alpha': alpha, 'visibility': trans, 'weights': weights,
                'depth': depth, 'depth_var': depth_var})
    if visibility2 is not None:                                           # :479-482, [N,S,nf-1,1] -> [N,nf-1]
        out['visibility2'] = torch.sum(weights[..., None] * visibility2[..., 0], dim=-2) / (acc[..., None] + 1e-6)
    return out


# --------------------------------------------------------------------------------------------
# a12  inverse-CDF resampling           (reference :328-361, static)
# --------------------------------------------------------------------------------------------


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, num_new: int,
               u: Optional[torch.Tensor] = None, return_debug: bool = False):
    """bins [N,B], weights [N,B-1] -> [N,num_new].  ``u`` None = deterministic linspace (:338)."""
    w = weights + 1e-5                                                    # :331
    pdf = w / torch.sum(w, -1, keepdim=True)                              # :332
    cdf = torch.cumsum(pdf, -1)                                           # :333
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)            # :334
    if u is None:
        u = torch.linspace(0., 1., steps=num_new).expand(list(cdf.shape[:-1]) + [num_new])
    u = u.contiguous()
    idx = torch.searchsorted(cdf, u, right=True)                          # :345
    lo = torch.clamp(idx - 1, min=0)                                      # :346
    hi = torch.clamp(idx, max=cdf.shape[-1] - 1)                          # :347
    cdf_lo, cdf_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)       # :353
    bin_lo, bin_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)     # :354
    span = cdf_hi - cdf_lo                                                # :356
    span = torch.where(span < 1e-5, torch.ones_like(span), span)          # :357
    frac = (u - cdf_lo) / span                                            # :358
    samples = bin_lo + frac * (bin_hi - bin_lo)                           # :359
    if return_debug:
        return samples, {'cdf': cdf, 'below': lo, 'above': hi}
    return samples


# --------------------------------------------------------------------------------------------
# a11  fine-pass depths                 (reference :304-315)
# --------------------------------------------------------------------------------------------


def fine_z(z_coarse: torch.Tensor, weights_coarse: torch.Tensor, num_new: int,
           u: Optional[torch.Tensor] = None) -> torch.Tensor:
    mids = .5 * (z_coarse[..., 1:] + z_coarse[..., :-1])                  # :310
    extra = sample_pdf(mids, weights_coarse[..., 1:-1], num_new, u=u).detach()   # :311-312
    merged, _ = torch.sort(torch.cat([z_coarse, extra], -1), -1)          # :314
    return merged


# --------------------------------------------------------------------------------------------
# a1/a2/a5/a13  the renderer            (reference :11-270, 363-428)
# --------------------------------------------------------------------------------------------

_MODEL_SLOTS = (          # attribute name (:22-41)      path into configs['model']
    ('coarse_model', ('coarse_mlp',)),
    ('fine_model', ('fine_mlp',)),
    ('pts_aug_coarse_model', ('points_augmentation', 'coarse_mlp')),
    ('pts_aug_fine_model', ('points_augmentation', 'fine_mlp')),
    ('views_aug_coarse_model', ('views_augmentation', 'coarse_mlp')),
    ('views_aug_fine_model', ('views_augmentation', 'fine_mlp')),
)


def model_slots(configs: dict) -> Dict[str, dict]:
    """slot name -> its mlp config, in the reference's construction order (build_nerf :45-65)."""
    slots = {}
    for name, path in _MODEL_SLOTS:
        node = configs['model']
        for key in path:
            node = node.get(key) if isinstance(node, dict) else None
            if node is None:
                break
        if node is not None:
            slots[name] = node
    return slots


class Randoms:
    """Source of the three kinds of random numbers the reference draws on the CPU generator
    (:299 t_rand, :341 u, :670 sigma noise).  The default draws them from torch's global CPU
    generator in the reference's order and slice sizes; tests inject recorded tensors."""

    def __init__(self, netchunk: Optional[int]):
        self.netchunk = netchunk

    def t_rand(self, n: int, s: int) -> torch.Tensor:
        return torch.rand([n, s])

    def u(self, n: int, s: int) -> torch.Tensor:
        return torch.rand([n, s])

    def sigma_noise(self, tag: str, p: int) -> torch.Tensor:
        step = self.netchunk or p
        return torch.cat([torch.randn([min(step, p - i), 1]) for i in range(0, p, step)], 0)


class FixedRandoms(Randoms):
    """Replays recorded draws: {'t_rand': [N,Sc], 'u': [N,Nf], 'noise_<slot>': [N*S,1]}."""

    def __init__(self, table: Dict[str, torch.Tensor]):
        super().__init__(None)
        self.table = table

    def t_rand(self, n, s):
        return self.table['t_rand'][:n, :s]

    def u(self, n, s):
        return self.table['u'][:n, :s]

    def sigma_noise(self, tag, p):
        return self.table[f'noise_{tag}'][:p]


class NerfOracle(torch.nn.Module):
    """Same constructor / forward contract and the same state_dict names as the reference
    ``SimpleNeRF`` (:11-75), computing with the functions above."""

    def __init__(self, configs: dict, model_configs: Optional[dict] = None):
        super().__init__()
        self.configs = configs
        self.model_configs = model_configs
        self.ndc = configs['data_loader']['ndc']
        self.specs: Dict[str, MlpSpec] = {}
        for slot, mlp_cfg in model_slots(configs).items():
            spec = MlpSpec(mlp_cfg)
            self.specs[slot] = spec
            holder = torch.nn.Module()
            # torch.nn.Linear gives the reference's default init *and* parameter names
            holder.pts_linears = torch.nn.ModuleList(
                [torch.nn.Linear(*reversed(spec.param_shapes()[f'pts_linears.{i}.weight'])) for i in range(spec.depth)])
            if spec.has_view_branch:
                holder.views_linears = torch.nn.ModuleList(
                    [torch.nn.Linear(*reversed(spec.param_shapes()[f'views_linears.{i}.weight']))
                     for i in range(spec.view_depth)])
            holder.pts_output_linear = torch.nn.Linear(spec.width, spec.head_out)
            if spec.has_view_branch:
                holder.feature_linear = torch.nn.Linear(spec.width, spec.width)
                holder.views_output_linear = torch.nn.Linear(spec.view_width, spec.view_head_out)
            setattr(self, slot, holder)
        self.randoms: Optional[Randoms] = None
        self.mlp_impl: Callable = mlp_forward   # tests may swap in oracle.bf16_emulation.mlp_forward_bf16

    # -- helpers ----------------------------------------------------------------------------
    def _params(self, slot: str) -> Dict[str, torch.Tensor]:
        return dict(getattr(self, slot).named_parameters())

    def _run_mlp(self, slot: str, pts: torch.Tensor, view_dirs: Optional[torch.Tensor],
                 rnd: Randoms, view_dirs2: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """run_network + batchify (:363-428): flatten, expand view dirs per point, reshape back."""
        spec = self.specs[slot]
        n, s = pts.shape[:2]
        flat = pts.reshape(-1, 3)
        vd = None
        if spec.use_view_dirs:
            vd = view_dirs[:, None].expand(pts.shape).reshape(-1, 3)     # :375-376
        vd2 = None
        if spec.use_view_dirs and spec.predict_visibility and view_dirs2 is not None:         # :379-382
            vd2 = view_dirs2.reshape(-1, view_dirs2.shape[-2], 3)
        noise = None
        if self.training and self.configs['model']['raw_noise_std'] > 0.:
            noise = rnd.sigma_noise(slot, n * s) * self.configs['model']['raw_noise_std']
        chunk = self.configs['model'].get('netchunk') or flat.shape[0]
        pieces: Dict[str, List[torch.Tensor]] = {}
        params = self._params(slot)
        for i in range(0, flat.shape[0], chunk):                          # :404
            extra = {} if vd2 is None else {'view_dirs2': vd2[i:i + chunk]}
            part = self.mlp_impl(spec, params, flat[i:i + chunk], None if vd is None else vd[i:i + chunk],
                               None if noise is None else noise[i:i + chunk], **extra)
            for k, v in part.items():
                pieces.setdefault(k, []).append(v)
        return {k: torch.cat(v, 0).reshape([n, s] + list(v[0].shape[1:])) for k, v in pieces.items()}    # :387-389

    def _stream(self, out: dict, prefix: str, suffix: str, slot: str, pts, view_dirs, z, batch, rnd, retraw, rays_o2=None):
        vd2 = None
        if rays_o2 is not None and self.specs[slot].predict_visibility:   # :149-151 / :214-216
            vd2 = other_view_dirs(z, batch['rays_o'], batch['rays_d'], rays_o2, self.ndc)
        raw = self._run_mlp(slot, pts, view_dirs, rnd, vd2)
        maps = composite(raw['sigma'][..., 0], raw['rgb'], z, self.ndc, batch['rays_o'], batch['rays_d'],
                         batch.get('rays_d_ndc'), self.configs['model']['white_bkgd'],
                         raw.get('visibility2') if rays_o2 is not None else None)
        for k, v in maps.items():
            out[f'{prefix}{k}_{suffix}'] = v
        if retraw:
            for k, v in raw.items():
                out[f'{prefix}raw_{k}_{suffix}'] = v
        return maps

    # -- a13 --------------------------------------------------------------------------------
    def render_rays(self, batch: dict, retraw: bool, rnd: Randoms, sec_views_vis: bool = False) -> Dict[str, torch.Tensor]:
        cfg = self.configs['model']
        n = batch['rays_o'].shape[0]
        rays_o2 = None
        if sec_views_vis and any(spec.predict_visibility for k, spec in self.specs.items() if k in ('coarse_model', 'fine_model')):   # :19-20, :119
            rays_o2 = batch['rays_o2'] if 'rays_o2' in batch else other_view_origins(
                batch['common_data']['poses'], batch['pixel_id'], batch['num_frames'])
        o, d = (batch['rays_o_ndc'], batch['rays_d_ndc']) if self.ndc else (batch['rays_o'], batch['rays_d'])
        near, far = (batch['near_ndc'], batch['far_ndc']) if self.ndc else (batch['near'], batch['far'])
        view_dirs = batch.get('view_dirs')
        perturb = bool(cfg['perturb']) and self.training
        out: Dict[str, torch.Tensor] = {}

        s_c = cfg['coarse_mlp']['num_samples']
        z_c = stratified_z(near, far, s_c, cfg['lindisp'], rnd.t_rand(n, s_c) if perturb else None)
        pts_c = o[..., None, :] + d[..., None, :] * z_c[..., :, None]     # :140/:142
        out['z_vals_coarse'] = z_c
        maps_c = self._stream(out, '', 'coarse', 'coarse_model', pts_c, view_dirs, z_c, batch, rnd, retraw, rays_o2)
        if self.training and 'pts_aug_coarse_model' in self.specs:        # :170
            self._stream(out, 'points_augmentation_', 'coarse', 'pts_aug_coarse_model', pts_c, view_dirs, z_c,
                         batch, rnd, retraw)
        if self.training and 'views_aug_coarse_model' in self.specs:      # :186
            self._stream(out, 'views_augmentation_', 'coarse', 'views_aug_coarse_model', pts_c, view_dirs, z_c,
                         batch, rnd, retraw)

        if 'fine_model' in self.specs:
            n_f = cfg['fine_mlp']['num_samples']
            z_f = fine_z(z_c, maps_c['weights'], n_f, rnd.u(n, n_f) if perturb else None)    # :202
            pts_f = o[..., None, :] + d[..., None, :] * z_f[..., :, None]
            out['z_vals_fine'] = z_f
            self._stream(out, '', 'fine', 'fine_model', pts_f, view_dirs, z_f, batch, rnd, retraw, rays_o2)
            if self.training and 'pts_aug_fine_model' in self.specs:      # :234
                self._stream(out, 'points_augmentation_', 'fine', 'pts_aug_fine_model', pts_f, view_dirs, z_f,
                             batch, rnd, retraw)
            if self.training and 'views_aug_fine_model' in self.specs:    # :249
                self._stream(out, 'views_augmentation_', 'fine', 'views_aug_fine_model', pts_f, view_dirs, z_f,
                             batch, rnd, retraw)
        if not retraw:                                                    # :265-269
            for level in ('coarse', 'fine'):
                for k in ('z_vals', 'visibility', 'weights'):
                    out.pop(f'{k}_{level}', None)
        return out

    # -- a1/a2 ------------------------------------------------------------------------------
    def forward(self, input_batch: dict, retraw: bool = False, sec_views_vis: bool = False):
        rnd = self.randoms or Randoms(self.configs['model'].get('netchunk'))
        retraw = retraw or self.training                                  # :74
        sec_views_vis = sec_views_vis or self.training
        if 'common_data' in input_batch:                                  # :69-73 (on a copy: the caller's dict stays as it was)
            input_batch = dict(input_batch, common_data={k: (v[0] if isinstance(v, torch.Tensor) else v)
                                                         for k, v in input_batch['common_data'].items()})
        n = input_batch['rays_o'].shape[0]
        chunk = self.configs['model']['chunk']
        parts: Dict[str, List[torch.Tensor]] = {}
        for i in range(0, n, chunk):                                      # :88
            sub = {k: (v[i:i + chunk] if isinstance(v, torch.Tensor) and v.shape[:1] == (n,) else v)
                   for k, v in input_batch.items()}
            for k, v in self.render_rays(sub, retraw, rnd, sec_views_vis).items():
                parts.setdefault(k, []).append(v)
        return {k: torch.cat(v, 0) for k, v in parts.items()}             # :504-512


def deterministic_state(shapes: Dict[str, tuple], seed: int) -> Dict[str, torch.Tensor]:
    """Host-independent stand-in for nn.Linear's default init (shared with the synthetic-input module)."""
    from simplenerf_b200.synthetic import deterministic_state as _det
    return _det(shapes, seed)
