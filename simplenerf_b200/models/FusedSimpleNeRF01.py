"""B200-native drop-in for the reference model ``models/SimpleNeRF01.py`` (class ``SimpleNeRF``).

Same plug-in contract as the reference (``ModelFactory.get_model`` picks the class named like the
file minus its two-digit suffix, ``src/models/ModelFactory.py:10-22``):

* ``FusedSimpleNeRF(configs, model_configs)`` -- same config dictionaries (``SimpleNeRF01.py:12``);
* ``forward(input_batch, retraw=False, sec_views_vis=False) -> dict`` -- same ray-batch input dict and
  the same output keys (``:67-75``, ``:163-269``);
* identical ``state_dict`` names / shapes, so reference checkpoints load unchanged.

Everything between the two dictionaries runs in the hand-written sm_100a kernels of
``libsimplenerf_b200.so`` (C ABI in ``include/simplenerf_b200.h``): there is no PyTorch or CPU
fallback -- a missing library or a CPU tensor raises.

Extra, optional keys read from ``configs['model']`` (absent in reference configs):
``precision``: ``'bf16'`` (tcgen05 tensor path, default) or ``'fp32'`` (CUDA-core precise path);
``rng``: ``'device'`` (default: the kernels that consume a random number draw it themselves -- counter-based Philox keyed by
``torch.initial_seed()``, nothing is materialised), ``'torch'`` (torch.rand / randn tensors on the GPU) or ``'reference'`` (draws
on the CPU generator in the reference's order, so equal seeds give equal random numbers, ``:299, :341, :670``);
``launch_rays``: rays per kernel launch group (default 65536; bounds workspace memory).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import ops
from .._lib import (FLAG_NDC, FLAG_PRECISE, FLAG_SAVE_FOR_BWD, FLAG_VIS_GRAD, FLAG_VIS_HEAD, FLAG_WHITE_BKGD, MlpDesc, P_COUNT, P_FEAT_B, P_FEAT_W,
                    P_HEAD_B, P_HEAD_W, P_RGB_B, P_RGB_W, P_VIEW_B, P_VIEW_W)

_SLOTS = (   # attribute (reference :22-41), path in configs['model'], output-key prefix, level
    ('coarse_model', ('coarse_mlp',), '', 'coarse'),
    ('pts_aug_coarse_model', ('points_augmentation', 'coarse_mlp'), 'points_augmentation_', 'coarse'),
    ('views_aug_coarse_model', ('views_augmentation', 'coarse_mlp'), 'views_augmentation_', 'coarse'),
    ('fine_model', ('fine_mlp',), '', 'fine'),
    ('pts_aug_fine_model', ('points_augmentation', 'fine_mlp'), 'points_augmentation_', 'fine'),
    ('views_aug_fine_model', ('views_augmentation', 'fine_mlp'), 'views_augmentation_', 'fine'),
)
_CTOR_ORDER = ('coarse_model', 'fine_model', 'pts_aug_coarse_model', 'pts_aug_fine_model',
               'views_aug_coarse_model', 'views_aug_fine_model')   # build_nerf :45-65 (keeps default init identical)


class MlpBlock(torch.nn.Module):
    """Parameter holder with the reference ``MLP``'s names, shapes and default init (:561-609)."""

    def __init__(self, mlp_cfg: dict):
        super().__init__()
        c = mlp_cfg
        self.predict_visibility = bool(c['predict_visibility'])         # row a14 / N4: fourth row of the view head (:602-603)
        if self.predict_visibility and not (c['use_view_dirs'] and c['view_dependent_rgb']):
            raise NotImplementedError('predict_visibility=True needs use_view_dirs and view_dependent_rgb (as the reference configs pair them)')
        if c['views_net_depth'] != 1:
            raise NotImplementedError('views_net_depth != 1 is not built')
        self.width, self.depth = c['points_net_width'], c['points_net_depth']
        self.pts_degree = c['points_positional_encoding_degree']
        self.trunk_degree = c.get('points_sigma_positional_encoding_degree', self.pts_degree)
        self.use_view_dirs = c['use_view_dirs']
        self.view_degree = c['views_positional_encoding_degree'] if self.use_view_dirs else 0
        self.has_view = bool(c['view_dependent_rgb'])
        if self.has_view != self.use_view_dirs:
            raise NotImplementedError('use_view_dirs and view_dependent_rgb must agree (as in the shipped configs)')
        self.view_width = c['views_net_width'] if self.has_view else 0
        enc, trunk_in = 3 * (1 + 2 * self.pts_degree), 3 * (1 + 2 * self.trunk_degree)
        view_in = self.width + (enc - trunk_in) + (3 * (1 + 2 * self.view_degree) if self.view_degree else 0)
        skips = (4,)
        self.pts_linears = torch.nn.ModuleList(
            [torch.nn.Linear(trunk_in, self.width)] +
            [torch.nn.Linear(self.width + (trunk_in if i in skips else 0), self.width) for i in range(self.depth - 1)])
        if self.has_view:
            self.views_linears = torch.nn.ModuleList([torch.nn.Linear(view_in, self.view_width)])
        self.pts_output_linear = torch.nn.Linear(self.width, 1 if self.has_view else 4)
        if self.has_view:
            self.feature_linear = torch.nn.Linear(self.width, self.width)
            self.views_output_linear = torch.nn.Linear(self.view_width, 3 + int(self.predict_visibility))
        self.desc = MlpDesc(depth=self.depth, width=self.width, skip_layer=4, pts_degree=self.pts_degree,
                            trunk_degree=self.trunk_degree, view_degree=self.view_degree, view_width=self.view_width,
                            head_out=1 if self.has_view else 4)
        self._packed: Optional[torch.Tensor] = None
        self._packed_key = None

    def param_table(self) -> List[Optional[torch.nn.Parameter]]:
        t: List[Optional[torch.nn.Parameter]] = [None] * P_COUNT
        for i, lin in enumerate(self.pts_linears):
            t[2 * i], t[2 * i + 1] = lin.weight, lin.bias
        t[P_HEAD_W], t[P_HEAD_B] = self.pts_output_linear.weight, self.pts_output_linear.bias
        if self.has_view:
            t[P_FEAT_W], t[P_FEAT_B] = self.feature_linear.weight, self.feature_linear.bias
            t[P_VIEW_W], t[P_VIEW_B] = self.views_linears[0].weight, self.views_linears[0].bias
            t[P_RGB_W], t[P_RGB_B] = self.views_output_linear.weight, self.views_output_linear.bias
        return t

    def packed(self, params: List[Optional[torch.Tensor]], force: bool = False) -> torch.Tensor:
        """bf16 weight image for the tensor path; a derived cache of the fp32 parameters.  It is rebuilt on every
        training forward (``force``: fused optimizers update parameters without bumping ``_version``) and, in eval,
        whenever a parameter's storage or version changed or the module switched between train() and eval()."""
        key = tuple((p.data_ptr(), p._version) for p in params if p is not None) + (self.training,)
        if force or self._packed is None or key != self._packed_key:
            dev = params[0].device
            if self._packed is None or self._packed.device != dev:
                self._packed = torch.empty(ops.packed_weights_bytes(self.desc), dtype=torch.uint8, device=dev)
            ops.pack_weights(self.desc, [None if p is None else p.detach() for p in params], self._packed)
            self._packed_key = key
        return self._packed


class DeviceRandoms:
    """Production random source: draws directly on the GPU."""

    def t_rand(self, n, s, device):
        return torch.rand((n, s), device=device)

    def u(self, n, s, device):
        return torch.rand((n, s), device=device)

    def sigma_noise(self, tag, p, device):
        return torch.randn(p, device=device)


class KernelRandoms(DeviceRandoms):
    """Production random source: every draw is an `ops.RngDraw` -- a (seed, offset) pair handed to the
    kernel that consumes the numbers (stratified sampler, resampler, sigma-head epilogue).  The seed is torch's
    (`torch.manual_seed` makes runs repeatable), the offset counts the draws of this model."""

    def __init__(self):
        self.seed = None
        self.draws = 0

    def _draw(self, scale: float = 1.0) -> 'ops.RngDraw':
        if self.seed is None:
            self.seed = torch.initial_seed()
        self.draws += 1
        return ops.RngDraw(self.seed, self.draws, scale)

    def t_rand(self, n, s, device):
        return self._draw()

    def u(self, n, s, device):
        return (self._draw(), s)

    def sigma_noise(self, tag, p, device):
        return self._draw()


class ReferenceOrderRandoms(DeviceRandoms):
    """Draws on torch's CPU generator with the reference's call order and slice sizes
    (t_rand :299; one randn per netchunk slice :670; u :341), then copies to the device."""

    def __init__(self, netchunk):
        self.netchunk = netchunk

    def t_rand(self, n, s, device):
        return torch.rand([n, s]).to(device)

    def u(self, n, s, device):
        return torch.rand([n, s]).to(device)

    def sigma_noise(self, tag, p, device):
        step = self.netchunk or p
        return torch.cat([torch.randn([min(step, p - i), 1]) for i in range(0, p, step)], 0).reshape(-1).to(device)


class FixedRandoms(DeviceRandoms):
    """Replays given draws (tests): {'t_rand': [N,Sc], 'u': [N,Nf], 'noise_<slot>': [N*S,1]}.  ``offset`` rays
    are skipped so that chunked launches read their own rows."""

    def __init__(self, table: Dict[str, torch.Tensor]):
        self.table = table
        self.offset = 0

    def t_rand(self, n, s, device):
        return self.table['t_rand'][self.offset:self.offset + n, :s].to(device)

    def u(self, n, s, device):
        return self.table['u'][self.offset:self.offset + n, :s].to(device)

    def sigma_noise(self, tag, p, device):
        s = p // max(1, self._n)
        return self.table[f'noise_{tag}'].reshape(-1)[self.offset * s:self.offset * s + p].to(device)

    _n = 1


class _RenderStream(torch.autograd.Function):
    """One MLP evaluated on [N,S] samples followed by compositing (run_network + volume_rendering,
    reference :363-483).  Gradients flow to the MLP parameters only (rays, depths and noise are data)."""

    @staticmethod
    def forward(ctx, block: MlpBlock, opts: dict, z, noise, rays_o, rays_d, pts_o, pts_d, view_dirs, rays_o2, *params):
        ndc, white = opts['ndc'], opts['white_bkgd']
        need_grad = opts['need_grad']
        flags = (FLAG_PRECISE if opts['precise'] else 0) | (FLAG_SAVE_FOR_BWD if need_grad else 0)
        if block.predict_visibility and not opts['precise']:
            flags |= FLAG_VIS_HEAD            # tensor path: the forward kernel also leaves what the visibility head reads (vis_tc.cu)
        n, s = z.shape
        table: List[Optional[torch.Tensor]] = [None] * P_COUNT
        it = iter(params)
        for i, slot in enumerate(opts['param_mask']):
            if slot:
                table[i] = next(it)
        packed = None if opts['precise'] else block.packed(table, force=need_grad)
        ws = torch.empty(ops.mlp_workspace_bytes(block.desc, n, s, flags), dtype=torch.uint8, device=z.device)
        sigma, rgb = ops.mlp_forward(block.desc, table, packed, pts_o, pts_d, view_dirs if block.view_degree else None,
                                     z, noise, ws, flags)
        maps = ops.composite_forward(sigma, rgb, z, rays_o, rays_d, pts_d if ndc else None, ndc, white)
        keys = [k for k in ('rgb', 'acc', 'depth', 'depth_var', 'depth_ndc', 'depth_var_ndc', 'alpha', 'visibility',
                            'weights') if k in maps]
        # row a14 / N4: secondary-view visibility head (:149-151, :379-382, :479-482, :640-649)
        vis = vis2 = vws = None
        vflags = flags | (FLAG_NDC if ndc else 0)
        if block.predict_visibility:
            vis, vis2, vws = ops.visibility_forward(block.desc, table, ws, rays_o, rays_d, z, rays_o2, vflags)
            if vis2 is not None:
                maps['visibility2'] = ops.visibility2_composite_forward(maps['weights'], maps['acc'], vis2)
                keys.append('visibility2')
        ctx.keys = keys
        if need_grad:
            ctx.block, ctx.opts, ctx.flags, ctx.ws, ctx.packed = block, opts, flags, ws, packed
            ctx.vis_state = (vis, vis2, vws, vflags, rays_o2, maps['weights'], maps['acc'], maps.get('visibility2'))
            ctx.save_for_backward(z, sigma, rgb, rays_o, rays_d, pts_o, pts_d, view_dirs, *params)
        ctx.set_materialize_grads(False)
        extra = tuple(t for t in (vis, vis2) if t is not None)     # network outputs of the head: differentiable (VisibilityLoss01.py:29)
        ctx.n_extra = len(extra)
        ctx.mark_non_differentiable(sigma, rgb)
        return tuple(maps[k] for k in keys) + (sigma, rgb) + extra

    @staticmethod
    def backward(ctx, *gouts):
        with torch.cuda.device(ctx.saved_tensors[0].device):       # autograd's thread may sit on another device
            return _RenderStream._backward(ctx, *gouts)

    @staticmethod
    def _backward(ctx, *gouts):
        z, sigma, rgb, rays_o, rays_d, pts_o, pts_d, view_dirs, *params = ctx.saved_tensors
        block, opts = ctx.block, ctx.opts
        grads_in = {k: g for k, g in zip(ctx.keys, gouts)}
        vis, vis2, vws, vflags, rays_o2, weights, acc, vis2_map = ctx.vis_state
        tail = gouts[len(ctx.keys) + 2:]                          # gradients of raw visibility / raw visibility2, if any
        d_vis = tail[0] if ctx.n_extra >= 1 else None
        d_vis2 = tail[1] if ctx.n_extra >= 2 else None
        if grads_in.get('visibility2') is not None:             # through the visibility2 map into the weights and acc
            d_map, d_w, d_acc = ops.visibility2_composite_backward(weights, acc, vis2, vis2_map, grads_in['visibility2'])
            d_vis2 = d_map if d_vis2 is None else d_vis2 + d_map
            grads_in['weights'] = d_w if grads_in.get('weights') is None else grads_in['weights'] + d_w
            grads_in['acc'] = d_acc if grads_in.get('acc') is None else grads_in['acc'] + d_acc
        grads_in.pop('visibility2', None)
        d_sigma, d_rgb = ops.composite_backward(sigma, rgb, z, rays_o, rays_d, pts_d if opts['ndc'] else None, opts['ndc'],
                                                opts['white_bkgd'], grads_in)
        table: List[Optional[torch.Tensor]] = [None] * P_COUNT
        gtable: List[Optional[torch.Tensor]] = [None] * P_COUNT
        # one contiguous, zero-initialised gradient bucket per MLP; every tensor starts 16-byte aligned so that the
        # wgrad kernel can flush with vector reductions
        offsets, total = [], 0
        for p in params:
            offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        flat = torch.zeros(total, device=z.device)
        views = [flat[o:o + p.numel()] for o, p in zip(offsets, params)]
        it = iter(zip(params, views))
        out_grads = []
        for i, slot in enumerate(opts['param_mask']):
            if slot:
                p, g = next(it)
                table[i], gtable[i] = p, g.view_as(p)
                out_grads.append(gtable[i])
        bwd_flags = ctx.flags
        if block.predict_visibility:      # first the head's own backward (it pre-fills d hv / d feature in the workspace)
            ops.visibility_backward(block.desc, table, ctx.ws, rays_o, rays_d, z, rays_o2, vis, vis2, d_vis, d_vis2, gtable, vws, vflags)
            bwd_flags |= FLAG_VIS_GRAD
        ops.mlp_backward(block.desc, table, ctx.packed, pts_o, pts_d, view_dirs if block.view_degree else None, z, sigma,
                         rgb, d_sigma, d_rgb, gtable, ctx.ws, bwd_flags)
        ctx.ws = ctx.vis_state = None
        return (None,) * 10 + tuple(out_grads)


class FusedSimpleNeRF(torch.nn.Module):
    def __init__(self, configs: dict, model_configs: Optional[dict] = None):
        super().__init__()
        self.configs = configs
        self.model_configs = model_configs
        mc = configs['model']
        self.ndc = configs['data_loader']['ndc']
        self.coarse_mlp_needed = 'coarse_mlp' in mc
        self.fine_mlp_needed = 'fine_mlp' in mc
        if not self.coarse_mlp_needed:
            raise NotImplementedError('a coarse_mlp is required (the reference cannot run without one either, :202)')
        cfgs = {}
        for attr, path, _, _ in _SLOTS:
            node = mc
            for key in path:
                node = node.get(key) if isinstance(node, dict) else None
                if node is None:
                    break
            if node is not None:
                cfgs[attr] = node
        for attr in _CTOR_ORDER:
            if attr in cfgs:
                setattr(self, attr, MlpBlock(cfgs[attr]))
        self.predict_visibility = any(getattr(self, a).predict_visibility for a in ('coarse_model', 'fine_model') if a in cfgs)   # :19-20
        self.slots = [(a, pre, lvl) for a, _, pre, lvl in _SLOTS if a in cfgs]
        self.precision = mc.get('precision', 'bf16')
        if self.precision not in ('bf16', 'fp32'):
            raise ValueError(f"configs['model']['precision'] must be 'bf16' or 'fp32', got {self.precision!r}")
        self.launch_rays = int(mc.get('launch_rays', 65536))
        self.fused_composite = bool(mc.get('fused_composite', True))   # evaluation: MLP + compositing in one pass (row X1)
        rng = mc.get('rng', 'device')
        if rng not in ('device', 'torch', 'reference'):
            raise ValueError(f"configs['model']['rng'] must be 'device', 'torch' or 'reference', got {rng!r}")
        self.randoms: DeviceRandoms = (ReferenceOrderRandoms(mc.get('netchunk')) if rng == 'reference'
                                       else DeviceRandoms() if rng == 'torch' else KernelRandoms())
        self._const_cache: Dict[tuple, torch.Tensor] = {}

    # torch.linspace is evaluated on the CPU exactly as the reference does (:285, :338) and cached per device
    def _linspace(self, steps: int, device) -> torch.Tensor:
        key = (steps, str(device))
        if key not in self._const_cache:
            self._const_cache[key] = torch.linspace(0., 1., steps=steps).to(device)
        return self._const_cache[key]

    def forward(self, input_batch: dict, retraw: bool = False, sec_views_vis: bool = False):
        retraw = retraw or self.training                                        # :74
        sec_views_vis = sec_views_vis or self.training
        rays_o = input_batch['rays_o']
        if not rays_o.is_cuda:
            raise RuntimeError('FusedSimpleNeRF runs on CUDA only: move the batch to the GPU (no CPU fallback exists)')
        with torch.cuda.device(rays_o.device):      # kernels go to the current device's stream: make it the batch's ('device': [k] configs)
            return self._forward(input_batch, retraw, sec_views_vis)

    def _forward(self, input_batch: dict, retraw: bool, sec_views_vis: bool):
        rays_o = input_batch['rays_o']
        n = rays_o.shape[0]
        if self.predict_visibility and sec_views_vis:                            # :119-133
            input_batch = dict(input_batch)
            if 'rays_o2' not in input_batch:          # the other views' camera centres, own view skipped (index plumbing only)
                poses = input_batch['common_data']['poses']
                poses = poses[0] if poses.dim() == 4 else poses                  # :69-73
                image_id = input_batch['pixel_id'][:, 0].long()
                cols = [poses[i + (i >= image_id).long()][:, :3, 3] for i in range(input_batch['num_frames'] - 1)]
                input_batch['rays_o2'] = torch.stack(cols, dim=1)
            input_batch['rays_o2'] = input_batch['rays_o2'].detach().float().contiguous()
        elif 'rays_o2' in input_batch:
            input_batch = {k: v for k, v in input_batch.items() if k != 'rays_o2'}
        parts: Dict[str, List[torch.Tensor]] = {}
        fixed = isinstance(self.randoms, FixedRandoms)
        for i in range(0, max(n, 1), self.launch_rays):                          # replaces batchify_rays :81-106
            sub = {k: (v[i:i + self.launch_rays] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n else v)
                   for k, v in input_batch.items() if k != 'common_data'}
            if fixed:
                self.randoms.offset = i
                self.randoms._n = sub['rays_o'].shape[0]
            out = self.render_rays(sub, retraw)
            if n <= self.launch_rays:
                return out
            for k, v in out.items():
                parts.setdefault(k, []).append(v)
        return {k: torch.cat(v, 0) for k, v in parts.items()}

    def _stream(self, out: dict, attr: str, prefix: str, level: str, z, batch, retraw: bool):
        block: MlpBlock = getattr(self, attr)
        mc = self.configs['model']
        n, s = z.shape
        dev = z.device
        noise = None
        if self.training and mc['raw_noise_std'] > 0.:                           # :669-671
            noise = self.randoms.sigma_noise(attr, n * s, dev)
            if isinstance(noise, ops.RngDraw):
                noise.scale = float(mc['raw_noise_std'])
            else:
                noise = (noise.reshape(-1).float() * mc['raw_noise_std']).contiguous()
        table = block.param_table()
        params = [p for p in table if p is not None]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        opts = dict(ndc=self.ndc, white_bkgd=bool(mc['white_bkgd']), precise=self.precision == 'fp32', need_grad=need_grad,
                    param_mask=[p is not None for p in table])
        f32 = lambda t: t.detach().float().contiguous()   # noqa: E731
        rays_o, rays_d = f32(batch['rays_o']), f32(batch['rays_d'])
        if self.ndc:
            pts_o, pts_d = f32(batch['rays_o_ndc']), f32(batch['rays_d_ndc'])    # :142
        else:
            pts_o, pts_d = rays_o, rays_d                                        # :140
        view_dirs = f32(batch['view_dirs']) if block.use_view_dirs else rays_d
        rays_o2 = batch.get('rays_o2') if block.predict_visibility else None
        if (self.fused_composite and not retraw and not need_grad and not self.training and self.precision == 'bf16'
                and block.has_view and not block.predict_visibility and s % 32 == 0 and s <= 1024):
            # nothing downstream wants the raw network outputs (:265-269), so the samples stay on chip
            # between the MLP and the compositing sums (snerf_render_forward); `weights` only where the resampler reads them
            detached = [None if q is None else q.detach() for q in table]
            maps = ops.render_forward(block.desc, detached, block.packed(detached), pts_o, pts_d, view_dirs, z, rays_o, rays_d,
                                      self.ndc, bool(mc['white_bkgd']), want_weights=(level == 'coarse' and self.fine_mlp_needed))
            for k, v in maps.items():
                out[f'{prefix}{k}_{level}'] = v
            return maps
        res = _RenderStream.apply(block, opts, z, noise, rays_o, rays_d, pts_o, pts_d, view_dirs, rays_o2, *params)
        keys = ['rgb', 'acc', 'depth', 'depth_var'] + (['depth_ndc', 'depth_var_ndc'] if self.ndc else []) + \
               ['alpha', 'visibility', 'weights'] + (['visibility2'] if rays_o2 is not None else [])
        maps = dict(zip(keys, res[:len(keys)]))
        for k, v in maps.items():
            out[f'{prefix}{k}_{level}'] = v
        if retraw:                                                               # :166-168
            sigma, rgb = res[len(keys)], res[len(keys) + 1]
            if block.predict_visibility:                                         # network outputs :710-713, :649
                out[f'{prefix}raw_visibility_{level}'] = res[len(keys) + 2].unsqueeze(-1)
                if rays_o2 is not None:
                    out[f'{prefix}raw_visibility2_{level}'] = res[len(keys) + 3].unsqueeze(-1)
            out[f'{prefix}raw_sigma_{level}'] = sigma.unsqueeze(-1)
            out[f'{prefix}raw_rgb_{level}'] = rgb
            which = 'rgb_view_dependent' if block.has_view else 'rgb_view_independent'
            out[f'{prefix}raw_{which}_{level}'] = rgb
        return maps

    def render_rays(self, batch: dict, retraw: bool) -> Dict[str, torch.Tensor]:
        mc = self.configs['model']
        dev = batch['rays_o'].device
        n = batch['rays_o'].shape[0]
        near, far = (batch['near_ndc'], batch['far_ndc']) if self.ndc else (batch['near'], batch['far'])
        perturb = bool(mc['perturb']) and self.training                          # :279-281
        out: Dict[str, torch.Tensor] = {}

        s_c = mc['coarse_mlp']['num_samples']
        t_rand = self.randoms.t_rand(n, s_c, dev) if perturb else None
        z_c = ops.sample_coarse(near, far, self._linspace(s_c, dev), t_rand, bool(mc['lindisp']))
        out['z_vals_coarse'] = z_c
        maps_c = None
        for attr, prefix, level in self.slots:
            if level != 'coarse' or (prefix and not self.training):              # aug models: training only (:170, :186)
                continue
            maps = self._stream(out, attr, prefix, level, z_c, batch, retraw)
            if not prefix:
                maps_c = maps
        if self.fine_mlp_needed:
            n_f = mc['fine_mlp']['num_samples']
            u = self.randoms.u(n, n_f, dev) if perturb else self._linspace(n_f, dev)
            z_f = ops.sample_fine(z_c, maps_c['weights'], u)                     # :202 (no gradient, :312)
            out['z_vals_fine'] = z_f
            for attr, prefix, level in self.slots:
                if level != 'fine' or (prefix and not self.training):
                    continue
                self._stream(out, attr, prefix, level, z_f, batch, retraw)
        if not retraw:                                                           # :265-269
            for level in ('coarse', 'fine'):
                for k in ('z_vals', 'visibility', 'weights'):
                    out.pop(f'{k}_{level}', None)
        return out
