"""Host mirror of the reference's loss front end for the masked per-ray losses.

Fused: MSE01-03, SparseDepthMSE01-03, DenseDepthMSE01, PointsAugmentationDepthLoss01/02, ViewsAugmentationDepthLoss01/02,
CoarseFineConsistencyLoss01/02, VisibilityLoss01, VisibilityPriorLoss01 -- all 15 loss modules of the reference.  `FusedLossComputer(configs).compute_losses(input_dict, output_dict)` has the contract of
`loss_functions.LossComputer01.LossComputer.compute_losses` (src/loss_functions/LossComputer01.py:33-52): it returns
`{loss_name: {'loss_value': tensor}, ..., 'TotalLoss': tensor}` and `TotalLoss.backward()` fills the gradients of the
model outputs.  The six losses that are plain masked means -- MSE01/02/03 (MSE01.py:26-67) and SparseDepthMSE01/02/03
(SparseDepthMSE01.py:26-71) -- run as ONE forward launch and ONE backward launch of `snerf_ray_losses_*` instead of
~20 eager kernels and a boolean-mask gather (a device synchronisation) per stream; the three patch-reprojection depth
losses PointsAugmentationDepthLoss02 / ViewsAugmentationDepthLoss02 / CoarseFineConsistencyLoss02 share one forward and
one backward launch of `snerf_reprojection_losses_*` (one warp per ray; the source patch and the main model's patch are
gathered once for all three).  Every other configured loss
is taken from `extra_losses` (name -> object with the reference's `compute_loss` signature, e.g. the reference's own
instance) and added with torch; a configured loss that is neither fused nor supplied raises unless its weight at this
iteration is 0.  There is no CPU path: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib, ops

# loss name -> (kind, config sub-key of configs['model'] holding its MLPs or None for the main model, output prefix)
_RGB = {'MSE01': (None, ''), 'MSE02': ('points_augmentation', 'points_augmentation_'),
        'MSE03': ('views_augmentation', 'views_augmentation_')}
_DEPTH = {'SparseDepthMSE01': (None, ''), 'SparseDepthMSE02': ('points_augmentation', 'points_augmentation_'),
          'SparseDepthMSE03': ('views_augmentation', 'views_augmentation_')}
# patch-reprojection depth losses: name -> (config sub-key of the other model or None for the fine model, other depth key)
_REPROJ = {'PointsAugmentationDepthLoss02': ('points_augmentation', 'points_augmentation_depth_coarse'),
           'ViewsAugmentationDepthLoss02': ('views_augmentation', 'views_augmentation_depth_coarse'),
           'CoarseFineConsistencyLoss02': (None, 'depth_fine')}
# plain (unmasked) MSE between two model depths, gradients to BOTH sides: name -> (config sub-key or None, other depth key per level)
_PAIR = {'PointsAugmentationDepthLoss01': ('points_augmentation', 'points_augmentation_depth_{level}'),
         'ViewsAugmentationDepthLoss01': ('views_augmentation', 'views_augmentation_depth_{level}'),
         'CoarseFineConsistencyLoss01': (None, None)}
FUSED_LOSSES = tuple(_RGB) + tuple(_DEPTH)


def get_loss_weight(loss_configs: dict, iter_num: int) -> float:
    """LossComputer.get_loss_weight (LossComputer01.py:54-69)."""
    if 'weight' in loss_configs:
        return loss_configs['weight']
    if 'iter_weights' in loss_configs:
        for key in sorted((int(k) for k in loss_configs['iter_weights']), reverse=True):
            if iter_num >= key:
                return loss_configs['iter_weights'][str(key)]
    raise RuntimeError(f"loss_weight is None for {loss_configs.get('name')} at iter {iter_num}")


def stream_plan(configs: dict, loss_name: str, input_dict: dict, output_dict: dict) -> List[Tuple[str, str, str]]:
    """The (prediction key, target key, mask key) triples one reference loss module reads, in its order."""
    model = configs['model']
    if loss_name in _RGB:
        sub, prefix = _RGB[loss_name]
        mlps = model if sub is None else model[sub]
        plan = []
        for level in ('coarse', 'fine'):                                                     # MSE01.py:32-45, MSE02.py:32-45
            key = f'{prefix}rgb_{level}'
            if f'{level}_mlp' in mlps and (sub is None or key in output_dict):
                plan.append((key, 'target_rgb', 'indices_mask_nerf'))
        return plan
    sub, prefix = _DEPTH[loss_name]
    if 'indices_mask_sparse_depth' not in input_dict:                                        # SparseDepthMSE01.py:31-32
        return []
    mlps = model if sub is None else model[sub]
    # SparseDepthMSE02.py:37-45 reads depth_fine (not the augmented fine depth) when an augmented fine MLP exists
    key = 'depth_fine' if 'fine_mlp' in mlps else f'{prefix}depth_coarse'
    return [(key, 'sparse_depth_values', 'indices_mask_sparse_depth')]


class _RayLosses(torch.autograd.Function):
    """values[n+1] = per-stream masked means, then their weighted sum; preds are differentiable."""

    @staticmethod
    def forward(ctx, targets, masks, weights, kinds, workspace, *preds):
        n_streams, n_rays = len(preds), preds[0].shape[0]
        dev = preds[0].device
        table = (_lib.LossStream * n_streams)()
        keep = []
        for s, (p, t, m, w) in enumerate(zip(preds, targets, masks, weights)):
            p32, t32 = ops._f32(p).reshape(n_rays, -1), ops._f32(t).reshape(n_rays, -1)
            if p32.shape != t32.shape:
                raise RuntimeError(f'stream {s}: prediction {tuple(p.shape)} vs target {tuple(t.shape)}')
            m8 = None
            if m is not None:
                m8 = m.detach().contiguous().view(torch.uint8) if m.dtype == torch.bool else m.detach().to(torch.uint8).contiguous()
                if m8.shape[0] != n_rays:
                    raise RuntimeError(f'stream {s}: mask of {m8.shape[0]} rays for {n_rays} rays')
            keep.append((p32, t32, m8))
            table[s].pred, table[s].target = ops._ptr(p32), ops._ptr(t32)
            table[s].mask = ops._ptr(m8, torch.uint8)
            table[s].grad = None
            table[s].channels, table[s].weight, table[s].kind = p32.shape[1], float(w), int(kinds[s])
        values = torch.empty(n_streams + 1, device=dev, dtype=torch.float32)
        counts = torch.empty(n_streams, device=dev, dtype=torch.int32)
        ops.LAUNCHES['count'] += 1
        _lib.check(_lib.load().snerf_ray_losses_forward(table, n_streams, n_rays, ops._ptr(values), ops._ptr(counts, torch.int32),
                                                        ops._ptr(workspace, torch.uint8), workspace.numel(), ops._stream()),
                   'snerf_ray_losses_forward')
        ctx.keep, ctx.counts, ctx.weights, ctx.shapes = keep, counts, [float(w) for w in weights], [p.shape for p in preds]
        ctx.kinds = [int(k) for k in kinds]
        return values

    @staticmethod
    def backward(ctx, g_values):
        n_streams, n_rays = len(ctx.keep), ctx.keep[0][0].shape[0]
        table = (_lib.LossStream * n_streams)()
        grads = []
        for s, (p32, t32, m8) in enumerate(ctx.keep):
            g = torch.empty_like(p32)
            grads.append(g)
            table[s].pred, table[s].target, table[s].mask = ops._ptr(p32), ops._ptr(t32), ops._ptr(m8, torch.uint8)
            table[s].grad, table[s].channels, table[s].weight, table[s].kind = ops._ptr(g), p32.shape[1], ctx.weights[s], ctx.kinds[s]
        ops.LAUNCHES['count'] += 1
        _lib.check(_lib.load().snerf_ray_losses_backward(table, n_streams, n_rays, ops._ptr(ctx.counts, torch.int32),
                                                         ops._ptr(ops._f32(g_values)), ops._stream()), 'snerf_ray_losses_backward')
        return (None,) * 5 + tuple(g.reshape(shape) for g, shape in zip(grads, ctx.shapes))


class _ReprojectionLosses(torch.autograd.Function):
    """values[k+1] = the reprojection loss of (main, other_k) for every k, then their weighted sum."""

    @staticmethod
    def forward(ctx, views, rays, weights, half_patch, threshold, symmetric, workspace, depth_main, *depth_others):
        n_rays, k = depth_main.shape[0], len(depth_others)
        dev = depth_main.device
        main32 = ops._f32(depth_main).reshape(-1)
        others32 = [ops._f32(d).reshape(-1) for d in depth_others]
        args = _lib.ReprojArgs()
        args.depth_main, args.n_others = ops._ptr(main32), k
        for i, d in enumerate(others32):
            if d.shape[0] != n_rays:
                raise RuntimeError(f'other depth {i}: {d.shape[0]} rays for {n_rays}')
            args.depth_other[i], args.weight[i] = ops._ptr(d), float(weights[i])
        rays_o, rays_d, pixel_id, mask = rays
        rays_o, rays_d = ops._f32(rays_o), ops._f32(rays_d)
        pixel_id = pixel_id.detach().to(torch.int32).contiguous()
        m8 = None if mask is None else (mask.detach().contiguous().view(torch.uint8) if mask.dtype == torch.bool
                                        else mask.detach().to(torch.uint8).contiguous())
        images, proj, origins, closest = views
        args.rays_o, args.rays_d, args.pixel_id = ops._ptr(rays_o), ops._ptr(rays_d), ops._ptr(pixel_id, torch.int32)
        args.mask_nerf, args.images = ops._ptr(m8, torch.uint8), ops._ptr(images)
        args.proj, args.origins, args.closest = ops._ptr(proj), ops._ptr(origins), ops._ptr(closest, torch.int32)
        args.n_views, args.height, args.width = images.shape[0], images.shape[1], images.shape[2]
        args.half_patch, args.rmse_threshold = int(half_patch), float(threshold)
        args.flags = _lib.REPROJ_SYMMETRIC if symmetric else 0
        codes = torch.empty((k, n_rays), device=dev, dtype=torch.uint8)
        values = torch.empty(k + 1, device=dev, dtype=torch.float32)
        counts = torch.empty(1, device=dev, dtype=torch.int32)
        ops.LAUNCHES['count'] += 1
        _lib.check(_lib.load().snerf_reprojection_losses_forward(C.byref(args), n_rays, ops._ptr(codes, torch.uint8), ops._ptr(values),
                                                                 ops._ptr(counts, torch.int32), ops._ptr(workspace, torch.uint8),
                                                                 workspace.numel(), ops._stream()),
                   'snerf_reprojection_losses_forward')
        ctx.keep = (main32, others32, codes, counts, [float(w) for w in weights], bool(symmetric))
        ctx.shapes = [depth_main.shape] + [d.shape for d in depth_others]
        ctx.mark_non_differentiable(codes)
        return values, codes

    @staticmethod
    def backward(ctx, g_values, _g_codes):
        main32, others32, codes, counts, weights, symmetric = ctx.keep
        n_rays, k = main32.shape[0], len(others32)
        args = _lib.ReprojArgs()
        g_main = torch.empty_like(main32)
        g_others = [torch.empty_like(d) for d in others32]
        args.depth_main, args.grad_main, args.n_others = ops._ptr(main32), ops._ptr(g_main), k
        for i, (d, g) in enumerate(zip(others32, g_others)):
            args.depth_other[i], args.grad_other[i], args.weight[i] = ops._ptr(d), ops._ptr(g), weights[i]
        args.flags = _lib.REPROJ_SYMMETRIC if symmetric else 0
        ops.LAUNCHES['count'] += 1
        _lib.check(_lib.load().snerf_reprojection_losses_backward(C.byref(args), n_rays, ops._ptr(codes, torch.uint8),
                                                                  ops._ptr(counts, torch.int32), ops._ptr(ops._f32(g_values)),
                                                                  ops._stream()), 'snerf_reprojection_losses_backward')
        grads = [g.reshape(shape) for g, shape in zip([g_main] + g_others, ctx.shapes)]
        return (None,) * 7 + tuple(grads)


def view_tables(poses: torch.Tensor, intrinsics: torch.Tensor):
    """Per-view tables of the reprojection kernel, with the reference's own torch expressions:
    proj[v] = intrinsics[:1] @ permuter @ R_v^T (CommonUtils01.py:60-69), origins, and the nearest other view
    (second smallest camera distance, PointsAugmentationDepthLoss02.py:126-130)."""
    permuter = torch.eye(3, device=poses.device)
    permuter[1:] *= -1
    proj = (intrinsics[:1].float() @ permuter[None] @ poses[:, :3, :3].float().transpose(1, 2)).reshape(-1, 9).contiguous()
    origins = poses[:, :3, 3].float().contiguous()
    dist = torch.sqrt(torch.sum(torch.square(origins[:, None, :] - origins[None, :, :]), dim=2))
    closest = torch.kthvalue(dist, 2, dim=1)[1].to(torch.int32).contiguous()
    return proj, origins, closest


def reprojection_losses(depth_main, depth_others, weights, rays_o, rays_d, pixel_id, mask_nerf, images, poses, intrinsics,
                        patch_size=(5, 5), rmse_threshold=0.1, symmetric=False, tables=None):
    """values[len(depth_others) + 1] (loss per pair, then the weighted total) and codes[len(depth_others), N]
    (bit 0: the main model is the more accurate one on the ray, bit 1: the other model is)."""
    if not depth_main.is_cuda:
        raise RuntimeError('simplenerf_b200 kernels need CUDA tensors (no CPU fallback exists)')
    if not 1 <= len(depth_others) <= _lib.REPROJ_MAX_OTHERS:
        raise RuntimeError(f'{len(depth_others)} other depths (1..{_lib.REPROJ_MAX_OTHERS} per call)')
    px, py = patch_size
    if px != py:
        raise NotImplementedError('square patches only (the reference pads W by hpy and H by hpx, :156)')
    proj, origins, closest = tables if tables is not None else view_tables(poses, intrinsics)
    views = (ops._f32(images), proj, origins, closest)
    return _ReprojectionLosses.apply(views, (rays_o, rays_d, pixel_id, mask_nerf), list(weights), px // 2, rmse_threshold,
                                     symmetric, _workspace(depth_main.device), depth_main, *depth_others)


_WORKSPACES: Dict[tuple, torch.Tensor] = {}


def _workspace(dev: torch.device) -> torch.Tensor:
    """Per (device, stream): the forward kernels keep block partials and a ticket counter in it, so two streams must not
    share one.  Zeroed once; the kernel leaves its ticket counter at zero."""
    key = (dev, torch.cuda.current_stream(dev).cuda_stream)
    if key not in _WORKSPACES:
        _WORKSPACES[key] = torch.zeros(_lib.load().snerf_ray_losses_workspace_bytes(), device=dev, dtype=torch.uint8)
    return _WORKSPACES[key]


def ray_losses(preds: Sequence[torch.Tensor], targets: Sequence[torch.Tensor], masks: Sequence[Optional[torch.Tensor]],
               weights: Sequence[float], kinds: Optional[Sequence[int]] = None) -> torch.Tensor:
    """values[len(preds) + 1]: masked mean of every stream, then sum_s weights[s] * values[s].  kinds[s]: _lib.LOSS_SQUARED
    (default: squared error), LOSS_ABSOLUTE (absolute error) or LOSS_PRIOR_SHORTFALL (sum_c target_c (1 - pred_c) per ray)."""
    if not 1 <= len(preds) <= _lib.LOSS_MAX_STREAMS:
        raise RuntimeError(f'{len(preds)} loss streams (1..{_lib.LOSS_MAX_STREAMS} per call)')
    if not preds[0].is_cuda:
        raise RuntimeError('simplenerf_b200 kernels need CUDA tensors (no CPU fallback exists)')
    kinds = [_lib.LOSS_SQUARED] * len(preds) if kinds is None else list(kinds)
    return _RayLosses.apply(list(targets), list(masks), list(weights), kinds, _workspace(preds[0].device), *preds)


def ray_loss_maps(preds: Sequence[torch.Tensor], targets: Sequence[torch.Tensor], masks: Sequence[Optional[torch.Tensor]],
                  kinds: Optional[Sequence[int]] = None) -> List[torch.Tensor]:
    """Per-ray error of every stream, averaged over its channels, compacted to the masked-in rays like the reference's
    `loss_maps` entries (MSE01.py:53-66 indexes with the mask first).  No gradient (a validation output)."""
    n_streams, n_rays = len(preds), preds[0].shape[0]
    kinds = [_lib.LOSS_SQUARED] * n_streams if kinds is None else list(kinds)
    table = (_lib.LossStream * n_streams)()
    keep, maps = [], []
    for s, (p, t, m) in enumerate(zip(preds, targets, masks)):
        p32, t32 = ops._f32(p).reshape(n_rays, -1), ops._f32(t).reshape(n_rays, -1)
        m8 = None if m is None else (m.detach().contiguous().view(torch.uint8) if m.dtype == torch.bool else m.detach().to(torch.uint8).contiguous())
        out = torch.empty(n_rays, device=p32.device, dtype=torch.float32)
        keep.append((p32, t32, m8))
        maps.append(out)
        table[s].pred, table[s].target, table[s].mask, table[s].grad = ops._ptr(p32), ops._ptr(t32), ops._ptr(m8, torch.uint8), ops._ptr(out)
        table[s].channels, table[s].weight, table[s].kind = p32.shape[1], 1.0, int(kinds[s])
    ops.LAUNCHES['count'] += 1
    _lib.check(_lib.load().snerf_ray_loss_maps(table, n_streams, n_rays, ops._stream()), 'snerf_ray_loss_maps')
    return [mp if m is None else mp[m.bool()] for mp, m in zip(maps, masks)]


class FusedLossComputer:
    def __init__(self, configs: dict, extra_losses: Optional[dict] = None, symmetric_reprojection: bool = False,
                 ray_sharded: bool = False, group=None):
        """ray_sharded: this process sees one shard of the step's rays (one process per GPU); the masked means are then
        weighted by  n_rank * world / n_global  per mask (distributed.mask_count_weights) so that the average of the ranks'
        gradients is the gradient of the reference's global masked means."""
        self.ray_sharded, self.group = ray_sharded, group
        self.configs = configs
        self.loss_configs = {lc['name']: lc for lc in configs['losses']}
        self.extra_losses = dict(extra_losses or {})
        self.symmetric_reprojection = symmetric_reprojection     # False = what the reference computes (see losses.cu)
        self._tables = None                                      # (key, per-view tables) of the last poses seen

    def _view_tables(self, poses, intrinsics):
        """Per-view tables, recomputed only when the poses / intrinsics change.  The key is the tensors' storage, offset, shape
        and version counter; the cache holds the tensors themselves, so their storage cannot be freed and handed to
        different data under the same address while the entry lives."""
        key = tuple((t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), t._version) for t in (poses, intrinsics))
        if self._tables is None or self._tables[0] != key:
            self._tables = (key, view_tables(poses, intrinsics), (poses, intrinsics))
        return self._tables[1]

    @staticmethod
    def _align_mirror(owner, mirror):
        out, j = [None] * len(owner), 0
        for i, tag in enumerate(owner):
            if tag is None:
                while mirror[j] is None:
                    j += 1
                out[i] = mirror[j]
                j += 1
        return out

    def _reprojection_plan(self, name: str):
        """(other depth key) if the reference module would compute a coarse pair, else None (:44-51, CoarseFine :34-37)."""
        model = self.configs['model']
        sub, other_key = _REPROJ[name]
        if sub is None:
            return other_key if ('coarse_mlp' in model and 'fine_mlp' in model) else None
        if 'fine_mlp' in model and 'fine_mlp' in model[sub]:
            raise NotImplementedError(f'{name}: fine-level augmentation pairs are not built (never enabled in a shipped config)')
        return other_key if ('coarse_mlp' in model and 'coarse_mlp' in model[sub]) else None

    def compute_losses(self, input_dict: dict, output_dict: dict, return_loss_maps: bool = False) -> dict:
        """return_loss_maps (validation, src/Trainer01.py:195-196): every loss whose terms are per-ray streams also returns
        `loss_maps` = {'<Module>_<level>': per-ray error of the masked-in rays}, the reference's names (LossUtils01.py:7-10).
        The patch-reprojection losses (`*02`) report values only: their maps are not built."""
        iter_num = input_dict['iter_num']
        preds, targets, masks, weights, owner = [], [], [], [], []
        kinds: Dict[int, int] = {}     # stream index -> kind, for the streams that are not squared errors
        map_names: Dict[int, str] = {}  # stream index -> loss-map name (return_loss_maps)
        mirror: list = []       # per stream: None, or the weight of a gradient-only mirror stream (two-sided losses)
        reproj: Dict[tuple, list] = {}
        extra_total = 0
        loss_values: Dict[str, dict] = {}
        for name, lc in self.loss_configs.items():
            weight = get_loss_weight(lc, iter_num)
            if name in FUSED_LOSSES:
                plan = stream_plan(self.configs, name, input_dict, output_dict)
                for pred_key, target_key, mask_key in plan:
                    target = input_dict[target_key]
                    if name in _RGB:     # the sparse-depth modules return an empty `loss_maps` (SparseDepthMSE01.py:67-70)
                        map_names[len(preds)] = f"{name}_{'fine' if pred_key.endswith('_fine') else 'coarse'}"
                    preds.append(output_dict[pred_key])
                    targets.append(target[:, 0] if target_key == 'sparse_depth_values' else target)   # SparseDepthMSE01.py:34
                    masks.append(input_dict[mask_key])
                    weights.append(weight)
                    owner.append(name)
                if not plan:
                    loss_values[name] = {'loss_value': torch.zeros((), device=input_dict['rays_o'].device)}
            elif name == 'VisibilityLoss01' and name not in self.extra_losses:
                # two-sided MAE between the predicted visibility and the transmittance, each side detached in turn
                # (VisibilityLoss01.py:26-74): stream (pred | T) moves the head, stream (T | pred) moves sigma; no ray mask
                for level in ('coarse', 'fine'):
                    if f'{level}_mlp' in self.configs['model']:
                        a, b = output_dict[f'raw_visibility_{level}'][..., 0], output_dict[f'visibility_{level}']
                        for p_t, t_t in ((a, b), (b, a)):
                            kinds[len(preds)] = _lib.LOSS_ABSOLUTE
                            preds.append(p_t)
                            targets.append(t_t.detach())
                            masks.append(None)
                            weights.append(weight)
                            owner.append(name)
            elif name == 'VisibilityPriorLoss01' and name not in self.extra_losses:
                # mean over the NeRF rays of sum_v prior_v (1 - visibility2_v)  (VisibilityPriorLoss01.py:25-80)
                model = self.configs['model']
                levels = [lv for lv in ('coarse', 'fine') if f'{lv}_mlp' in model]
                if any(f'raw_visibility2_{lv}' not in output_dict for lv in levels):         # :29-31: the module returns None
                    continue
                prior = input_dict.get('visibility_prior_masks', input_dict.get('visibility_prior_weights'))
                for lv in levels:
                    pred = output_dict[f'visibility2_{lv}']
                    kinds[len(preds)] = _lib.LOSS_PRIOR_SHORTFALL
                    preds.append(pred)
                    targets.append(torch.ones_like(pred).detach() if prior is None else prior.to(pred.dtype))
                    masks.append(input_dict['indices_mask_nerf'])
                    weights.append(weight)
                    owner.append(name)
            elif name == 'DenseDepthMSE01' and name not in self.extra_losses:
                # masked MSE against the dense depth prior (DenseDepthMSE01.py:26-68).  The reference slices the fine depth with
                # an attribute it never sets (`self.num_rays`, :41) and cannot run with a fine MLP; here the whole batch is used.
                for level in ('coarse', 'fine'):
                    if f'{level}_mlp' in self.configs['model']:
                        map_names[len(preds)] = f'{name}_{level}'
                        preds.append(output_dict[f'depth_{level}'])
                        targets.append(input_dict['dense_depth_values'][:, 0])
                        masks.append(input_dict['indices_mask_nerf'])
                        weights.append(weight)
                        owner.append(name)
            elif name in _PAIR and name not in self.extra_losses:
                # mean((a - b)^2) over every ray with gradients to both depths (PointsAugmentationDepthLoss01.py:27-74,
                # CoarseFineConsistencyLoss01.py:25-49): stream (a | b) carries the value and a's gradient, stream (b | a)
                # carries b's gradient only (its value enters the total as v - v.detach()).
                model = self.configs['model']
                sub, other = _PAIR[name]
                if sub is None:
                    pairs = [('depth_coarse', 'depth_fine')] if ('coarse_mlp' in model and 'fine_mlp' in model) else []
                else:
                    pairs = [(f'depth_{lv}', other.format(level=lv)) for lv in ('coarse', 'fine')
                             if f'{lv}_mlp' in model and f'{lv}_mlp' in model[sub]]
                for a_key, b_key in pairs:
                    for p_key, t_key, tag in ((a_key, b_key, name), (b_key, a_key, None)):
                        preds.append(output_dict[p_key])
                        targets.append(output_dict[t_key].detach())
                        masks.append(None)
                        weights.append(weight if tag else 0.0)
                        owner.append(tag)
                        mirror.append(None if tag else weight)
                if not pairs:
                    loss_values[name] = {'loss_value': torch.zeros((), device=input_dict['rays_o'].device)}
            elif name in _REPROJ and name not in self.extra_losses:
                other_key = self._reprojection_plan(name)
                if other_key is None:
                    loss_values[name] = {'loss_value': torch.zeros((), device=input_dict['rays_o'].device)}
                    continue
                setting = (tuple(lc['patch_size']), float(lc['rmse_threshold']))
                reproj.setdefault(setting, []).append((name, other_key, weight))
                if name == 'CoarseFineConsistencyLoss02' and 'sparse_depth' in self.configs['data_loader'] \
                        and input_dict.get('indices_mask_sparse_depth') is not None:
                    # compute_loss_sd (CoarseFineConsistencyLoss02.py:174-189): the fine depth supervises the coarse depth
                    preds.append(output_dict['depth_coarse'])
                    targets.append(output_dict['depth_fine'].detach())
                    masks.append(input_dict['indices_mask_sparse_depth'])
                    weights.append(weight)
                    owner.append(name)
            elif name in self.extra_losses:
                loss_dict = self.extra_losses[name].compute_loss(input_dict, output_dict, return_loss_maps=False)
                if loss_dict is not None:                                                     # LossComputer01.py:46
                    loss_values[name] = loss_dict
                    extra_total = extra_total + weight * loss_dict['loss_value']
            elif weight != 0:
                raise RuntimeError(f'Unknown Loss Function: {name} (not fused; pass an object for it in extra_losses)')
        mirror_full = self._align_mirror(owner, mirror)      # stream index -> weight of a gradient-only mirror stream
        total = extra_total
        scales = None
        if self.ray_sharded:
            from ..distributed import mask_count_weights
            named = {k: input_dict[k] for k in ('indices_mask_nerf', 'indices_mask_sparse_depth') if input_dict.get(k) is not None}
            from ..distributed import ALL_RAYS
            scales = mask_count_weights(named, self.group, n_rays=input_dict['rays_o'].shape[0])
            by_ptr = {m.data_ptr(): scales[k] for k, m in named.items()}
            all_rays = scales[ALL_RAYS]          # mask-less means (pair losses, VisibilityLoss01): weight of the shard's ray count
        for i in range(0, len(preds), _lib.LOSS_MAX_STREAMS):
            sl = slice(i, i + _lib.LOSS_MAX_STREAMS)
            values = ray_losses(preds[sl], targets[sl], masks[sl], weights[sl],
                                [kinds.get(i + j, _lib.LOSS_SQUARED) for j in range(len(preds[sl]))])
            for j, name in enumerate(owner[sl]):        # a module with a coarse and a fine stream reports their sum (MSE01.py:35,42)
                if name is None:                        # gradient-only mirror of a two-sided loss
                    mw = mirror_full[i + j] if scales is None else mirror_full[i + j] * all_rays
                    total = total + mw * (values[j] - values[j].detach())
                    continue
                prev = loss_values.get(name, {}).get('loss_value')
                loss_values[name] = {'loss_value': values[j] if prev is None else prev + values[j]}
            if scales is None:
                total = total + values[-1]
            else:       # per-stream weight * count scale, applied on the device
                w = torch.stack([(all_rays if m is None else by_ptr[m.data_ptr()]) * wt for m, wt in zip(masks[sl], weights[sl])])
                total = total + (values[:-1] * w).sum()
        if reproj:
            cd = input_dict['common_data']
            poses, images, intrinsics = cd['poses'], cd['images'], cd['intrinsics']
            if poses.dim() == 4:      # still carrying the replica dimension that LossComputer01.py:34-38 strips
                poses, images, intrinsics = poses[0], images[0], intrinsics[0]
            tables = self._view_tables(poses, intrinsics)
            for (patch, threshold), items in reproj.items():
                values, _ = reprojection_losses(output_dict['depth_coarse'], [output_dict[k] for _, k, _ in items],
                                                [w for _, _, w in items], input_dict['rays_o'], input_dict['rays_d'],
                                                input_dict['pixel_id'], input_dict['indices_mask_nerf'], images, poses, intrinsics,
                                                patch, threshold, self.symmetric_reprojection, tables)
                for j, (name, _, _) in enumerate(items):
                    prev = loss_values.get(name, {}).get('loss_value')
                    loss_values[name] = {'loss_value': values[j] if prev is None else prev + values[j]}
                if scales is None:
                    total = total + values[-1]
                else:       # the reprojection losses are means over the rays of the NeRF mask
                    w = torch.tensor([wt for _, _, wt in items], device=values.device) * scales['indices_mask_nerf']
                    total = total + (values[:-1] * w).sum()
        if return_loss_maps and map_names:
            idx = sorted(map_names)
            for i in range(0, len(idx), _lib.LOSS_MAX_STREAMS):
                part = idx[i:i + _lib.LOSS_MAX_STREAMS]
                with torch.no_grad():
                    maps = ray_loss_maps([preds[j] for j in part], [targets[j] for j in part], [masks[j] for j in part],
                                         [kinds.get(j, _lib.LOSS_SQUARED) for j in part])
                for j, mp in zip(part, maps):
                    loss_values.setdefault(owner[j], {'loss_value': torch.zeros((), device=mp.device)}).setdefault('loss_maps', {})[map_names[j]] = mp
        if return_loss_maps:
            for v in loss_values.values():
                if isinstance(v, dict):
                    v.setdefault('loss_maps', {})
        loss_values['TotalLoss'] = total
        return loss_values
