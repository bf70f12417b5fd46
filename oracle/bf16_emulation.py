"""Torch restatement of the ROUNDING POINTS of the bf16 tensor path (test infrastructure only).

Same algorithm as ``oracle.nerf_oracle.mlp_forward`` (reference src/models/SimpleNeRF01.py:626-715) with
bf16 rounding applied exactly where the tcgen05 kernels round: the encoded points and every hidden activation
that feed a tensor-core GEMM, and the weights of those GEMMs.  The rgb head and the view-direction part of the view layer stay
fp32, as in the kernels; the sigma head of an MLP with a view branch is one more row of the view step's tensor-core product
(bf16 row, bf16 activation), that of an MLP without one is an fp32 dot product on the un-rounded activation.  feature_linear has no activation (:691-697), and the kernels multiply the
last trunk activation by the merged matrix  W_view[:, :256] @ W_feat  (formed in fp32, rounded to bf16 once; the feature
bias reaches the view layer in fp32): the feature vector itself is never formed, so it is not rounded here either.  Rounding is straight-through in the
backward pass, so autograd yields the gradient the kernels are expected to produce (it
isolates kernel bugs from the unavoidable ReLU-mask flips of a bf16 forward)."""
import torch
import torch.nn.functional as F

from . import nerf_oracle as orc


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


bf = _RoundBf16.apply


def mlp_forward_bf16(spec, params, pts, view_dirs, sigma_noise=None, view_dirs2=None):
    enc = orc.positional_encoding(pts, spec.pts_degree)
    e_bf = bf(enc)
    x = e_bf[:, :spec.trunk_in]
    h32 = None
    for i in range(spec.depth):
        acc = F.linear(x, bf(params[f'pts_linears.{i}.weight']))
        if i < spec.depth - 1 or spec.view_dep_rgb:
            # hidden layers (and the last trunk layer of an MLP with a view branch): packed epilogue,
            # HFMA2.BF16.RELU(bf16(acc), 1, bf16(bias)) = one rounding of the sum
            h32 = F.relu(bf(bf(acc) + bf(params[f'pts_linears.{i}.bias'])))
        else:
            # last trunk layer without a view branch: fp32 epilogue (its un-rounded output feeds the fp32 sigma + rgb head)
            h32 = F.relu(acc + params[f'pts_linears.{i}.bias'])
        x = bf(h32)
        if i in spec.skips:
            x = torch.cat([e_bf[:, :spec.trunk_in], x], -1)
    if spec.view_dep_rgb:
        # the sigma row rides the view step of the tensor core: bf16 row times the bf16 activation, fp32 accumulation and bias
        head = F.linear(x, bf(params['pts_output_linear.weight'])) + params['pts_output_linear.bias']
    else:
        head = F.linear(h32, params['pts_output_linear.weight'], params['pts_output_linear.bias'])
    sigma = head[..., 0:1]
    if sigma_noise is not None:
        sigma = sigma + sigma_noise
    out = {'sigma': F.relu(sigma)}
    if not spec.view_dep_rgb:
        out['rgb'] = out['rgb_view_independent'] = torch.sigmoid(head[..., 1:4])
        return out
    wv = params['views_linears.0.weight']
    n_hi = spec.pts_enc_dim - spec.trunk_in
    venc = orc.positional_encoding(view_dirs, spec.view_degree)
    merged = bf(wv[:, :spec.width] @ params['feature_linear.weight'])
    pre_pt = F.linear(x, merged)                 # the part every view shares: the tensor core's accumulator
    if n_hi:
        pre_pt = pre_pt + F.linear(e_bf[:, spec.trunk_in:], bf(wv[:, spec.width:spec.width + n_hi]))
    b_common = F.linear(params['feature_linear.bias'], wv[:, :spec.width]) + params['views_linears.0.bias']
    w_dir = wv[:, spec.width + n_hi:]
    hv = F.relu(pre_pt + b_common + F.linear(venc, w_dir))
    vout = F.linear(hv, params['views_output_linear.weight'], params['views_output_linear.bias'])
    out['rgb'] = out['rgb_view_dependent'] = torch.sigmoid(vout[..., 0:3])
    if spec.predict_visibility:
        # visibility head (vis_tc.cu): the shared accumulator reaches the CUDA-core kernel rounded to bf16, the per-view
        # direction part, the ReLU and the fourth row are fp32 (own view and other views alike)
        w_vis, b_vis = params['views_output_linear.weight'][3:4], params['views_output_linear.bias'][3:4]
        shared = bf(pre_pt) + b_common
        out['visibility'] = torch.sigmoid(F.linear(F.relu(shared + F.linear(venc, w_dir)), w_vis, b_vis))
        if view_dirs2 is not None:
            venc2 = orc.positional_encoding(view_dirs2, spec.view_degree)                     # [P, nf-1, 27]
            hv2 = F.relu(shared[:, None, :] + F.linear(venc2, w_dir))
            out['visibility2'] = torch.sigmoid(F.linear(hv2, w_vis, b_vis))
    return out
