"""N>1 host logic on CPU: two gloo ranks, ray-sharded, gradient all-reduce == single-process gradient.
The compute inside each rank is the CPU oracle (the CUDA path needs a GPU); what is under test is the sharding and the
gradient exchange of simplenerf_b200/distributed.py, which bench.py uses unchanged over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nerf_oracle as orc
from simplenerf_b200 import synthetic
from simplenerf_b200.distributed import GradientExchange, allreduce_gradients, shard_bounds, shard_rays

N_RAYS = 24


def _setup():
    configs = synthetic.make_configs('vanilla')
    configs['model']['coarse_mlp']['num_samples'] = 8
    configs['model']['fine_mlp']['num_samples'] = 8
    model = orc.NerfOracle(configs)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synthetic.densify_state(synthetic.deterministic_state(shapes, 4)))
    model.eval()      # deterministic sampling; gradients still flow
    batch = synthetic.make_ray_batch('llff', N_RAYS, 9)
    target = torch.rand((N_RAYS, 3), generator=torch.Generator().manual_seed(1))
    return model, batch, target


def _loss(model, batch, target):
    out = model(batch, retraw=True)
    return ((out['rgb_fine'] - target) ** 2).mean() + ((out['rgb_coarse'] - target) ** 2).mean()


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    model, batch, target = _setup()
    lo, hi = shard_bounds(N_RAYS, rank, world)
    _loss(model, shard_rays(batch, rank, world), target[lo:hi]).backward()
    allreduce_gradients(model.parameters(), weight=(hi - lo) / N_RAYS)
    if rank == 0:
        ret['grads'] = {k: p.grad.clone() for k, p in model.named_parameters()}
    dist.destroy_process_group()


def _bucket_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # gradients laid out the way the drop-in's backward hands them out: views of one zero-padded flat bucket per MLP
    shapes = [(5, 3), (7,), (2, 2), (1,)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    total = sum((p.numel() + 3) // 4 * 4 for p in params)
    flat = torch.zeros(total)
    off = 0
    for i, p in enumerate(params):
        view = flat[off:off + p.numel()].view_as(p)
        view.fill_(float(10 * rank + i + 1))
        p.grad = view
        off += (p.numel() + 3) // 4 * 4
    loose = torch.nn.Parameter(torch.zeros(3))
    loose.grad = torch.full((3,), float(rank + 1))
    allreduce_gradients(params + [loose], weight=0.5)
    if rank == 0:
        ret['bucket'] = [p.grad.clone() for p in params] + [loose.grad.clone()]
        ret['same_storage'] = all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in params)
    dist.destroy_process_group()


def test_bucketed_gradients_are_reduced_in_place():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_bucket_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret['same_storage']
    for i, g in enumerate(ret['bucket'][:4]):       # 0.5 * ((i + 1) + (10 + i + 1))
        torch.testing.assert_close(g, torch.full_like(g, 0.5 * (2 * i + 12)))
    torch.testing.assert_close(ret['bucket'][4], torch.full((3,), 1.5))


class _BucketFn(torch.autograd.Function):
    """Stand-in for the drop-in's MLP backward: hands out the gradients of its parameters as views of one flat bucket."""

    @staticmethod
    def forward(ctx, x, *params):
        ctx.save_for_backward(x, *params)
        return sum((p * x).sum() for p in params)

    @staticmethod
    def backward(ctx, g):
        x, *params = ctx.saved_tensors
        total = sum((p.numel() + 3) // 4 * 4 for p in params)
        flat = torch.zeros(total)
        out, off = [], 0
        for p in params:
            view = flat[off:off + p.numel()].view_as(p)
            view.copy_(g * x.expand_as(p))
            out.append(view)
            off += (p.numel() + 3) // 4 * 4
        return (None,) + tuple(out)


def _exchange_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    blocks = [[torch.nn.Parameter(torch.ones(s)) for s in ((4, 3), (5,))], [torch.nn.Parameter(torch.ones(s)) for s in ((2, 2), (1,), (6,))]]
    params = [p for blk in blocks for p in blk]
    ex = GradientExchange(params, weight=0.5)
    got = []
    for step in range(3):                      # step 0 learns the buckets, steps 1-2 exchange from the hooks
        for p in params:
            p.grad = None
        x = torch.tensor(float(rank + 1 + step))
        (_BucketFn.apply(x, *blocks[0]) + _BucketFn.apply(x, *blocks[1])).backward()
        hooked = len(ex._handles)
        ex.finish()
        got.append((hooked, [p.grad.clone() for p in params]))
    ex.close()
    if rank == 0:
        ret['exchange'] = got
    dist.destroy_process_group()


def test_gradient_exchange_overlapped_with_backward():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_exchange_worker, args=(2, port, ret), nprocs=2, join=True)
    for step, (hooked, grads) in enumerate(ret['exchange']):
        assert hooked == (0 if step == 0 else 2)            # both buckets launched from the hooks after the first step
        want = 0.5 * ((1 + step) + (2 + step))              # weight * (x of rank 0 + x of rank 1)
        for g in grads:
            torch.testing.assert_close(g, torch.full_like(g, want))


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 4096, 762048):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_two_rank_gradient_equals_single_process():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    model, batch, target = _setup()
    _loss(model, batch, target).backward()
    for k, p in model.named_parameters():
        torch.testing.assert_close(ret['grads'][k], p.grad, rtol=1e-4, atol=1e-7, msg=lambda m, k=k: f'{k}: {m}')


# ---- masked means under ray sharding: count-weighted rank losses average to the global masked mean ----
def _masked_case():
    g = torch.Generator().manual_seed(8)
    n = 40
    pred = torch.rand((n, 3), generator=g)
    target = torch.rand((n, 3), generator=g)
    mask_nerf = torch.rand((n,), generator=g) < 0.7
    mask_nerf[:20] = torch.rand((20,), generator=g) < 0.3      # the two shards hold different numbers of masked rays
    return pred, target, mask_nerf, ~mask_nerf


def _mask_worker(rank, world, port, ret):
    from oracle import loss_oracle
    from simplenerf_b200.distributed import mask_count_weights
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    pred, target, m1, m2 = _masked_case()
    lo, hi = shard_bounds(pred.shape[0], rank, world)
    p = pred[lo:hi].clone().requires_grad_()
    scales = mask_count_weights({'indices_mask_nerf': m1[lo:hi], 'indices_mask_sparse_depth': m2[lo:hi]})
    loss = scales['indices_mask_nerf'] * loss_oracle.masked_mse(p, target[lo:hi], m1[lo:hi]) + \
        0.1 * scales['indices_mask_sparse_depth'] * loss_oracle.masked_mse(p, target[lo:hi], m2[lo:hi])
    loss.backward()
    grad = torch.zeros_like(pred)
    grad[lo:hi] = p.grad / world            # what averaging the ranks' parameter gradients does to each ray's contribution
    dist.all_reduce(grad)
    total = loss.detach().clone() / world
    dist.all_reduce(total)
    if rank == 0:
        ret['grad'], ret['loss'] = grad, total
    dist.destroy_process_group()


def test_count_weighted_masked_means_equal_the_global_mean():
    from oracle import loss_oracle
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_mask_worker, args=(2, port, ret), nprocs=2, join=True)
    pred, target, m1, m2 = _masked_case()
    p = pred.clone().requires_grad_()
    want = loss_oracle.masked_mse(p, target, m1) + 0.1 * loss_oracle.masked_mse(p, target, m2)
    want.backward()
    torch.testing.assert_close(ret['loss'], want.detach(), rtol=1e-6, atol=0)
    torch.testing.assert_close(ret['grad'], p.grad, rtol=1e-5, atol=1e-9)


# ---- gradient accumulation over sub-batches (Trainer01.py:84-101) under the overlapped exchange ----
def _accumulate_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    blocks = [[torch.nn.Parameter(torch.ones(s)) for s in ((4, 3), (5,))], [torch.nn.Parameter(torch.ones(s)) for s in ((2, 2), (1,), (6,))]]
    params = [p for blk in blocks for p in blk]
    ex = GradientExchange(params, weight=0.5)
    got = []
    n_sub = 3
    for step in range(3):
        for p in params:
            p.grad = None
        hooked_early = 0
        for k in range(n_sub):                 # every sub-batch accumulates into the same flat buckets
            ex.arm(k == n_sub - 1)
            x = torch.tensor(float(rank + 1 + step + 10 * k))
            (_BucketFn.apply(x, *blocks[0]) + _BucketFn.apply(x, *blocks[1])).backward()
            if k < n_sub - 1:
                hooked_early += len(ex._handles)
        hooked = len(ex._handles)
        ex.finish()
        got.append((hooked_early, hooked, [p.grad.clone() for p in params]))
    ex.close()
    if rank == 0:
        ret['accumulate'] = got
    dist.destroy_process_group()


def test_gradient_exchange_with_sub_batches_equals_single_process():
    """The buckets may be exchanged only in the LAST backward of a step: earlier sub-batches accumulate into them."""
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_accumulate_worker, args=(2, port, ret), nprocs=2, join=True)
    for step, (hooked_early, hooked, grads) in enumerate(ret['accumulate']):
        assert hooked_early == 0                              # nothing is launched before the last sub-batch
        assert hooked == (0 if step == 0 else 2)
        want = 0.5 * sum((1 + step + 10 * k) + (2 + step + 10 * k) for k in range(3))
        for g in grads:
            torch.testing.assert_close(g, torch.full_like(g, want))


def test_train_step_pieces_are_the_references_sub_batches_split_over_the_ranks():
    """every rank runs the same number of backward passes, also when shards are unequal."""
    from simplenerf_b200.trainer import RayShardedTrainStep
    for n, sub, world in ((4098, 1024, 2), (4099, 1024, 2), (4096, 2048, 8), (11, 4, 3), (4096, 4096, 1)):
        per_rank = []
        for rank in range(world):
            step = RayShardedTrainStep.__new__(RayShardedTrainStep)
            step.rank, step.world = rank, world
            per_rank.append(step._pieces(n, sub, shard=True))
        assert len({len(p) for p in per_rank}) == 1                      # same count everywhere
        covered = sorted(span for p in per_rank for span in p)
        assert covered[0][0] == 0 and covered[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))   # a partition of the batch
        for k in range(len(per_rank[0])):                                # piece k of every rank lies in the reference's sub-batch k
            assert all(k * sub <= p[k][0] < p[k][1] <= min(n, (k + 1) * sub) for p in per_rank)
    lone = RayShardedTrainStep.__new__(RayShardedTrainStep)          # a sub-batch with fewer rays than ranks cannot be split: loud, not a hang
    lone.rank, lone.world = 0, 2
    try:
        lone._pieces(4097, 1024, shard=True)
    except ValueError as exc:
        assert 'cannot be split' in str(exc)
    else:
        raise AssertionError('expected a ValueError')
