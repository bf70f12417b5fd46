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
from simplenerf_b200.distributed import allreduce_gradients, shard_bounds, shard_rays

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
