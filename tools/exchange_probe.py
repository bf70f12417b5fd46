"""How long does the gradient exchange itself take when the ranks arrive together?  (torchrun, one rank per GPU)
Four fp32 buckets of the model's sizes, all-reduced (a) as one coalesced NCCL launch, (b) one collective each, (c) as ONE flat
9 MB buffer; CUDA events on the compute stream around exchange + wait, after a tiny all-reduce that lines the ranks up."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
sizes = [595844, 595844, 579716, 494084]
bufs = [torch.randn(n, device=dev) for n in sizes]
flat = torch.randn(sum(sizes), device=dev)
tick = torch.zeros(1, device=dev)


def coalesced():
    with dist._coalescing_manager(async_ops=True) as cm:
        for b in bufs:
            dist.all_reduce(b, op=dist.ReduceOp.AVG)
    cm.wait()


def separate():
    hs = [dist.all_reduce(b, op=dist.ReduceOp.AVG, async_op=True) for b in bufs]
    for h in hs:
        h.wait()


def one_flat():
    dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=True).wait()


def sync_call():
    for b in bufs:
        dist.all_reduce(b, op=dist.ReduceOp.AVG)


for name, fn in (('coalesced', coalesced), ('separate', separate), ('one flat buffer', one_flat), ('blocking calls', sync_call)):
    ts = []
    for it in range(30):
        dist.all_reduce(tick)                 # line the ranks up
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[5:])
    t = torch.tensor([ts[len(ts) // 2]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f'{world} ranks, {name}: median {float(t):.0f} us (max over ranks) for {sum(sizes) * 4 / 1e6:.1f} MB')
dist.destroy_process_group()
