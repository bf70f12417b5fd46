"""Next-row N3: batch assembly.  The fixtures come from the unmodified reference methods (oracle/make_golden_batch.py);
gathers are pure data movement, so everything is compared bit for bit."""
import pytest
import torch

import golden_util as gu
from oracle import batch_oracle


def _case(g, tag, device='cpu'):
    nerf = {k[len(tag) + 6:]: v.to(device) for k, v in g.items() if k.startswith(f'{tag}_nerf_')}
    sd = {k[len(tag) + 4:]: v.to(device) for k, v in g.items() if k.startswith(f'{tag}_sd_')}
    idx = {k[len(tag) + 5:]: v.to(device) for k, v in g.items() if k.startswith(f'{tag}_idx_')}
    want = {k[len(tag) + 7:]: v for k, v in g.items() if k.startswith(f'{tag}_batch_')}
    return nerf, sd, idx, want


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_batch_oracle_matches_reference(tag):
    nerf, sd, idx, want = _case(gu.load('batch.npz'), tag)
    got = batch_oracle.assemble_batch(nerf, sd, idx['indices'], idx['indices_mask_nerf'], idx.get('indices_mask_sparse_depth'))
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k], want[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_assemble_batch_matches_reference(tag):
    from simplenerf_b200.batching import assemble_batch
    nerf, sd, idx, want = _case(gu.load('batch.npz'), tag, 'cuda:0')
    tables = dict(nerf, sparse_depth_data=sd)
    got = assemble_batch(tables, idx['indices'], idx['indices_mask_nerf'], idx.get('indices_mask_sparse_depth'), ndc=True)
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and torch.equal(got[k].cpu(), want[k]), k


@pytest.mark.gpu
def test_gather_rows_many_tables_and_empty_batch():
    from simplenerf_b200.batching import gather_rows
    g = torch.Generator().manual_seed(4)
    idx = torch.randint(0, 1000, (3001,), generator=g).cuda()
    mask = (torch.rand(3001, generator=g) < 0.5).cuda()
    entries = []
    for k in range(30):                                   # more tables than one launch takes
        src = torch.rand((1000, 1 + k % 5), generator=g).cuda()
        entries.append((src, torch.empty((3001, 1 + k % 5), device='cuda'), mask if k % 2 else None))
    gather_rows(entries, idx)
    for src, dst, m in entries:
        want = src[idx] if m is None else torch.where(m[:, None], src[idx], torch.full_like(src[idx], -1))
        assert torch.equal(dst, want)
    gather_rows([(entries[0][0], torch.empty((0, 1), device='cuda'), None)], idx[:0])


def test_gather_refuses_cpu_tensors():
    from simplenerf_b200.batching import gather_rows
    with pytest.raises(RuntimeError, match='CUDA'):
        gather_rows([(torch.zeros(4, 3), torch.zeros(2, 3), None)], torch.zeros(2, dtype=torch.int64))


@pytest.mark.gpu
def test_host_batch_stager_round_trip_and_slot_reuse():
    from simplenerf_b200 import synthetic
    from simplenerf_b200.batching import HostBatchStager
    example = synthetic.make_ray_batch('llff', 777, 3)
    example['pixel_id'] = torch.randint(0, 1000, (777, 3), dtype=torch.int32)
    stager = HostBatchStager(example, 'cuda:0', slots=2)
    assert stager.bytes_per_batch == sum(v.numel() * v.element_size() for v in example.values() if isinstance(v, torch.Tensor))
    for it in range(5):                       # slots are reused; every batch must arrive intact
        batch = {k: (v + it if isinstance(v, torch.Tensor) else v) for k, v in example.items()}
        slot = it % 2
        stager.stage(slot, batch)
        got = stager.device_batch(slot)
        assert got['iter_num'] == example['iter_num'] and got['num_frames'] == 3
        for k, v in batch.items():
            if isinstance(v, torch.Tensor):
                assert got[k].dtype == v.dtype and got[k].shape == v.shape and torch.equal(got[k].cpu(), v), k
        stager.release(slot)
