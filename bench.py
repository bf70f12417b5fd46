#!/usr/bin/env python
"""Benchmark of the SimpleNeRF volumetric-rendering hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config C2 of BASELINE.json): LLFF-shaped SimpleNeRF training step -- 1008x756 camera, 3 views,
4096 rays per GPU, 64 coarse + 128 extra fine samples (192-point fine pass), coarse + fine + points-augmented
+ views-augmented MLPs, forward + backward (+ Adam update), random-init weights, synthetic rays.
A "step" is one pass of the hot path over one 4096-ray batch per GPU (weak scaling: the reference's
RealEstate config C4 is 32768 rays over 8 GPUs = 4096 per GPU); under torchrun the MLP gradients are
all-reduced over NCCL every step.  Metric: train rays/s over all GPUs.

Our arm   : the drop-in model (simplenerf_b200.models.FusedSimpleNeRF01, bf16 tcgen05 path) -- `value` with
            the batch resident in HBM, `e2e` with the batch in pinned host memory copied every step and the
            loss read back.
Reference : `--impl reference` times the UNMODIFIED reference -- its own Trainer01.train_one_iter on the same C2
            step (DataPreprocessor01 batch, SimpleNeRF01, LossComputer01 with the nine shipped losses, Adam) -- on the host
            cores, from the copy `__graft_entry__.build()` stages under baseline/_ref (kind "reference"); when that copy is
            absent it falls back to the CPU restatement in oracle/ (kind "port").
Beside the headline the line carries: `roofline` (MLP kernels against the bf16 tensor peak, per-kernel HBM fractions),
`cpu_baseline`, `render` (ms per 1008x756 frame), `c5` (compositing / resampling sweep against the HBM roof),
`trainer` (the drop-in behind the reference's own Trainer01 on this GPU, and the reference model itself on this GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096


def measured_traffic():
    """DRAM bytes of the three MLP kernels per 4096-ray step (dram__bytes_read.sum + dram__bytes_write.sum over the 12 launches
    of one step, ncu --set full): kept in profiles/mlp_dram_traffic.json next to the capture it was read from, so that the
    number changes when the profile does.  Algorithmic figure: 28.8 GB."""
    path = os.path.join(ROOT, 'profiles', 'mlp_dram_traffic.json')
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d['bytes_per_step']), f"{d['source']} ({d['how']})"
    except (OSError, KeyError, ValueError):
        return None, 'profiles/mlp_dram_traffic.json missing'
# algorithmic GB per step and kernel: main / fine / pts-aug MLPs keep 8.5 panels x 512 B per point
# (8 trunk activations + the 128-wide view layer: 4.25 KB; the feature vector is never formed -- merged view branch) + 0.25 KB
# of sign bits, views-aug 8 panels (4 KB); 4096 rays x (64 + 64 + 192) points with view layers + 4096 x 64 without.
_P_VIEW, _P_NOVIEW = 4096 * (64 + 64 + 192), 4096 * 64
ALG_GB = {'forward': (_P_VIEW * (4352 + 256) + _P_NOVIEW * (4096 + 256)) / 1e9,
          'dgrad': (_P_VIEW * (4352 + 256 + 256) + _P_NOVIEW * (4096 + 256)) / 1e9,
          'wgrad': (_P_VIEW * 72 * 128 + _P_NOVIEW * 68 * 128) / 1e9}
TRAIN_FLOP_PER_RAY = 2 * 648_585_216          # fwd+bwd MACs per ray (4 MLPs) x 2 -- the REFERENCE's arithmetic
# what the kernels execute: the merged view branch saves 65 536 MACs per point in each of forward, dgrad and wgrad of the three
# MLPs with a view branch (64 + 64 + 192 points per ray)
TRAIN_FLOP_PER_RAY_EXECUTED = 2 * (648_585_216 - 3 * 65_536 * (64 + 64 + 192))
RENDER_FLOP_PER_RAY_EXECUTED = 2 * 256 * (593_408 - 65_536)
RENDER_FLOP_PER_RAY = 2 * 256 * 593_408       # vanilla coarse+fine eval
STREAMS = (('rgb_coarse', 'depth_coarse'), ('rgb_fine', 'depth_fine'),
           ('points_augmentation_rgb_coarse', 'points_augmentation_depth_coarse'),
           ('views_augmentation_rgb_coarse', 'views_augmentation_depth_coarse'))


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(tflops=d.get('bf16_tflops_sustained', d['bf16_tflops']), tflops_burst=d['bf16_tflops'],
                    hbm=d['hbm_gbs'], source='measured (MEASURED_PEAKS.json: burst bf16 = best of 10 matmuls, sustained = 4 s back to back)')
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """Samples SM clocks / throttle reasons during the timed region (NVML; falls back to nvidia-smi)."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap',
               0x80: 'hw_power_brake_slowdown', 0x2: 'applications_clocks_setting'}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def _sample(self):
        import pynvml
        self.samples.append(pynvml.nvmlDeviceGetClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _run(self):
        try:
            while not self._stop.is_set():
                self._sample()
                time.sleep(0.005)
        except Exception as exc:   # noqa: BLE001
            self.reasons.add(f'sampler_error:{type(exc).__name__}')

    def __enter__(self):
        try:   # NVML is initialised before the timed region starts, so the first sample falls inside it
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as exc:   # noqa: BLE001
            self.reasons.add(f'sampler_error:{type(exc).__name__}')
        return self

    def mark(self):
        """One synchronous sample (called while the GPU is still busy, just before the closing synchronize)."""
        try:
            self._sample()
        except Exception as exc:   # noqa: BLE001
            self.reasons.add(f'sampler_error:{type(exc).__name__}')

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def make_state(model, seed=0):
    """Deterministic random-init weights (an opaque field) for any module with the reference's names."""
    from simplenerf_b200 import synthetic
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    return synthetic.densify_state(synthetic.deterministic_state(shapes, seed))


def training_loss(out, target, tdepth):
    return sum(((out[a] - target) ** 2).mean() + 0.1 * ((out[b] - tdepth) ** 2).mean() for a, b in STREAMS)


def fused_training_loss(out, target, tdepth):
    """The same function as training_loss in one forward and one backward launch (snerf_ray_losses_*: MSE01-03 /
    SparseDepthMSE01-03 semantics with every ray masked in): 8 streams, rgb weight 1, depth weight 0.1."""
    from simplenerf_b200.loss_functions import ray_losses
    preds = [out[a] for a, _ in STREAMS] + [out[b] for _, b in STREAMS]
    n = len(STREAMS)
    return ray_losses(preds, [target] * n + [tdepth] * n, [None] * (2 * n), [1.0] * n + [0.1] * n)[-1]


# ------------------------------------------------------------------------------------------------
# reference arm: the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_training_pass(n_rays: int, sub_batch: int, threads: int):
    """One training pass of the oracle over n_rays (grad accumulation in sub-batches like Trainer01.py:82-101)."""
    import torch
    from oracle import nerf_oracle as orc
    from simplenerf_b200 import synthetic
    torch.set_num_threads(threads)
    configs = synthetic.make_configs('simplenerf')
    model = orc.NerfOracle(configs)
    model.load_state_dict(make_state(model))
    model.train()
    batch = synthetic.make_ray_batch('llff', n_rays, 1021)
    g = torch.Generator().manual_seed(3)
    target, tdepth = torch.rand((n_rays, 3), generator=g), 1 + 4 * torch.rand((n_rays,), generator=g)

    def run():
        t0 = time.perf_counter()
        for i in range(0, n_rays, sub_batch):
            sub = {k: (v[i:i + sub_batch] if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            out = model(sub)
            training_loss(out, target[i:i + sub_batch], tdepth[i:i + sub_batch]).backward()
        return time.perf_counter() - t0
    return run


def reference_step_runner(threads: int):
    """-> (run, kind, sample): run() performs ONE full C2 training step of the reference on the host cores and returns seconds.
    Unmodified reference when staged (baseline/_ref), else the oracle port."""
    import torch
    from baseline import ref_harness as rh
    torch.set_num_threads(threads)
    if rh.available():
        trainer, _ = rh.build_trainer('SimpleNeRF01', device=[0], resolution=(756, 1008), num_rays=2048, sparse_rays=2048,
                                      sub_batch_size=2048)
        if next(trainer.model.parameters()).is_cuda:
            raise RuntimeError('the CPU arm must run with the GPUs hidden (CUDA_VISIBLE_DEVICES="")')
        it = [20000]

        def run():
            t0 = time.perf_counter()
            trainer.train_one_iter(it[0])
            it[0] += 1
            return time.perf_counter() - t0
        return run, 'reference', ('one full C2 step per step: the unmodified Trainer01.train_one_iter (src/Trainer01.py:61-107) on a synthetic '
                                  '3-view 756x1008 scene -- 2048 image + 2048 sparse-depth rays as 2 sub-batches of 2048, SimpleNeRF01 (4 MLPs), '
                                  'LossComputer01 with the 9 shipped losses, Adam -- from baseline/_ref, fp32, all host threads')
    run = cpu_training_pass(4096, 2048, threads)
    return run, 'port', ('one full 4096-ray step per step as 2 sub-batches of 2048 (fwd+bwd, 4 MLPs): oracle/nerf_oracle.py on the host '
                         '(baseline/_ref not staged)')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ['CUDA_VISIBLE_DEVICES'] = ''          # the reference picks the GPU by itself when it sees one (CommonUtils01.py:15-27)
    import torch
    threads = os.cpu_count() or 1
    run, kind, sample = reference_step_runner(threads)
    for _ in range(args.warmup):
        run()
    times = [run() for _ in range(args.steps)]
    total = sum(times)
    value = RAYS_PER_GPU * args.steps / total
    line = {
        'impl': 'reference', 'metric': 'train rays/s (64+128 samples, fwd+bwd)', 'value': value, 'unit': 'rays/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'rays/s', 'cores': torch.get_num_threads(), 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': 'rays/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(quick: bool):
    """The reference arm on the host cores, in a child process with the GPUs hidden: 1 warm-up + 2 timed full steps (~15-25 s)."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1' if quick else '2', '--warmup', '0' if quick else '1']
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE')}
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    for ln in reversed(res.stdout.strip().splitlines()):
        if ln.startswith('{'):
            return json.loads(ln)['cpu_baseline']
    return {'value': None, 'unit': 'rays/s', 'cores': os.cpu_count(), 'kind': 'failed', 'sample': (res.stderr or res.stdout)[-300:]}


def trainer_context(steps: int, warmup: int):
    """N=1 context numbers through the reference's own Trainer01.train_one_iter on THIS GPU (full C2 step incl. the reference's
    batch assembly and its per-loss .item() synchronisations): (a) the unmodified reference model -- what a user of the
    reference runs today; (b) the same with ONE string changed (model name -> FusedSimpleNeRF01); (c) also FusedLossComputer +
    FusedAdam.  None when the reference copy is not staged."""
    import torch
    from baseline import ref_harness as rh
    if not rh.available():
        return None
    from simplenerf_b200.loss_functions.FusedLossComputer01 import FusedLossComputer
    from simplenerf_b200.optim import FusedAdam
    scene = dict(device=[torch.cuda.current_device()], resolution=(756, 1008), num_rays=2048, sparse_rays=2048, sub_batch_size=2048)
    out = {'note': 'rays/s of Trainer01.train_one_iter on one B200, wall clock with a device synchronize on both sides; 4096 rays per step'}
    variants = (('reference_gpu', 'SimpleNeRF01', {}, {}),
                ('dropin_one_string', 'FusedSimpleNeRF01', {'precision': 'bf16'}, {}),
                ('dropin_fused_losses_adam', 'FusedSimpleNeRF01', {'precision': 'bf16'},
                 dict(loss_computer_factory=FusedLossComputer, optimizer_factory=lambda ps: FusedAdam(ps, lr=5e-4, betas=(0.9, 0.999)))))
    for key, name, extra, factories in variants:
        trainer, _ = rh.build_trainer(name, model_extra=extra or None, **scene, **factories)
        for i in range(warmup):
            trainer.train_one_iter(20000 + i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            trainer.train_one_iter(20100 + i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        out[key] = {'value': RAYS_PER_GPU / dt, 'unit': 'rays/s', 'ms_per_step': 1e3 * dt, 'steps': steps}
        if key == 'dropin_fused_losses_adam':
            # (d) the same model, losses and optimizer driven by the drop-in's own per-rank step (simplenerf_b200.trainer.
            # RayShardedTrainStep = the body of Trainer01.train_one_iter without its ten .item() reads per sub-batch) on a batch the
            # reference's preprocessor assembled, resident on the device: device time of the full C2 step with the shipped losses
            from simplenerf_b200.trainer import RayShardedTrainStep
            batch = trainer.train_data_loader.get_next_batch(20200)
            model = trainer.model.module if hasattr(trainer.model, 'module') else trainer.model
            step = RayShardedTrainStep(trainer.configs, model, trainer.loss_computer, trainer.optimizer)
            for _ in range(max(warmup, 3)):
                step(batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4 * steps):
                step(batch)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (4 * steps)
            out['train_full'] = {'value': RAYS_PER_GPU / (ms * 1e-3), 'unit': 'rays/s', 'ms_per_step': ms, 'steps': 4 * steps,
                                 'what': 'RayShardedTrainStep: 2 sub-batches of 2048 rays (2048 image + 2048 sparse-depth rays), FusedSimpleNeRF01 bf16, '
                                         'FusedLossComputer with the 9 shipped losses, FusedAdam; CUDA events, batch resident in HBM'}
        del trainer
        torch.cuda.empty_cache()
    return out


def workload_config(n_gpus: int):
    return {'workload': 'C2: LLFF-shaped SimpleNeRF training step, 3 views 1008x756, 4096 rays/GPU, 64 coarse + 192 fine '
                        'points/ray, coarse+fine+points-aug+views-aug MLPs, fwd+bwd+Adam',
            'rays_per_gpu': RAYS_PER_GPU, 'global_rays': RAYS_PER_GPU * n_gpus, 'parallelism': f'ray-sharded dp{n_gpus}',
            'l2': 'per-step working set (bf16 activation + gradient stash, ~14 GB) exceeds the 126 MB L2'}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from simplenerf_b200 import _lib, ops, synthetic
    from simplenerf_b200.distributed import GradientExchange
    from simplenerf_b200.models import get_model

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    if not _lib.load().snerf_has_tensor_path():
        raise RuntimeError('libsimplenerf_b200.so was built without the tensor path')

    configs = synthetic.make_configs('simplenerf')
    model = get_model(configs, None)
    model.load_state_dict(make_state(model))
    model = model.to(dev).train()
    from simplenerf_b200.optim import FusedAdam
    opt = (torch.optim.Adam(model.parameters(), lr=5e-4, betas=(0.9, 0.999), fused=True) if args.torch_adam
           else FusedAdam(model.parameters(), lr=5e-4, betas=(0.9, 0.999)))    # Trainer01.py:516, one launch
    params = [p for p in model.parameters()]
    exchange = GradientExchange(params, weight=1.0 / world, overlap=args.exchange == 'overlap') if world > 1 else None
    n = RAYS_PER_GPU
    host = synthetic.make_ray_batch('llff', n, 1021 + rank)
    g = torch.Generator().manual_seed(3 + rank)
    host['target_rgb'] = torch.rand((n, 3), generator=g)
    host['target_depth'] = 1 + 4 * torch.rand((n,), generator=g)
    host = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    resident = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))

    def step(batch):
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        loss = (training_loss if args.torch_loss else fused_training_loss)(out, batch['target_rgb'], batch['target_depth'])
        loss.backward()
        if exchange is not None:   # ray-sharded data parallel: sum of shard gradients / world == gradient of the global mean loss;
            exchange.finish()      # the buckets were launched from autograd hooks while the backward pass was still running
        opt.step()
        return loss

    host_ms = {'last': 0.0}

    def timed(fn, steps, mark=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms['last'] = 1e3 * (time.perf_counter() - t_host) / max(steps, 1)   # host time to ENQUEUE a step (no sync inside)
        e1.record()
        if mark is not None:
            mark()   # the host runs ahead of the device: the GPU is still inside the timed region here
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    n_warm = args.warmup        # exactly the W untimed steps the caller asked for
    for _ in range(n_warm):
        step(resident)

    # ---- value: inputs resident in HBM; MLP kernels timed with CUDA events inside the timed region ----
    events = {'mlp_forward': [], 'mlp_backward': [], 'dgrad': [], 'wgrad': []}

    class Timer:
        def __init__(self, name):
            self.name = name
            self.split_event = None

        def __enter__(self):
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if self.name == 'mlp_backward':      # the C side records this event between the dgrad and the wgrad launch
                self.split_event = torch.cuda.Event(enable_timing=True)
                self.split_event.record()        # materialises the handle
            self.e0.record()
            return self

        def __exit__(self, *exc):
            self.e1.record()
            events[self.name].append((self.e0, self.e1))
            if self.split_event is not None:
                events['dgrad'].append((self.e0, self.split_event))
                events['wgrad'].append((self.split_event, self.e1))
            return False

    ops.TIMER['hook'] = Timer
    launches0 = ops.LAUNCHES['count']
    with ClockSampler(local) as clocks:
        ms_total = timed(lambda: step(resident), args.steps, clocks.mark)
    launches = ops.LAUNCHES['count'] - launches0
    ops.TIMER['hook'] = None
    mlp_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in events.items()}
    ms_step = ms_total / args.steps
    host_enqueue_ms = host_ms['last']
    value = world * n / (ms_step * 1e-3)

    # ---- e2e: batch in pinned host memory, copied every step; loss read back every step ----
    # The loss of every step is copied to pinned host memory on the compute stream (non-blocking, like the input copies),
    # so the host keeps enqueueing the next step while the device works; the synchronize that closes the timed region
    # covers all the reads.
    loss_host = torch.zeros(max(args.steps, 1) + 1, dtype=torch.float32).pin_memory()
    e2e_i = [0]

    # Input pipeline: HostBatchStager = one pinned staging buffer and ONE host -> device copy per step (not one per tensor), on a
    # copy stream, double buffered: while step i runs, the host packs batch i+1 and its copy travels under step i's kernels.
    from simplenerf_b200.batching import HostBatchStager
    stager = HostBatchStager(host, dev, slots=2)

    def e2e_step():
        slot = e2e_i[0] % 2
        batch = stager.device_batch(slot)                      # this step's inputs: copied during the previous step
        stager.stage((slot + 1) % 2, host)                     # pack + submit the NEXT step's inputs (pinned host -> device)
        loss = step(batch)
        stager.release(slot)
        loss_host[e2e_i[0] % loss_host.numel()].copy_(loss.detach(), non_blocking=True)     # device -> host read of the loss
        e2e_i[0] += 1
    stager.stage(0, host)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    e2e_value = world * n / (ms_e2e * 1e-3)
    if not bool(torch.isfinite(loss_host[:min(e2e_i[0], loss_host.numel())]).all()):
        raise RuntimeError('non-finite loss read back in the e2e run')

    # ---- secondary metric: ms per 1008x756 frame (vanilla coarse+fine eval, tile-sharded rows, no collective) ----
    render = None
    if not args.no_render:
        vcfg = synthetic.make_configs('vanilla')
        vmodel = get_model(vcfg, None)
        vmodel.load_state_dict(make_state(vmodel))
        vmodel = vmodel.to(dev).eval()
        h, w = synthetic.CAMERAS['llff']['resolution']
        rows = (h + world - 1) // world
        r0, r1 = rank * rows, min(h, (rank + 1) * rows)
        frame = synthetic.make_ray_batch('llff', (r1 - r0) * w, 0, frame=True, start=r0 * w)
        frame = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in frame.items()}
        with torch.no_grad():
            vmodel(frame)
            ms_frame = timed(lambda: vmodel(frame), args.frames) / args.frames
        # the same frame through the one-call renderer: pose on the host -> device ray generation -> render ->
        # device post-processing -> uint8 image and depth maps in pinned host memory, all inside the timed region
        from simplenerf_b200.render import FrameRenderer
        cam = synthetic.CAMERAS['llff']
        intrinsic = [[cam['focal'], 0, cam['centre'][0]], [0, cam['focal'], cam['centre'][1]], [0, 0, 1]]
        fr = FrameRenderer(vmodel, (h, w), intrinsic, cam['near'], cam['far'], rows=(r0, r1))
        import numpy as np
        pose = np.eye(4, dtype=np.float32)
        fr.render(pose)
        ms_frame_e2e = timed(lambda: fr.render(pose), args.frames) / args.frames
        d2h = (r1 - r0) * w * (3 + 4 * 4)
        render = {'ms_per_frame': ms_frame, 'rays_per_s': h * w / (ms_frame * 1e-3), 'resolution': [h, w], 'frames_timed': args.frames,
                  'tensor_frac_of_burst_peak': h * w * RENDER_FLOP_PER_RAY / (ms_frame * 1e-3) / (peaks()['tflops_burst'] * 1e12 * world),
                  'tensor_frac_of_sustained_peak': h * w * RENDER_FLOP_PER_RAY / (ms_frame * 1e-3) / (peaks()['tflops'] * 1e12 * world),
                  'tensor_frac_of_burst_peak_executed': h * w * RENDER_FLOP_PER_RAY_EXECUTED / (ms_frame * 1e-3) / (peaks()['tflops_burst'] * 1e12 * world),
                  'sharding': f'{world} row bands, no collective',
                  'e2e': {'ms_per_frame': ms_frame_e2e, 'h2d_bytes_per_frame': 48 + 36, 'd2h_bytes_per_frame': d2h,
                          'note': 'FrameRenderer.render(pose): rays generated on the device, uint8 image + 4 depth maps copied to pinned host memory'}}

    # ---- N=1 extras (rank 0): C5 sweep, the drop-in behind the reference's Trainer01, the reference's CPU path ----
    c5 = trainer_ctx = cpu = None
    if rank == 0 and world == 1:
        if not args.no_c5:
            from tools import scan_microbench
            rows = scan_microbench.run(iters=5)
            c5 = {'unit': 'GB/s', 'peak': peaks()['hbm'], 'target_frac': 0.70,
                  'bytes': 'algorithmic bytes per ray (SURVEY.md 8d): composite fwd 20S+68 (24S+68 with weights), bwd 36S+48, sample_pdf+merge 1792 (1280 with the linspace row), stratified 8S',
                  'rows': [{'kernel': r['kernel'], 'S': r['S'], 'log2_rays': r['rays'].bit_length() - 1, 'us': round(r['seconds'] * 1e6, 1),
                            'gbs': round(r['gbs']), 'frac': round(r['frac'], 3)} for r in rows]}
        if not args.no_trainer:
            trainer_ctx = trainer_context(steps=5, warmup=2)
        if not args.no_cpu:
            cpu = cpu_baseline_subprocess(args.quick_cpu)

    if rank == 0:
        pk = peaks()
        mlp_total = mlp_ms['mlp_forward'] + mlp_ms['mlp_backward']
        achieved = n * TRAIN_FLOP_PER_RAY / (mlp_total * 1e-3) / 1e12
        traffic, traffic_note = measured_traffic()
        line = {
            'metric': 'train rays/s (64+128 samples, fwd+bwd)', 'value': value, 'unit': 'rays/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': n_warm, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': workload_config(world),
            'e2e': {'value': e2e_value, 'unit': 'rays/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': stager.nbytes,
                    'd2h_bytes_per_step': 4,
                    'note': 'inputs: packed into one pinned buffer and copied host -> device every step on a copy stream (HostBatchStager, double buffered: the copy of step i+1 overlaps step i); loss: device -> pinned host every step (non-blocking copy on the compute stream)'},
            'gpu_launches': launches,
            'host_enqueue_ms_per_step': host_enqueue_ms,
            'clocks': clocks.summary(),
            'roofline': {'bound': 'tensor', 'kernel': 'tc_forward_kernel + tc_dgrad_kernel + tc_wgrad_kernel (all 4 MLPs)',
                         # `peak` is the BURST figure (the timed region is short and runs above the clock the sustained figure was
                         # measured at); the fraction of the sustained figure is given beside it
                         'achieved': achieved, 'peak': pk['tflops_burst'], 'unit': 'TFLOP/s', 'frac': achieved / pk['tflops_burst'],
                         'frac_of_sustained_peak': achieved / pk['tflops'], 'sustained_peak': pk['tflops'],
                         # `achieved` counts the reference's arithmetic (1.2972 GFLOP per ray).  The kernels execute
                         # 9.7 % fewer MACs (feature_linear folded into the view layer): the fraction of the
                         # tensor peak the hardware actually sustains is the `executed` one
                         'executed': {'flop_per_ray': TRAIN_FLOP_PER_RAY_EXECUTED,
                                      'achieved': achieved * TRAIN_FLOP_PER_RAY_EXECUTED / TRAIN_FLOP_PER_RAY,
                                      'frac': achieved * TRAIN_FLOP_PER_RAY_EXECUTED / TRAIN_FLOP_PER_RAY / pk['tflops_burst']},
                         'traffic': traffic, 'peak_source': pk['source'],
                         'traffic_note': 'DRAM bytes per step of the same three kernels (12 launches): ' + traffic_note,
                         # the training step moves ~28.5 GB through HBM for 4.8 executed TFLOP: it sits between the two roofs
                         'hbm': None if traffic is None else {'achieved': traffic / (mlp_total * 1e-3) / 1e9, 'peak': pk['hbm'], 'unit': 'GB/s',
                                                              'frac': traffic / (mlp_total * 1e-3) / 1e9 / pk['hbm']},
                         # per kernel, against the roof that bounds it in training: ALGORITHMIC bytes (DESIGN.md section 4:
                         # KB per point x 1 572 864 points of a step; views-aug has no view layer) / live CUDA-event time
                         'kernels': {k: {'bound': 'hbm', 'ms_per_step': mlp_ms[t], 'algorithmic_gb': gb,
                                         'achieved': gb / (mlp_ms[t] * 1e-3), 'peak': pk['hbm'], 'unit': 'GB/s',
                                         'frac': gb / (mlp_ms[t] * 1e-3) / pk['hbm']}
                                     for k, t, gb in (('tc_forward_kernel', 'mlp_forward', ALG_GB['forward']),
                                                      ('tc_dgrad_kernel', 'dgrad', ALG_GB['dgrad']),
                                                      ('tc_wgrad_kernel', 'wgrad', ALG_GB['wgrad'])) if mlp_ms.get(t)},
                         'ms_per_step': {'mlp_forward': mlp_ms['mlp_forward'], 'mlp_backward': mlp_ms['mlp_backward'],
                                         'other': ms_step - mlp_total},
                         'frac_forward': n * 2 * 220_348_416 / (mlp_ms['mlp_forward'] * 1e-3) / 1e12 / pk['tflops_burst'],
                         'frac_backward': n * 2 * 428_236_800 / (mlp_ms['mlp_backward'] * 1e-3) / 1e12 / pk['tflops_burst'],
                         'step_frac_of_peak': n * TRAIN_FLOP_PER_RAY / (ms_step * 1e-3) / 1e12 / pk['tflops_burst']},
            'cpu_baseline': cpu,
            'render': render,
            'c5': c5,
            'trainer': trainer_ctx,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-render', action='store_true')
    ap.add_argument('--frames', type=int, default=10, help='frames timed for the render metric')
    ap.add_argument('--no-c5', action='store_true', help='skip the compositing / resampling sweep (N=1 only)')
    ap.add_argument('--no-trainer', action='store_true', help='skip the runs behind the reference Trainer01 (N=1 only)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--quick-cpu', action='store_true')
    ap.add_argument('--exchange', default='after', choices=['after', 'overlap'],
                    help='gradient all-reduce: one coalesced launch after the backward pass (default) or per bucket from autograd hooks while it runs')
    ap.add_argument('--torch-loss', action='store_true', help='eager torch loss (8 masked means, ~80 launches) instead of snerf_ray_losses_*')
    ap.add_argument('--torch-adam', action='store_true', help='torch.optim.Adam(fused=True) instead of simplenerf_b200.optim.FusedAdam')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
