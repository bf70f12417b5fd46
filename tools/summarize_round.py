"""Turns the files a profiling pass left in gpurun_out/ (tools/final_round.sh) into the tracked summaries under profiles/.
    python tools/summarize_round.py <tag of the ncu files, e.g. r1i> <tag of the bench files, e.g. final3> <output tag, e.g. r1_final3>"""
import csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
ncu_tag, bench_tag, out_tag = sys.argv[1:4]


def last_json(name):
    path = os.path.join(G, name)
    return json.loads(open(path).read().strip().splitlines()[-1]) if os.path.exists(path) else None


# ---- scan / loss / gather kernels under ncu --set full ----
rows = list(csv.reader(open(os.path.join(G, f'{ncu_tag}_aux_raw.csv'))))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda d, k: float(d[ix[k]].replace(',', ''))                                        # noqa: E731
to_bytes = lambda v, u: v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]        # noqa: E731
n = 1 << 20
alg = {'composite_fwd_kernel<4': (24 * 64 + 68) * n, 'composite_bwd_kernel<4': (36 * 64 + 48) * n,
       'composite_fwd_kernel<6': (24 * 192 + 68) * n, 'composite_bwd_kernel<6': (36 * 192 + 48) * n, 'sample_coarse': 8 * 64 * n}
out = [f'# Round 1 ({out_tag}): samplers, compositing, losses and batch gather under `ncu --set full --clock-control none`', '',
       'Command (tools/final_round.sh): `ncu --set full --clock-control none --import-source on -k "regex:sample_|composite_|ray_losses|reproj_|gather_rows" python tools/aux_kernels_prof.py`',
       "(2^20 rays for the samplers / compositing at 64 and 192 samples per ray = the model's coarse and fine passes; 4096 / 1500 rays for the loss "
       'kernels; 4096 rows of 4 tables for the gather).  ncu times are cold-cache and serialised; the bandwidth fractions of the C5 table come from '
       f'CUDA-event timing (profiles/{out_tag}_c5_scan_microbench.md).', '',
       '| kernel | grid | time us | dram read MB | dram write MB | algorithmic MB | dram GB/s | L1 data pipe % | issue slots % | warps active % | smem wavefronts / ray | instructions / ray | regs |',
       '|---|---|---|---|---|---|---|---|---|---|---|---|---|']
fine_seen = 0
for d in data:
    short = d[ix['Kernel Name']].replace('void ', '').replace('snerf::', '').split('(')[0]
    t_us = num(d, 'gpu__time_duration.sum') * {'ms': 1e3, 'us': 1, 'ns': 1e-3, 's': 1e6}[units[ix['gpu__time_duration.sum']]]
    rd = to_bytes(num(d, 'dram__bytes_read.sum'), units[ix['dram__bytes_read.sum']])
    wr = to_bytes(num(d, 'dram__bytes_write.sum'), units[ix['dram__bytes_write.sum']])
    a = next((v for k, v in alg.items() if k in short), None)
    per_ray = ['', '']
    if 'sample_fine' in short:
        a = (1792 if fine_seen == 0 else 1280) * n
        short += ' (random u)' if fine_seen == 0 else ' (linspace row)'
        fine_seen += 1
        per_ray = [f"{num(d, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / n:.0f}", f"{num(d, 'smsp__inst_executed.sum') / n:.0f}"]
    out.append(f"| `{short}` | {d[ix['launch__grid_size']]} | {t_us:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {'' if a is None else f'{a / 1e6:.1f}'} | "
               f"{(rd + wr) / t_us / 1e3:.0f} | {num(d, 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'):.0f} | "
               f"{num(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {num(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
               f"{per_ray[0]} | {per_ray[1]} | {d[ix['launch__registers_per_thread']]} |")
out += ['', 'Reading:',
        "* compositing and the stratified sampler move exactly their algorithmic bytes (DRAM read + write = the table's algorithmic MB within 2 %) and are HBM-bound;",
        '* `sample_fine_fast_kernel` is NOT HBM-bound: its DRAM traffic equals the algorithmic 1.8 / 1.3 KB per ray, but it needs several hundred warp '
        'instructions and ~120-190 shared-memory wavefronts (loads, stores, shuffles) per ray, and ncu shows issue slots and the L1 data pipe both at '
        '75-85 %.  History of the round for 2^22 rays (CUDA events): 6.37 -> 3.74 ms with random uniforms, 4.50 -> 2.42 ms with the linspace row '
        '(branch-free descents over breadth-first tables: -117 bank-conflict cycles per ray; all-ascending sorting network; merge from an OR-reduced '
        'occupancy mask; zero numerators kept out of the division slow path).  What is left is the exact-arithmetic floor (fp64 prefix sums, IEEE '
        'divisions, 128-element sort, 192-way merge): 0.31-0.43 of the HBM roof;',
        '* the loss and gather kernels are launch-latency sized (5-25 us at 4096 rays): they exist to replace a few hundred eager launches and the '
        'device synchronisations of boolean-mask indexing, not to move bytes.']
open(os.path.join(P, f'{out_tag}_aux_kernels_ncu.md'), 'w').write('\n'.join(out) + '\n')

# ---- bench lines ----
n1 = last_json(f'{bench_tag}_n1.json')
lines = [('`python bench.py` (N=1, 20 steps, 25 warm-up)', n1)]
for g in (2, 4, 8):
    d = last_json(f'{bench_tag}_n{g}.json')
    if d is not None:
        lines.append((f'`torchrun --nproc-per-node {g} bench.py --gpus {g} --steps 20 --warmup 5 --no-cpu`', d))
ref = last_json(f'{bench_tag}_ref.json')
txt = [f'# Round 1 ({out_tag}): bench lines (one B200 box each; SM clocks / throttle reasons as sampled inside the timed region)', '',
       '| run | value rays/s | ms/step | e2e rays/s | MLP fwd / bwd / other ms | frame ms | clocks |', '|---|---|---|---|---|---|---|']
for name, d in lines:
    m, r = d['roofline']['ms_per_step'], d.get('render') or {}
    txt.append(f"| {name} | {d['value']:.0f} | {d['ms_per_step']:.2f} | {d['e2e']['value']:.0f} | {m['mlp_forward']:.2f} / {m['mlp_backward']:.2f} / {m['other']:.2f} | "
               f"{r.get('ms_per_frame', float('nan')):.1f} | {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']} |")
txt += ['']
for name, d in lines[1:]:
    txt.append(f"Weak scaling at {d['n_gpus']} GPUs (4096 rays per GPU): {d['value'] / n1['value'] / d['n_gpus']:.3f} of linear (from the values above; the driver computes its own figure).")
if ref is not None:
    txt.append(f"Reference arm (`bench.py --impl reference --steps 3 --warmup 1`, {ref['cpu_baseline']['sample']}, {ref['cpu_baseline']['cores']} host cores): {ref['value']:.0f} rays/s.")
txt += [f"`cpu_baseline` of the N=1 line ({n1['cpu_baseline']['sample']}): {n1['cpu_baseline']['value']:.0f} rays/s on {n1['cpu_baseline']['cores']} cores.", '',
        'Roofline object of the N=1 line:', '', '```json', json.dumps(n1['roofline'], indent=1), '```', '',
        'Notes.',
        '* The training step runs at the board power cap (`sw_power_cap`, SM clock 1700-1800 of 1965 MHz under sustained load).  The fused loss kernels of this round took the',
        '  non-MLP part of the step from 0.59 to 0.32 ms; the MLP kernels, no longer separated by ~80 eager launches, then ran at a slightly lower clock and the step time did not',
        '  move: it is set by energy per step -- 31.4 GB of HBM traffic and 5.3 TFLOP of tensor work -- not by launch gaps.',
        f"* `e2e`: the batch is packed into one pinned buffer and copied host -> device every step on a copy stream (HostBatchStager, double buffered: the copy of step i+1",
        f"  travels under step i), the loss is read back every step; the host needs {n1['host_enqueue_ms_per_step']:.1f} ms to enqueue a step ({n1['gpu_launches'] // n1['steps']} launches through the C ABI).",
        '* Experiment not kept: capping NCCL\'s CTAs (`NCCL_MAX_CTAS=4` / `2`) so that the overlapped gradient all-reduce takes fewer SMs from the persistent 148-CTA backward',
        '  kernels: 8 GPUs 7.78 / 8.12 ms per step against 7.56 ms with NCCL\'s defaults (the exchange then outlasts the backward pass) -- defaults stay.']
open(os.path.join(P, f'{out_tag}_bench_lines.md'), 'w').write('\n'.join(txt) + '\n')
shutil.copy(os.path.join(G, f'{ncu_tag}_scan.md'), os.path.join(P, f'{out_tag}_c5_scan_microbench.md'))
shutil.copy(os.path.join(G, f'{ncu_tag}_launches.csv'), os.path.join(P, f'{out_tag}_launches.csv'))
print('\n'.join(out[5:20]))
print('\n'.join(txt[:12]))
