"""profiles/<out>_aux_kernels_ncu.md from gpurun_out/<tag>_aux_raw.csv (ncu --set full of tools/aux_kernels_prof.py).
    python tools/summarize_aux_ncu.py <tag of the ncu files> <output tag>"""
import csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
ncu_tag, out_tag = sys.argv[1:3]


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
out = [f'# {out_tag}: samplers, compositing, losses and batch gather under `ncu --set full --clock-control none`', '',
       'Command: `ncu --set full --clock-control none --import-source on -k "regex:sample_|composite_|ray_losses|reproj_|gather_rows" python tools/aux_kernels_prof.py`',
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

print('\n'.join(out[:24]))
