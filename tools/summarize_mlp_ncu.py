"""profiles/<out>_ncu_full_summary.md and profiles/mlp_dram_traffic.json from the raw csv of one `ncu --set full` capture of the
MLP kernels of one training step (tools/profile_round.sh).   python tools/summarize_mlp_ncu.py <gpurun_out csv> <out tag> [note]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, tag = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda d, k: float(d[ix[k]].replace(',', ''))                                         # noqa: E731
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
val = lambda d, k: num(d, k) * scale[units[ix[k]]]                                           # noqa: E731
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6548.2
out = [f'# {tag}: one C2 training step (4096 rays, 4 MLPs) under `ncu --set full --clock-control none`, MLP kernels', '',
       'Command: `ncu --set full --clock-control none --import-source on -k "regex:tc_(forward|dgrad|wgrad)" -s 36 -c 12 python bench.py --steps 2 --warmup 3 --no-cpu --no-render --no-c5 --no-trainer`',
       'Launch order inside a step: forward coarse, points-aug, views-aug, fine; then dgrad + wgrad per MLP (fine first).  ncu times are serialised and cold-cache: compare shares, not absolutes.', note, '',
       '| kernel | time ms | dram read GB | dram write GB | dram GB/s | frac of measured %.0f GB/s | tensor pipe (hmma) active %% | issue slots %% | regs |' % peak,
       '|---|---|---|---|---|---|---|---|---|']
agg = {}
for d in data:
    name = d[ix['Kernel Name']].split('(')[0].replace('snerf::', '').replace('void ', '')
    t, rd, wr = val(d, 'gpu__time_duration.sum'), val(d, 'dram__bytes_read.sum'), val(d, 'dram__bytes_write.sum')
    gbs = (rd + wr) / 1e9 / (t * 1e-3)
    out.append(f"| `{name}` | {t:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {gbs:.0f} | {gbs / peak:.2f} | "
               f"{num(d, 'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active') * 100 if False else num(d, 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active') if 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active' in ix else float('nan'):.1f} | "
               f"{num(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | {d[ix['launch__registers_per_thread']]} |")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += t; a[2] += rd + wr
out += ['', 'Per step (sum over the launches of each kernel):', '', '| kernel | launches | ms | dram GB | GB/s | frac of %.0f GB/s |' % peak, '|---|---|---|---|---|---|']
tot = [0, 0.0, 0.0]
for k, (n, t, b) in agg.items():
    out.append(f'| `{k}` | {n} | {t:.3f} | {b / 1e9:.2f} | {b / 1e9 / (t * 1e-3):.0f} | {b / 1e9 / (t * 1e-3) / peak:.2f} |')
    tot[0] += n; tot[1] += t; tot[2] += b
out.append(f'| all | {tot[0]} | {tot[1]:.3f} | {tot[2] / 1e9:.2f} | {tot[2] / 1e9 / (tot[1] * 1e-3):.0f} | {tot[2] / 1e9 / (tot[1] * 1e-3) / peak:.2f} |')
dst = os.path.join(ROOT, 'profiles', f'{tag}_ncu_full_summary.md')
open(dst, 'w').write('\n'.join(out) + '\n')
json.dump({'bytes_per_step': tot[2], 'source': f'profiles/{tag}_ncu_full_summary.md',
           'how': 'ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum summed over the 12 tc_forward / tc_dgrad / tc_wgrad launches of one 4096-ray C2 step',
           'per_kernel_gb': {k: round(v[2] / 1e9, 2) for k, v in agg.items()}},
          open(os.path.join(ROOT, 'profiles', 'mlp_dram_traffic.json'), 'w'), indent=1)
print(open(dst).read())
