#!/bin/bash
# Final measurement pass on the GPU box (under gpurun): default bench, reference arm, ncu launch list, ncu --set full of the
# scan / loss / gather kernels, C5 microbench.  NT / BT = tags of the ncu and bench files in gpurun_out/ (tools/summarize_round.py).
python bench.py > gpurun_out/${BT:-final3}_n1.json 2> gpurun_out/${BT:-final3}_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${BT:-final3}_ref.json 2> gpurun_out/${BT:-final3}_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-render"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${NT:-r1i}_launches.csv $B > gpurun_out/${NT:-r1i}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:sample_|composite_|ray_losses|reproj_|gather_rows" -f -o gpurun_out/${NT:-r1i}_aux python tools/aux_kernels_prof.py > gpurun_out/${NT:-r1i}_ncu_aux.log 2>&1
ncu -i gpurun_out/${NT:-r1i}_aux.ncu-rep --page raw --csv > gpurun_out/${NT:-r1i}_aux_raw.csv 2>/dev/null
rm -f gpurun_out/${NT:-r1i}_aux.ncu-rep
python tools/scan_microbench.py > gpurun_out/${NT:-r1i}_scan.md 2>&1
tail -2 gpurun_out/${BT:-final3}_n1.json | cut -c1-600
cat gpurun_out/${BT:-final3}_ref.json | cut -c1-400
