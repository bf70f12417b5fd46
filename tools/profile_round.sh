#!/bin/bash
# One profiling pass on the GPU box (run under gpurun): plain bench, ncu launch list, ncu --set full of one training
# step's MLP kernels and of the scan / loss / gather kernels, C5 scan microbench.  Outputs land in gpurun_out/ with the tag.
TAG=${1:-r1}
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-render"
$B > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:tc_(forward|dgrad|wgrad)" -s 36 -c 12 -f -o gpurun_out/${TAG}_full $B > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_full.ncu-rep     # 40 MB each: gpurun brings back at most 64 MiB
python tools/aux_kernels_prof.py > gpurun_out/${TAG}_aux_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "regex:sample_|composite_|ray_losses|reproj_|gather_rows" -f -o gpurun_out/${TAG}_aux python tools/aux_kernels_prof.py > gpurun_out/${TAG}_ncu_aux.log 2>&1
ncu -i gpurun_out/${TAG}_aux.ncu-rep --page raw --csv > gpurun_out/${TAG}_aux_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_aux.ncu-rep
python tools/scan_microbench.py > gpurun_out/${TAG}_scan.md 2>&1
tail -3 gpurun_out/${TAG}_scan.md
