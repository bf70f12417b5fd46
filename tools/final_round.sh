python bench.py > gpurun_out/final2_n1.json 2> gpurun_out/final2_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final2_ref.json 2> gpurun_out/final2_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-render"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1h_launches.csv $B > gpurun_out/r1h_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:sample_|composite_|ray_losses|reproj_|gather_rows" -f -o gpurun_out/r1h_aux python tools/aux_kernels_prof.py > gpurun_out/r1h_ncu_aux.log 2>&1
ncu -i gpurun_out/r1h_aux.ncu-rep --page raw --csv > gpurun_out/r1h_aux_raw.csv 2>/dev/null
rm -f gpurun_out/r1h_aux.ncu-rep
python tools/scan_microbench.py > gpurun_out/r1h_scan.md 2>&1
tail -2 gpurun_out/final2_n1.json | cut -c1-600
cat gpurun_out/final2_ref.json | cut -c1-400
