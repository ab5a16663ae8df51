mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02i_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-config5 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlapm_sym_kernel -s 1 -c 1 -f -o gpurun_out/prof_mlapm_sym_kernel python scripts/profile_workloads.py --reps 2 > gpurun_out/ncu_sym.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pinnsf_tc16_kernel -s 3 -c 1 -f -o gpurun_out/prof_pinnsf_tc16_kernel python bench.py --workload nn --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_tc16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:features_cells_kernel -s 3 -c 1 -f -o gpurun_out/prof_features_cells_kernel python bench.py --workload nn --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_cells.log 2>&1
ls -la gpurun_out/*.ncu-rep
