mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01c_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches_workloads.csv python scripts/profile_workloads.py --reps 2 > gpurun_out/wl_under_ncu.log 2>&1
for k in mlapm_sym_kernel pinnsf_tc_kernel features_cells_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python scripts/profile_workloads.py --reps 2 > gpurun_out/ncu_$k.log 2>&1
  ls -la gpurun_out/prof_$k.ncu-rep
done
