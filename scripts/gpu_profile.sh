mkdir -p gpurun_out
python scripts/profile_workloads.py --reps 1 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_workloads.csv python scripts/profile_workloads.py --reps 2 > gpurun_out/wl_under_ncu.log 2>&1
for k in mlapm_pairs2_kernel pinnsf_tile_kernel pinnsf_bwd_tile_kernel pinnsf_dw_kernel features_cells_kernel; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python scripts/profile_workloads.py --reps 2 > gpurun_out/ncu_$k.log 2>&1
  ls -la gpurun_out/prof_$k.ncu-rep
done
