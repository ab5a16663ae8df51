# Round check on one B200: GPU parity tests, smoke, headline bench (+ reference arm), stage timings, launch list + ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-700
timeout 900 python scripts/bench_stages.py > gpurun_out/stages.log 2>&1; tail -3 gpurun_out/stages.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r01d_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlapm_sym_kernel -s 1 -c 1 -f -o gpurun_out/prof_mlapm_sym_kernel python scripts/profile_workloads.py --reps 2 > gpurun_out/ncu_sym.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01d_launches_bench.csv
