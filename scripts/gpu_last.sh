mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sfm_rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_sfm_rollout_kernel python scripts/prof_sfm_rollout.py --scenes 64 > gpurun_out/ncu_sfmroll.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pinnsf_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_pinnsf_tc_kernel python scripts/tc_prof.py --no-prof > gpurun_out/ncu_tc.log 2>&1
ls -la gpurun_out/*.ncu-rep
