mkdir -p gpurun_out
python scripts/prof_sfm_rollout.py --scenes 1; python scripts/prof_sfm_rollout.py --scenes 64; python scripts/prof_sfm_rollout.py --scenes 4096 --frames 100
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sfm_rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_sfm_rollout_kernel python scripts/prof_sfm_rollout.py --scenes 64 > gpurun_out/ncu_sfmroll.log 2>&1; ls -la gpurun_out/prof_sfm_rollout_kernel.ncu-rep
