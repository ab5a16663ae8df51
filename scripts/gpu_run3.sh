set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python scripts/probe_pipes.py > gpurun_out/probes.json 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 python scripts/bench_stages.py > gpurun_out/stages.log 2>&1; echo "rc=$?" >> gpurun_out/stages.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlapm_pairs -s 3 -c 1 -o gpurun_out/prof_mlapm_v1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_mlapm.log 2>&1
tail -5 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench.log gpurun_out/probes.json
