set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_stages.py > gpurun_out/stages.log 2>&1; echo "rc=$?" >> gpurun_out/stages.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlapm_pairs -s 3 -c 1 -o gpurun_out/prof_mlapm python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_mlapm.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
echo done
