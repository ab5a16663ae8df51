mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-600
timeout 900 python scripts/bench_stages.py > gpurun_out/stages.log 2>&1; tail -60 gpurun_out/stages.log
timeout 600 python scripts/drift_report.py > gpurun_out/drift.log 2>&1; tail -12 gpurun_out/drift.log
