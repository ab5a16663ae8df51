# Round check on one B200: GPU parity tests, smoke, headline bench, per-stage timings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 600 python scripts/bench_stages.py > gpurun_out/stages.log 2>&1; tail -40 gpurun_out/stages.log
