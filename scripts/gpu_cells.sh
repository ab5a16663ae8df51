mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -k "cell_list or features" > gpurun_out/pytest_cells.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_cells.log
tail -30 gpurun_out/pytest_cells.log
timeout 600 python scripts/bench_stages.py > gpurun_out/stages.log 2>&1; tail -45 gpurun_out/stages.log
