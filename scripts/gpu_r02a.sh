# round 2, first GPU session (one B200): new parity / drop-in tests first, then the whole suite, bench, sanitizers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity_sizes.py tests/test_gpu_dropin.py -m gpu -q -s > gpurun_out/r02a_new_tests.log 2>&1
echo "new tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r02a_new_tests.log | tail -3
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity_sizes.py --deselect tests/test_gpu_dropin.py > gpurun_out/r02a_pytest_gpu.log 2>&1
echo "suite rc=$?"; tail -3 gpurun_out/r02a_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02a_bench.log 2>&1; tail -1 gpurun_out/r02a_bench.log | cut -c1-600
SAN_TOOLS="memcheck synccheck" SAN_TIMEOUT=420 bash scripts/gpu_sanitize.sh
timeout 600 python scripts/drift_report.py > gpurun_out/r02a_drift.log 2>&1; tail -5 gpurun_out/r02a_drift.log | cut -c1-300
