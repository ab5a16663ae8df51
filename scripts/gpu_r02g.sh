mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bench_size" 2>&1 | tail -2
SAN_TOOLS="initcheck" SAN_TIMEOUT=600 bash scripts/gpu_sanitize.sh
grep -A12 "Uninitialized" gpurun_out/sanitizer_initcheck.log | grep -E " at piml|in .*\.cu:" | awk '{$1=$1};1' | sort | uniq -c | sort -rn | head -20
