mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_sizes.py tests/test_gpu_parity.py -m gpu -q -s -k "mlapm" > gpurun_out/r02f_mlapm_tests.log 2>&1; grep -E "^MLAPM|passed|failed" gpurun_out/r02f_mlapm_tests.log | tail -14
timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'])"
timeout 600 python bench.py --no-cpu --agents 1000000 --steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench 1M', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
SAN_TOOLS="racecheck initcheck" SAN_TIMEOUT=600 bash scripts/gpu_sanitize.sh
