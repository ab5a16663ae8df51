mkdir -p gpurun_out
timeout 300 python scripts/probe_tc_fwd.py 2>&1 | tail -2
PIML_TC_F16=0 timeout 300 python scripts/probe_tc_fwd.py 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --workload nn --steps 10 > gpurun_out/r02d_bench_nn.log 2>&1; tail -1 gpurun_out/r02d_bench_nn.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['parity'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['gpu_launches'])"
