mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "row_range" 2>&1 | tail -3
timeout 900 python bench.py --agents 1000000 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_1m.log 2>&1; tail -1 gpurun_out/bench_1m.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('1M agents: ms/step', d['ms_per_step'], 'value', d['value'], 'frac', r['frac'], 'e2e', d['e2e']['value'], 'nn', d['nn_path'])"
nvidia-smi --query-gpu=memory.used --format=csv
