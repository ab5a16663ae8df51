mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "tensor_core or tensor_cores or rollout" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu --steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('nn_path', d['nn_path'])"
timeout 600 python scripts/bench_stages.py 2>&1 | tail -45
