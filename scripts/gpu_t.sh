timeout 900 python -m pytest tests -m gpu -q -x -k "tensor_core or tensor_cores or rollout or smoke or pinnsf" 2>&1 | tail -3
timeout 600 python scripts/bench_stages.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(round(v.get('ms', v.get('ms_per_step')),4)) for k,v in d.items() if 'tcgen05' in k or 'nn_rollout' in k})"
