timeout 900 python -m pytest tests -m gpu -q -x -k "feature or cell or rollout or row_range or smoke" 2>&1 | tail -3
for g in 1 4 8; do echo "lanes=$g"; PIML_CELLS_LANES=$g timeout 300 python scripts/bench_stages.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:(round(v.get('ms', v.get('ms_per_step')),4)) for k,v in d.items() if 'features' in k or 'nn_rollout' in k or 'rollout_S64' in k})"; done
