timeout 900 python -m pytest tests -m gpu -q -x -k "sfm or rollout or integrate" -s 2>&1 | tail -6
timeout 900 python scripts/bench_stages.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:v for k,v in d.items() if 'socialforce' in k or 'train' in k})"
