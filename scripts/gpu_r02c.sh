mkdir -p gpurun_out
PIML_TC_F16=0 timeout 300 python scripts/probe_tc_fwd.py 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x -k "tensor or rollout or nn or sharded or scene or dropin or smoke" 2>&1 | tail -3
timeout 300 python scripts/bench_stages.py 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())
for k,v in d.items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})"
