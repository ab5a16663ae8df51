mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for g in 0 1; do PIML_ROLLOUT_GRAPH=$g timeout 600 python scripts/bench_stages.py 2>/dev/null > gpurun_out/r02h_stages_graph$g.json; python - <<PY
import json
d=json.load(open('gpurun_out/r02h_stages_graph$g.json'))
print('graph=$g', {k:round(v.get('ms_per_step',v.get('ms',0)),4) for k,v in d.items() if k.startswith('rollout') or k.startswith('nn_rollout')})
PY
done
timeout 900 python bench.py > gpurun_out/r02h_bench.log 2>&1; tail -1 gpurun_out/r02h_bench.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('mlapm', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity']['pass'], d['parity']['max_rel_force'])
print('nn', d['nn_path']['ms_per_step'], d['nn_path']['stage_ms'], d['nn_path']['parity'])
print('crowd_1m', d['crowd_1m']); print('scenes', d['scenes_4096'])"
SAN_TOOLS="memcheck" SAN_TIMEOUT=420 bash scripts/gpu_sanitize.sh
