set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "mlapm" > gpurun_out/pytest_mlapm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mlapm.log
tail -3 gpurun_out/pytest_mlapm.log
for exp in 2,2,0 2,1,0 2,4,0 1,2,0 1,4,0 4,1,0 4,2,0 2,2,1; do
  PIML_MLAPM_EXP=$exp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_exp_$exp.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_exp_$exp.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("exp $exp", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"]["sm_mhz"])
else:
    print("exp $exp FAILED", open("gpurun_out/bench_exp_$exp.log").read()[-500:])
PY
done
for sp in 4 9 18 27 36; do
  PIML_MLAPM_SPLIT=$sp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_split_$sp.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_split_$sp.log") if x.startswith("{")]
d=json.loads(l[-1]); print("split $sp", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
PY
done
