mkdir -p gpurun_out
for exp in 2,2,1 2,2,2 2,2,3; do
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
