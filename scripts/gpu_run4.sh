set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "mlapm" > gpurun_out/pytest_mlapm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mlapm.log
for rp in 1 2 4; do
  PIML_MLAPM_RP=$rp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_rp$rp.log 2>&1
done
PIML_MLAPM_RP=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlapm_pairs2 -s 3 -c 1 -o gpurun_out/prof_mlapm_v2_rp2 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_mlapm_v2.log 2>&1
tail -3 gpurun_out/pytest_mlapm.log
for rp in 1 2 4; do python - <<PY
import json
l=[x for x in open("gpurun_out/bench_rp$rp.log") if x.startswith("{")]
d=json.loads(l[-1]); print("rp$rp", d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["clocks"])
PY
done
