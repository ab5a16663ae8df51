# symmetric MLAPM kernel: tests, bench, per-split tuning, launch list and one ncu --set full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "mlapm or MLAPM" 2>&1 | tail -3
for per in 2 3 4 6 8; do
  echo "per=$per"; PIML_MLAPM_SYM_PER=$per timeout 300 python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step', d['ms_per_step'], 'kernel_ms', r['kernel_ms'], 'frac', r['frac'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlapm_sym_kernel -s 1 -c 1 -f -o gpurun_out/prof_mlapm_sym_kernel python scripts/profile_workloads.py --reps 2 > gpurun_out/ncu_sym.log 2>&1
ls -la gpurun_out/
