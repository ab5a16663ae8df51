mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "mlapm_symmetric" 2>&1 | tail -2
for cfg in 0 1 2 3; do for per in 2 4; do
  echo "cfg=$cfg per=$per"; PIML_MLAPM_SYM_CFG=$cfg PIML_MLAPM_SYM_PER=$per timeout 300 python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step', d['ms_per_step'], 'kernel_ms', r['kernel_ms'], 'frac', r['frac'])"
done; done
