# MLAPM kernel iteration: parity tests of the MLAPM path + headline bench without the CPU leg.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "mlapm or MLAPM" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step', d['ms_per_step'], 'kernel_ms', r['kernel_ms'], 'frac', r['frac'], 'peak', r['peak'], 'e2e ms', d['e2e']['ms_per_step'], 'clk', d['clocks'])"
