mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02j_bench.log 2>&1; tail -1 gpurun_out/r02j_bench.log | python -c "
import sys,json; d=json.loads(sys.stdin.read())
print('mlapm', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], 'parity', d['parity']['pass'])
print('nn', d['nn_path']['ms_per_step'], d['nn_path']['stage_ms'], d['nn_path']['gpu_launches'], d['nn_path']['parity']['pass'], d['nn_path']['e2e']['ms_per_step'], d['nn_path']['cpu_baseline']['value'])
print('training', d['training'])"
