# G x B200 (G = $1): fused exchange checks (incl. oracle rows at N = 100k), NN sharded check, the agent-sharded bench
G=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 scripts/check_push_exchange.py 2>&1 | grep -v Warning | tail -$((G+3))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29523 scripts/check_nn_sharded.py 2>&1 | grep -v Warning | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $G --steps 20 --warmup 3 2>gpurun_out/bench_${G}gpu.err | tail -1 > gpurun_out/r02_bench_${G}gpu.json
tail -3 gpurun_out/bench_${G}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${G}gpu.json'))
print('mlapm', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['parallelism']['exchange'], d['parallelism']['exchange_note'])
print('nn', d.get('nn_path'))
print('timeline', d.get('timeline'))
print('crowd_1m', d.get('crowd_1m'))
print('scenes', d.get('scenes_4096'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus $G --steps 20 --warmup 3 --workload nn 2>/dev/null | tail -1 | cut -c1-400
