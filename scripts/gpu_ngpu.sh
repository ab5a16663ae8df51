# G x B200 (G = $1): fused exchange checks and the agent-sharded bench
G=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 scripts/check_push_exchange.py 2>&1 | grep -v Warning | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $G --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_${G}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${G}gpu.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'], d['config']['exchange'], d['config']['exchange_note'])"
