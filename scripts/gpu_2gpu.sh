# 2 x B200: fused exchange checks (ordered + symmetric) and the agent-sharded bench
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_push_exchange.py 2>&1 | grep -v Warning | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e'], d['config']['parallelism'], d['config']['exchange_note'])"
PIML_FORCE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --exchange nccl 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('nccl ordered:', d['value'], d['ms_per_step'])"
