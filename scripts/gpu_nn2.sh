mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/check_nn_sharded.py > gpurun_out/nn2.log 2>&1
grep "sharded\|Error" gpurun_out/nn2.log | head
