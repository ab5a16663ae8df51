# round 2, tc16 bring-up: correctness of the 16-bit tensor-core forward and timing against the tf32 kernel
mkdir -p gpurun_out
timeout 300 python scripts/probe_tc_fwd.py > gpurun_out/r02b_tc16_fwd.log 2>&1; tail -12 gpurun_out/r02b_tc16_fwd.log
PIML_TC_F16=0 timeout 300 python scripts/probe_tc_fwd.py 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor or rollout or nn or sharded or scene" 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --workload nn --steps 10 > gpurun_out/r02b_bench_nn.log 2>&1; tail -1 gpurun_out/r02b_bench_nn.log | cut -c1-1500
