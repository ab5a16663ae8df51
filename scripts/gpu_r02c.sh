mkdir -p gpurun_out
for d in 0 1; do PIML_TC_DEBUG=$d PIML_TC_PROF=1 timeout 300 python scripts/tc_time.py 2>&1 | grep -E "tc16 prof|tcgen05" | tail -2; done
