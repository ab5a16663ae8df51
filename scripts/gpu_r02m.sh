# Round 2, second half, on one B200: what the numbers and summaries named r02l / r02m under profiles/ were produced with.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as G; G.smoke()" 2>&1 | tail -1
timeout 600 python scripts/fuzz_parity.py --seconds 150 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/r02m_bench.log 2>&1; tail -c 300 gpurun_out/r02m_bench.log
timeout 300 python bench.py --workload nn --steps 20 --warmup 3 > gpurun_out/r02l_bench_nn.log 2>&1
timeout 300 python scripts/time_nn_step.py | tail -2
timeout 300 python scripts/time_scenes.py 4096 100 | cut -c1-400
# launch lists (cold-cache, serialised: shares, not absolutes) and the two --set full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches_bench_nn.csv \
    python bench.py --workload nn --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pinnsf_tc16 -s 8 -c 1 -f -o gpurun_out/prof_pinnsf_tc16_kernel \
    python scripts/time_nn_step.py 100000 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:features_sorted -s 12 -c 1 -f -o gpurun_out/prof_features_sorted \
    python scripts/time_nn_step.py 100000 3 > /dev/null 2>&1
# sanitizers over the fused step
for t in memcheck racecheck synccheck initcheck; do
    timeout 600 compute-sanitizer --tool $t --log-file gpurun_out/r02l_sanitizer_$t.log python scripts/sanitize_workload.py --part nnstep > /dev/null 2>&1
    tail -1 gpurun_out/r02l_sanitizer_$t.log
done
# here (no GPU): python scripts/ncu_summary.py <rep> profiles/<name>.txt "<note>"; python scripts/ncu_lines.py <rep> <cubin> <kernel>
