# launch list of the NN workload + ncu --set full of the 16-bit tensor-core forward and the cell-list kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02e_launches_bench_nn.csv python bench.py --workload nn --steps 3 --warmup 3 --no-cpu > gpurun_out/r02e_nn_under_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02e_launches_bench_nn.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows[-90:]:
    k=r[4][:60]; agg.setdefault(k,[]).append(float(r[-1]))
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} avg={sum(v)/len(v)/1000:8.1f} us")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pinnsf_tc16_kernel -s 3 -c 1 -f -o gpurun_out/prof_pinnsf_tc16_kernel python bench.py --workload nn --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_tc16.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
