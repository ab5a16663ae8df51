# compute-sanitizer over small invocations of every kernel family (one B200).  Logs -> gpurun_out/sanitizer_<tool>.log
mkdir -p gpurun_out
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 200 --error-exitcode 3 \
      --log-file gpurun_out/sanitizer_$tool.log python scripts/sanitize_workload.py > gpurun_out/sanitizer_$tool.out 2>&1
  echo "$tool rc=$? $(grep -c 'part .*: ok' gpurun_out/sanitizer_$tool.out) parts ok; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_$tool.log | tail -1)"
done
