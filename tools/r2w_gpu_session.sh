set -x
for tool in synccheck racecheck initcheck memcheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/duo_sanitize.py > gpurun_out/r2w_$tool.log 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Barrier error|Uninitialized|Invalid|^ok" gpurun_out/r2w_$tool.log | sort | uniq -c | head -12
done
