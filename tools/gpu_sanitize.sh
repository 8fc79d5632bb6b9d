mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|leaf2|flow|levels|bush" gpurun_out/sanitize_$tool.log | head -12
done
