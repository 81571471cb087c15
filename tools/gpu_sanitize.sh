mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok|Error|hazard" | head -8
done 2>&1 | tee gpurun_out/sanitizer.txt
