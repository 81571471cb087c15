mkdir -p gpurun_out
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool (final code)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|sanitize run ok|Error" | head -6
done 2>&1 | tee gpurun_out/r2_sanitizer_final.txt
