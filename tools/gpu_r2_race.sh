mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python tools/sanitize_small.py > gpurun_out/r2_racecheck_full.txt 2>&1
grep -E "RACECHECK SUMMARY|sanitize run ok" gpurun_out/r2_racecheck_full.txt
grep -E "Race reported|Warning|hazard" gpurun_out/r2_racecheck_full.txt | sed 's/\[.*//' | sort | uniq -c | sort -rn | head -20
B200DPF_MAC_TMA=0 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 40 python tools/sanitize_small.py 2>&1 | grep -E "RACECHECK SUMMARY|sanitize run ok|Race reported" | sed 's/\[.*//' | sort | uniq -c | head
