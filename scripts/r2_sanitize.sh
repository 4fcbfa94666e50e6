mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_forward.py 256 > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "rc=$?"; grep -E "forward ok|ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/r2_sanitizer_$tool.log | head -8
done
