set -x
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "--tool $tool:" >> gpurun_out/r2y_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E "calls|SUMMARY|ERROR|hazard|Race|Invalid|at 0x|by thread" | head -40 >> gpurun_out/r2y_sanitizer.txt
done
cat gpurun_out/r2y_sanitizer.txt
