set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3w_sanitizer.txt
for tool in memcheck racecheck; do
  echo "--tool $tool (tools/sanitize_target.py):" >> gpurun_out/r3w_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_target.py 2>&1 | grep -E "calls|SUMMARY|ERROR|hazard|Race|Invalid|at 0x|by thread" | head -30 >> gpurun_out/r3w_sanitizer.txt
done
echo "--tool memcheck (tools/sanitize_target_index.py):" >> gpurun_out/r3w_sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_target_index.py 2>&1 | grep -E "cases|launches|SUMMARY|ERROR|Invalid|at 0x|by thread|Assertion" | head -30 >> gpurun_out/r3w_sanitizer.txt
cat gpurun_out/r3w_sanitizer.txt
