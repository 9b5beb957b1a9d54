set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3y_quick.txt
for p in 0 3 6; do
echo "== persist $p" >> gpurun_out/r3y_quick.txt
ZFP_B200_PERSIST=$p timeout 200 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error|error' >> gpurun_out/r3y_quick.txt
done
cat gpurun_out/r3y_quick.txt
