set -x
mkdir -p gpurun_out
rm -f gpurun_out/r4d_quick.txt
for lib in _d96 _d64 ""; do
echo "== lib='$lib'" >> gpurun_out/r4d_quick.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 200 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error|error' >> gpurun_out/r4d_quick.txt
done
cat gpurun_out/r4d_quick.txt
