set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2v_quick.txt
for v in "PS:_ps2x8" ":"; do
ps=${v%%:*}; lib=${v##*:}
echo "== ps='$ps' lib='$lib'" >> gpurun_out/r2v_quick.txt
if [ -n "$ps" ]; then export ZFP_B200_PS=1; else unset ZFP_B200_PS; fi
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 100 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error|error' >> gpurun_out/r2v_quick.txt
done
unset ZFP_B200_PS
cat gpurun_out/r2v_quick.txt
