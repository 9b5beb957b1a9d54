set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3v.txt
for lib in "" _q6; do
echo "== lib='$lib'" >> gpurun_out/r3v.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "4D" >> gpurun_out/r3v.txt 2>&1
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "4D" >> gpurun_out/r3v.txt 2>&1
done
cat gpurun_out/r3v.txt
