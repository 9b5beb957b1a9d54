set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3c.txt
for lib in "" _rc8 _rc9; do
echo "== lib='$lib'" >> gpurun_out/r3c.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "i32" >> gpurun_out/r3c.txt 2>&1
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_reversible.py >> gpurun_out/r3c.txt 2>&1
done
cat gpurun_out/r3c.txt
