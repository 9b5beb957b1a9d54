set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2u.txt
for lib in "" _t128 _t256; do
echo "== lib='$lib'" >> gpurun_out/r2u.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "f32" >> gpurun_out/r2u.txt 2>&1
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "i32" >> gpurun_out/r2u.txt 2>&1
done
cat gpurun_out/r2u.txt
