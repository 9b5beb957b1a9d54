set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3l.txt
for lib in "" _f32a _f32b; do
echo "== lib='$lib'" >> gpurun_out/r3l.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "3D f32" >> gpurun_out/r3l.txt 2>&1
done
cat gpurun_out/r3l.txt
