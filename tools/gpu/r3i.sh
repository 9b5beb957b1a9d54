set -x
mkdir -p gpurun_out
rm -f gpurun_out/r3i.txt
for lib in "" _s128; do
echo "== lib='$lib'" >> gpurun_out/r3i.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_configs.py "2D" >> gpurun_out/r3i.txt 2>&1
done
cat gpurun_out/r3i.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r3i_pytest.txt 2>&1
tail -3 gpurun_out/r3i_pytest.txt
