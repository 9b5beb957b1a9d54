set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py "1D" > gpurun_out/r3d_1d.txt 2>&1
cat gpurun_out/r3d_1d.txt
rm -f gpurun_out/r3c.txt
for lib in "" _rc8 _rc9; do
echo "== lib='$lib'" >> gpurun_out/r3c.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/bench_reversible.py >> gpurun_out/r3c.txt 2>&1
done
cat gpurun_out/r3c.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r3d_pytest.txt 2>&1
tail -5 gpurun_out/r3d_pytest.txt
