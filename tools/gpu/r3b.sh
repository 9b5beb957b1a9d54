set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py "2D" > gpurun_out/r3b_2d.txt 2>&1
cat gpurun_out/r3b_2d.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200_d3.so timeout 200 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error|error' > gpurun_out/r3b_quick_d3.txt
cat gpurun_out/r3b_quick_d3.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r3b_pytest.txt 2>&1
tail -5 gpurun_out/r3b_pytest.txt
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200_d3.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r3b_pytest_d3.txt 2>&1
tail -5 gpurun_out/r3b_pytest_d3.txt
