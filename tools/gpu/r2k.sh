set -x
mkdir -p gpurun_out
./tools/ubench/fp64_lift > gpurun_out/r2k_ubench.txt 2>&1
cat gpurun_out/r2k_ubench.txt
timeout 300 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error' > gpurun_out/r2k_quick.txt
cat gpurun_out/r2k_quick.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.txt 2>&1
tail -15 gpurun_out/r2k_pytest.txt
timeout 300 python tools/bench_configs.py "C3 4D" > gpurun_out/r2k_4d.txt 2>&1
ZFP_B200_4D_OLD=1 timeout 300 python tools/bench_configs.py "C3 4D" >> gpurun_out/r2k_4d.txt 2>&1
cat gpurun_out/r2k_4d.txt
