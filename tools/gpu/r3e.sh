set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py "1D" > gpurun_out/r3e_1d.txt 2>&1
cat gpurun_out/r3e_1d.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r3e_pytest.txt 2>&1
tail -5 gpurun_out/r3e_pytest.txt
