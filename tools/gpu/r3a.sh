set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py "2D" > gpurun_out/r3a_2d.txt 2>&1
cat gpurun_out/r3a_2d.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r3a_pytest.txt 2>&1
tail -5 gpurun_out/r3a_pytest.txt
