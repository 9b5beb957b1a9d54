set -x
mkdir -p gpurun_out
timeout 200 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error|error' > gpurun_out/r4c_quick.txt
cat gpurun_out/r4c_quick.txt
timeout 600 python tools/bench_variable.py > gpurun_out/r4c_var.txt 2>&1
cat gpurun_out/r4c_var.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r4c_pytest.txt 2>&1
tail -4 gpurun_out/r4c_pytest.txt
