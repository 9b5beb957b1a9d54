set -x
mkdir -p gpurun_out
timeout 200 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error|error' > gpurun_out/r3z_quick.txt
cat gpurun_out/r3z_quick.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3z_pytest.txt 2>&1
tail -4 gpurun_out/r3z_pytest.txt
python __graft_entry__.py smoke > gpurun_out/r3z_smoke.txt 2>&1; tail -1 gpurun_out/r3z_smoke.txt
