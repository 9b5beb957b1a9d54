set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_variable.py > gpurun_out/r3u_var.txt 2>&1
cat gpurun_out/r3u_var.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3u_pytest.txt 2>&1
tail -4 gpurun_out/r3u_pytest.txt
