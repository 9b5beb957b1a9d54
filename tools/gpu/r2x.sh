set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2x_pytest.txt 2>&1
tail -8 gpurun_out/r2x_pytest.txt
