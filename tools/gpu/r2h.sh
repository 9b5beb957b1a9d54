set -x
mkdir -p gpurun_out
timeout 300 python tools/quick_gpu_check.py 512 > gpurun_out/r2h_quick.txt 2>&1
tail -9 gpurun_out/r2h_quick.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.txt 2>&1
tail -15 gpurun_out/r2h_pytest.txt
timeout 600 python tools/bench_variable.py > gpurun_out/r2h_var.txt 2>&1
tail -12 gpurun_out/r2h_var.txt
ZFP_B200_NO_VAR1=1 timeout 600 python tools/bench_variable.py > gpurun_out/r2h_var_old.txt 2>&1
tail -12 gpurun_out/r2h_var_old.txt
