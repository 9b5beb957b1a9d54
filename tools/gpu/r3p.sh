set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_variable.py > gpurun_out/r3p_var.txt 2>&1
ZFP_B200_NO_OVERLAP=1 timeout 600 python tools/bench_variable.py > gpurun_out/r3p_var_noov.txt 2>&1
cat gpurun_out/r3p_var.txt gpurun_out/r3p_var_noov.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3p_pytest.txt 2>&1
tail -4 gpurun_out/r3p_pytest.txt
