set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.txt 2>&1
tail -15 gpurun_out/r2j_pytest.txt
timeout 300 python tools/bench_configs.py "C3 4D" > gpurun_out/r2j_4d.txt 2>&1
ZFP_B200_4D_OLD=1 timeout 300 python tools/bench_configs.py "C3 4D" >> gpurun_out/r2j_4d.txt 2>&1
cat gpurun_out/r2j_4d.txt
