set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.txt 2>&1
tail -15 gpurun_out/r2f_pytest.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 6000 gpurun_out/r2f_bench_n1.json; tail -5 gpurun_out/r2f_bench_n1.err
echo "== Q4 encode" > gpurun_out/r2f_q4.txt
ZFP_B200_Q4=1 timeout 300 python tools/quick_gpu_check.py 1024 >> gpurun_out/r2f_q4.txt 2>&1
tail -8 gpurun_out/r2f_q4.txt
