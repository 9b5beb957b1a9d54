set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3o_pytest.txt 2>&1
tail -5 gpurun_out/r3o_pytest.txt
timeout 900 python bench.py > gpurun_out/r3o_bench_n1.json 2> gpurun_out/r3o_bench_n1.err
tail -c 300 gpurun_out/r3o_bench_n1.json; tail -3 gpurun_out/r3o_bench_n1.err
python __graft_entry__.py smoke > gpurun_out/r3o_smoke.txt 2>&1; tail -1 gpurun_out/r3o_smoke.txt
