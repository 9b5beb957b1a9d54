set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r4e_bench_n1.json 2> gpurun_out/r4e_bench_n1.err
tail -c 300 gpurun_out/r4e_bench_n1.json; tail -3 gpurun_out/r4e_bench_n1.err
python __graft_entry__.py smoke > gpurun_out/r4e_smoke.txt 2>&1; tail -1 gpurun_out/r4e_smoke.txt
