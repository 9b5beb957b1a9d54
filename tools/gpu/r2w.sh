set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err
tail -c 600 gpurun_out/r2w_bench_n1.json; tail -5 gpurun_out/r2w_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err
tail -c 600 gpurun_out/r2w_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2w_ncu_bench.log 2>&1
tail -3 gpurun_out/r2w_ncu_bench.log
python __graft_entry__.py smoke > gpurun_out/r2w_smoke.txt 2>&1; tail -2 gpurun_out/r2w_smoke.txt
