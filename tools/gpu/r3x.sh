set -x
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r3x_bench_n4.json 2> gpurun_out/r3x_bench_n4.err
tail -c 400 gpurun_out/r3x_bench_n4.json; tail -2 gpurun_out/r3x_bench_n4.err
