set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3j_bench_n2.json 2> gpurun_out/r3j_bench_n2.err
tail -c 1500 gpurun_out/r3j_bench_n2.json; tail -3 gpurun_out/r3j_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r3j_pytest_multi.txt 2>&1
tail -3 gpurun_out/r3j_pytest_multi.txt
