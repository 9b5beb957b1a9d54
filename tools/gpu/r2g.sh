set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py::test_random_access_box_decode -m gpu -x -q > gpurun_out/r2g_pytest.txt 2>&1
tail -25 gpurun_out/r2g_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -c 3000 gpurun_out/r2g_bench_n2.json; tail -5 gpurun_out/r2g_bench_n2.err
