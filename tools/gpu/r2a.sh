set -x
mkdir -p gpurun_out
nvidia-smi -L
python tools/quick_gpu_check.py 1024 > gpurun_out/r2a_quick.txt 2>&1
tail -8 gpurun_out/r2a_quick.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.txt 2>&1
tail -5 gpurun_out/r2a_pytest.txt
timeout 600 python tools/bench_ref_cuda.py > gpurun_out/r2a_refcuda.jsonl 2>&1
tail -12 gpurun_out/r2a_refcuda.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_staged -c 1 -f -o gpurun_out/r2a_enc python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2a_ncu_enc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_staged -c 1 -f -o gpurun_out/r2a_dec python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2a_ncu_dec.log 2>&1
ls -la gpurun_out | tail -5
