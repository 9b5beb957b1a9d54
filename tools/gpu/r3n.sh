set -x
mkdir -p gpurun_out
timeout 900 python tools/gpu/index_rebuild_time.py > gpurun_out/r3n_index.txt 2>&1
cat gpurun_out/r3n_index.txt
