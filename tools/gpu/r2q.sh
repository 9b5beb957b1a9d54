set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_staged -c 1 -f -o gpurun_out/r2q_enc python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2q_ncu_enc.log 2>&1
tail -2 gpurun_out/r2q_ncu_enc.log
