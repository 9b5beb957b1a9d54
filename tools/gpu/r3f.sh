set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -c 1 -f -o gpurun_out/r3f_dec1d python tools/prof_target.py 268435456x f64 8 1 > gpurun_out/r3f_ncu.log 2>&1
tail -2 gpurun_out/r3f_ncu.log
