set -x
mkdir -p gpurun_out
ZFP_B200_PS=1 ZFP_B200_LIB=zfp_b200/lib/libzfp_b200_ps5.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_ps -c 1 -f -o gpurun_out/r2o_dec_ps5 python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2o_ncu.log 2>&1
tail -3 gpurun_out/r2o_ncu.log
