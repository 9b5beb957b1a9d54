set -x
mkdir -p gpurun_out
ZFP_B200_LIB=zfp_b200/lib/libzfp_b200_e384.so timeout 400 ncu --set full --clock-control none -k regex:encode_staged -c 1 -f -o /tmp/r3g_enc384 python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r3g_ncu.log 2>&1
ncu -i /tmp/r3g_enc384.ncu-rep --page raw --csv > gpurun_out/r3g_enc384.csv 2>/dev/null
tail -2 gpurun_out/r3g_ncu.log
