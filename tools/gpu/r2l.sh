set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_staged -c 1 -f -o gpurun_out/r2l_dec python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2l_ncu_dec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_staged -c 1 -f -o gpurun_out/r2l_enc python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2l_ncu_enc.log 2>&1
ZFP_B200_VAR1=1 timeout 600 python tools/gpu/var1_diag.py > gpurun_out/r2l_var1.txt 2>&1
cat gpurun_out/r2l_var1.txt
timeout 600 python tools/gpu/var1_diag.py > gpurun_out/r2l_var0.txt 2>&1
cat gpurun_out/r2l_var0.txt
