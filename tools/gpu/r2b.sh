set -x
mkdir -p gpurun_out
for v in "" _nonarrow _t384; do
  echo "== variant '$v'" >> gpurun_out/r2b_quick.txt
  ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$v.so python tools/quick_gpu_check.py 1024 >> gpurun_out/r2b_quick.txt 2>&1
done
cat gpurun_out/r2b_quick.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encode_staged -c 1 -f -o gpurun_out/r2b_enc python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2b_ncu_enc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_staged -c 1 -f -o gpurun_out/r2b_dec python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2b_ncu_dec.log 2>&1
