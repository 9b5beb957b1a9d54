set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2c_quick.txt
for v in "" _narrow; do
  echo "== variant '$v'" >> gpurun_out/r2c_quick.txt
  ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$v.so python tools/quick_gpu_check.py 1024 >> gpurun_out/r2c_quick.txt 2>&1
done
cat gpurun_out/r2c_quick.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.txt 2>&1
tail -15 gpurun_out/r2c_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_staged -c 1 -f -o gpurun_out/r2c_dec python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2c_ncu_dec.log 2>&1
