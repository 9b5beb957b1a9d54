set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2e_quick.txt
echo "== Q4=1" >> gpurun_out/r2e_quick.txt
ZFP_B200_Q4=1 timeout 300 python tools/quick_gpu_check.py 1024 >> gpurun_out/r2e_quick.txt 2>&1
cat gpurun_out/r2e_quick.txt
ZFP_B200_Q4=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r2e_pytest.txt 2>&1
tail -5 gpurun_out/r2e_pytest.txt
ZFP_B200_Q4=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_q4 -c 1 -f -o gpurun_out/r2e_dec python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2e_ncu_dec.log 2>&1
