set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2i_quick.txt
for v in "WS:" "WS:_wsnosync" ":_esync" ":"; do
  ws=${v%%:*}; lib=${v#*:}
  echo "== ws='$ws' lib='$lib'" >> gpurun_out/r2i_quick.txt
  if [ -n "$ws" ]; then export ZFP_B200_WS=1; else unset ZFP_B200_WS; fi
  ZFP_B200_LIB=zfp_b200/lib/libzfp_b200$lib.so timeout 300 python tools/quick_gpu_check.py 1024 2>&1 | grep -E "mismatch|float64|Error" >> gpurun_out/r2i_quick.txt
done
unset ZFP_B200_WS
cat gpurun_out/r2i_quick.txt
ZFP_B200_WS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_ws -c 1 -f -o gpurun_out/r2i_dec_ws python tools/prof_target.py 1024 f64 8 1 > gpurun_out/r2i_ncu.log 2>&1
