set -x
mkdir -p gpurun_out
FUZZ_LO=2000 FUZZ_HI=2100 timeout 1500 python tools/fuzz_big.py > gpurun_out/r3t_fuzz.txt 2>&1
tail -5 gpurun_out/r3t_fuzz.txt
