set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2m_quick.txt
for pad in 0 61440; do
echo "== pad $pad" >> gpurun_out/r2m_quick.txt
ZFP_B200_SMEM_PAD=$pad timeout 300 python tools/quick_gpu_check.py 1024 2>&1 | grep -E 'mismatch|float64|Error' >> gpurun_out/r2m_quick.txt
done
cat gpurun_out/r2m_quick.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.txt 2>&1
tail -15 gpurun_out/r2m_pytest.txt
