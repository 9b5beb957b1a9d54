set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "foreign or untrusted or index" > gpurun_out/r3m_pytest.txt 2>&1
tail -15 gpurun_out/r3m_pytest.txt
timeout 600 python tools/gpu/index_rebuild_time.py > gpurun_out/r3m_index.txt 2>&1
cat gpurun_out/r3m_index.txt
