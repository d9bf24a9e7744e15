set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "softcorr" > gpurun_out/dev_sc_pytest.log 2>&1; tail -15 gpurun_out/dev_sc_pytest.log
