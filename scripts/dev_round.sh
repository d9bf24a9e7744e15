set -x
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -x -q -k "cabi or graphed or cta_pair" > gpurun_out/dev_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/dev_pytest.log; tail -15 gpurun_out/dev_pytest.log
