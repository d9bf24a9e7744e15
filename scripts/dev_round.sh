set -x
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -x -q -k "graph or iter or vcrnet" > gpurun_out/dev_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/dev_pytest.log; tail -25 gpurun_out/dev_pytest.log
