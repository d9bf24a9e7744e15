set -x
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/knn3_diag.py > gpurun_out/dev_knn3_diag.txt 2>&1; echo "rc=$?" >> gpurun_out/dev_knn3_diag.txt
cat gpurun_out/dev_knn3_diag.txt
timeout -s KILL 200 python -m pytest tests -m gpu -x -q -k "knn" > gpurun_out/dev_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/dev_pytest.log; tail -8 gpurun_out/dev_pytest.log
