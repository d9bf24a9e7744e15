set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "knn" > gpurun_out/dev_knn_pytest.log 2>&1; tail -15 gpurun_out/dev_knn_pytest.log
timeout 200 python scripts/knn_bench.py > gpurun_out/dev_knn_bench.txt 2>&1; cat gpurun_out/dev_knn_bench.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "transform_nets" > gpurun_out/dev_tnet_pytest.log 2>&1; tail -15 gpurun_out/dev_tnet_pytest.log
