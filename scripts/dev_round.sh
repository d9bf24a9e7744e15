set -x
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/pair_diag.py > gpurun_out/dev_pair_diag4.txt 2>&1; echo "rc=$?" >> gpurun_out/dev_pair_diag4.txt
tail -40 gpurun_out/dev_pair_diag4.txt
nvidia-smi --query-gpu=name,temperature.gpu --format=csv
