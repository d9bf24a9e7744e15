set -x
mkdir -p gpurun_out
timeout -s KILL 120 python scripts/flash_diag.py > gpurun_out/dev_flash_diag.txt 2>&1; echo "rc=$?" >> gpurun_out/dev_flash_diag.txt
cat gpurun_out/dev_flash_diag.txt
