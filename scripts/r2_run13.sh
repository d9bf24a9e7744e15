#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I vcr_net_b200/csrc -o /tmp/mma_rate scripts/mma_rate.cu && timeout 60 /tmp/mma_rate | tee gpurun_out/mma_rate.txt | grep "grid 148\|TMEM (lane"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flash" | tail -5
timeout -s KILL 300 python scripts/flash_diag.py | tee gpurun_out/flash_diag.txt
