#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flash or colsum" | tail -3
timeout -s KILL 300 python scripts/colsum_diag.py | tee gpurun_out/colsum_diag.txt
