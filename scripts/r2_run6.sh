#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flash" | tail -5
timeout -s KILL 300 python scripts/flash_diag.py | tee gpurun_out/flash_diag.txt
