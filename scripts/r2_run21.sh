#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flash or transformer or mha" | tail -3
timeout -s KILL 300 python scripts/flash_diag.py | tee gpurun_out/flash_diag.txt | cut -c1-150
timeout 300 python scripts/flash_fit.py h3 | tail -2
