#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/flash_diag.py | tee gpurun_out/flash_diag.txt | cut -c1-60,200-400
