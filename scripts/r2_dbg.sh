#!/bin/bash
mkdir -p gpurun_out
echo "=== alone h3"; timeout 600 python -m pytest tests/test_gpu_headline.py -q -x -k "cfg2 and h3" 2>&1 | grep -v "parity achieved" | tail -4
echo "=== fp32 then h3"; timeout 600 python -m pytest tests/test_gpu_headline.py -q -x -k "cfg2" 2>&1 | grep -v "parity achieved" | tail -4
echo "=== blocking"; CUDA_LAUNCH_BLOCKING=1 timeout 600 python -m pytest tests/test_gpu_headline.py -q -x -k "cfg2" 2>&1 | grep -n "failed:\|Error\|vcr_net_b200/.*: in " | head -20
