#!/bin/bash
exec 2>&1
mkdir -p gpurun_out
for kn in flash_attn_ts attn_colsum_tc; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kn -s 2 -c 1 -f -o /tmp/$kn python scripts/attn_prof.py > gpurun_out/ncu_$kn.log 2>&1
ncu -i /tmp/$kn.ncu-rep --page source --csv > gpurun_out/src_$kn.csv 2>> gpurun_out/ncu_$kn.log
ncu -i /tmp/$kn.ncu-rep --page raw --csv > gpurun_out/raw_$kn.csv 2>> gpurun_out/ncu_$kn.log
python scripts/ncu_source_phases.py gpurun_out/src_$kn.csv | head -60
done
