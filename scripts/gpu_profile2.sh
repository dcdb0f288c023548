#!/bin/bash
mkdir -p gpurun_out
# launch list of 21 consecutive ticks (one doCycle tick) in the heavy phase: 19 launches per tick
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 30210 -c 400 --csv --log-file gpurun_out/launches_131k_v3.csv \
   python scripts/gpu_trace.py 131072 1700 100 > gpurun_out/ncu_launch_run.log 2>&1
for k in k_cond_scan k_cond_score k_cond_select k_node; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s 1410 -c 1 -o gpurun_out/prof3_$k \
   python scripts/gpu_trace.py 32768 1500 100 > gpurun_out/ncu_${k}_run.log 2>&1
done
ls -la gpurun_out | tail -8
