#!/bin/bash
# ncu launch list of 22 consecutive ticks (one doCycle tick) in the heavy phase of GSF-131072, final round-1 pipeline
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30300 -c 420 --csv --log-file gpurun_out/launches_131k_v6.csv \
   python scripts/gpu_trace.py 131072 1700 100 > gpurun_out/ncu_launch_run_v6.log 2>&1
tail -2 gpurun_out/ncu_launch_run_v6.log | cut -c1-200
wc -l gpurun_out/launches_131k_v6.csv
