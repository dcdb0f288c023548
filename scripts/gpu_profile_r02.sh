#!/bin/bash
# round 2: ncu launch list (gpu__time_duration.sum, no clock control) of 22 consecutive ticks around t = 1 590 ms of the metric run
# (GSFSignature 131 072 nodes, 19 launches per tick, one doCycle tick inside) -> profiles/r02_launches_gsf131072_t1590.csv
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 30300 -c 420 --csv --log-file gpurun_out/r02_launches_gsf131072_t1590.csv \
   python scripts/gpu_node_ticks.py 1700 > gpurun_out/r02_ncu_launch_run.log 2>&1
tail -2 gpurun_out/r02_ncu_launch_run.log | cut -c1-200
wc -l gpurun_out/r02_launches_gsf131072_t1590.csv
