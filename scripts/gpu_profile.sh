#!/bin/bash
# round-1 profiling pass: per-kernel event timing in the heavy phase, ncu launch list, ncu full capture
mkdir -p gpurun_out
timeout 600 python scripts/gpu_trace.py 131072 1800 100 1500 > gpurun_out/trace_prof_131k.log 2>&1
# launch list of the heavy phase (tick ~1600): 17 launches per tick
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 27200 -c 340 --csv --log-file gpurun_out/launches_131k.csv \
   python scripts/gpu_trace.py 131072 1700 100 > gpurun_out/ncu_launch_run.log 2>&1
# full capture of the two handler kernels at 32768 nodes (kernel replay), heavy phase
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_cond -s 1400 -c 2 -o gpurun_out/prof_k_cond \
   python scripts/gpu_trace.py 32768 1500 100 > gpurun_out/ncu_cond_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_node -s 1400 -c 2 -o gpurun_out/prof_k_node \
   python scripts/gpu_trace.py 32768 1500 100 > gpurun_out/ncu_node_run.log 2>&1
ls -la gpurun_out/
tail -2 gpurun_out/trace_prof_131k.log
