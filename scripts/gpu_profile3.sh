#!/bin/bash
# full ncu captures of the handler kernels in the heavy phase of GSF-131072 (tick ~1000)
mkdir -p gpurun_out
for k in k_node_tasks k_node_msgs k_cond_score k_emit k_ms_scatter; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1003 -c 1 -o gpurun_out/prof4_$k \
   python scripts/gpu_trace.py 131072 1100 100 > gpurun_out/ncu_${k}_run.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cond_nodes -s 2006 -c 2 -o gpurun_out/prof4_k_cond_nodes \
   python scripts/gpu_trace.py 131072 1100 100 > gpurun_out/ncu_k_cond_nodes_run.log 2>&1
ls -la gpurun_out | tail -8
