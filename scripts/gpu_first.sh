#!/bin/bash
# first end-to-end GPU pass: environment facts, smoke, parity tests, exploratory bench
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread" 
which java javac 2>&1 | head -2
} > gpurun_out/env.txt 2>&1
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --nodes 16384 --steps 5 --warmup 2 --no-cpu > gpurun_out/bench_16k.log 2>&1; echo "rc=$?" >> gpurun_out/bench_16k.log
timeout 900 python bench.py --nodes 131072 --steps 4 --warmup 1 --no-cpu > gpurun_out/bench_131k.log 2>&1; echo "rc=$?" >> gpurun_out/bench_131k.log
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_16k.log gpurun_out/bench_131k.log
