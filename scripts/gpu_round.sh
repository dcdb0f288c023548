#!/bin/bash
# round-style GPU check: tests, smoke, bench (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -c 300 gpurun_out/bench_reference.err
cat gpurun_out/bench_default.json | cut -c1-3000
cat gpurun_out/bench_reference.json | cut -c1-1200
