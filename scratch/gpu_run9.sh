#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 16 --warmup 3 > gpurun_out/r02_b_bench_110m.json 2> gpurun_out/r02_b_bench_110m.err
tail -4 gpurun_out/r02_pytest_gpu.log; head -c 2500 gpurun_out/r02_b_bench_110m.json; tail -3 gpurun_out/r02_b_bench_110m.err
