#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --config eou-120m-stream --no-cpu-baseline > gpurun_out/r02_i_bench_stream64.json 2> gpurun_out/r02_i_bench_stream64.err
timeout 900 python bench.py --config eou-120m-stream --streams 512 --steps 125 --no-cpu-baseline > gpurun_out/r02_i_bench_stream512.json 2> gpurun_out/r02_i_bench_stream512.err
tail -5 gpurun_out/r02_pytest_gpu.log; head -c 500 gpurun_out/r02_i_bench_stream64.json; echo; tail -2 gpurun_out/r02_i_bench_stream64.err; head -c 400 gpurun_out/r02_i_bench_stream512.json; echo
