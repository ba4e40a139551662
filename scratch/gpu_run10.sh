#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q -k "stream or resampler" > gpurun_out/r02_pytest_stream.log 2>&1
timeout 900 python bench.py --config eou-120m-stream > gpurun_out/r02_c_bench_stream64.json 2> gpurun_out/r02_c_bench_stream64.err
timeout 900 python bench.py --config eou-120m-stream --streams 512 --steps 125 --no-cpu-baseline > gpurun_out/r02_c_bench_stream512.json 2> gpurun_out/r02_c_bench_stream512.err
tail -4 gpurun_out/r02_pytest_stream.log; cat gpurun_out/r02_c_bench_stream64.json; tail -3 gpurun_out/r02_c_bench_stream64.err; cat gpurun_out/r02_c_bench_stream512.json; tail -3 gpurun_out/r02_c_bench_stream512.err
