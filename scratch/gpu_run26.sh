#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_j_bench_110m.json 2> gpurun_out/r02_j_bench_110m.err
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 > gpurun_out/r02_j_bench_600m.json 2> gpurun_out/r02_j_bench_600m.err
timeout 900 python bench.py --config eou-120m-stream > gpurun_out/r02_j_bench_stream64.json 2> gpurun_out/r02_j_bench_stream64.err
timeout 900 python bench.py --config eou-120m-stream --streams 512 --steps 125 --no-cpu-baseline > gpurun_out/r02_j_bench_stream512.json 2> gpurun_out/r02_j_bench_stream512.err
tail -4 gpurun_out/r02_pytest_gpu.log; for f in 110m 600m stream64 stream512; do head -c 330 gpurun_out/r02_j_bench_$f.json; echo; tail -1 gpurun_out/r02_j_bench_$f.err; done
