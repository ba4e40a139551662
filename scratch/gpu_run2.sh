#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python scratch/gemm_dbg.py > gpurun_out/r02_gemm_dbg.log 2>&1
timeout 600 python bench.py --steps 16 --warmup 3 > gpurun_out/r02_a_bench_110m.json 2> gpurun_out/r02_a_bench_110m.err
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_a_bench_600m.json 2> gpurun_out/r02_a_bench_600m.err
cat gpurun_out/r02_gemm_dbg.log; head -c 3000 gpurun_out/r02_a_bench_110m.json; tail -3 gpurun_out/r02_a_bench_110m.err; head -c 3000 gpurun_out/r02_a_bench_600m.json; tail -3 gpurun_out/r02_a_bench_600m.err
