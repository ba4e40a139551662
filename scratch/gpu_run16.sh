#!/bin/bash
cd /root/repo
python scratch/gemm_hint.py > gpurun_out/r02_gemm_hint.log 2>&1
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/r02_e_bench_110m.json 2> gpurun_out/r02_e_bench_110m.err
cat gpurun_out/r02_gemm_hint.log; head -c 600 gpurun_out/r02_e_bench_110m.json; tail -3 gpurun_out/r02_e_bench_110m.err
