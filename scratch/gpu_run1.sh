#!/bin/bash
# first GPU pass of round 2: tests, GEMM sweep, both bench configs, ncu of the small kernels
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
timeout 300 python scratch/gemm_sweep.py > gpurun_out/r02_gemm_sweep.log 2>&1
timeout 600 python bench.py --steps 16 --warmup 3 > gpurun_out/r02_a_bench_110m.json 2> gpurun_out/r02_a_bench_110m.err
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_a_bench_600m.json 2> gpurun_out/r02_a_bench_600m.err
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:mel_|layernorm|dwconv_bn_silu|subsample_|relpos_attention|ctc_' -c 36 -o gpurun_out/r02_small_kernels python scratch/one_step.py ctc 1 > gpurun_out/r02_ncu_small.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; tail -30 gpurun_out/r02_gemm_sweep.log; cat gpurun_out/r02_a_bench_110m.json | head -c 1500; tail -2 gpurun_out/r02_a_bench_110m.err;  cat gpurun_out/r02_a_bench_600m.json | head -c 1500; tail -3 gpurun_out/r02_a_bench_600m.err
