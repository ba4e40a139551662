#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
python scratch/gemm_tma.py > gpurun_out/r02_gemm_tma3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_q_bench_110m.json 2> gpurun_out/r02_q_bench_110m.err
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_q_bench_600m.json 2> gpurun_out/r02_q_bench_600m.err
tail -4 gpurun_out/r02_pytest_gpu.log
for f in 110m 600m; do python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_q_bench_$f.json') if l.startswith('{')][-1]); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_class_ms_per_step'])"; tail -1 gpurun_out/r02_q_bench_$f.err; done
grep -A9 "N=1536" gpurun_out/r02_gemm_tma3.log | cut -c1-200 | head -44
