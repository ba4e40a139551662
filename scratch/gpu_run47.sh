#!/bin/bash
# full GPU suite with the tcgen05 attention on by default + bench lines of the three single-GPU configs
cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_t47_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t47_pytest_gpu.log
tail -6 gpurun_out/r02_t47_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_t47_bench_110m.json 2> gpurun_out/r02_t47_bench_110m.err
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t47_bench_600m.json 2> gpurun_out/r02_t47_bench_600m.err
timeout 600 python bench.py --config eou-120m-stream --no-cpu-baseline > gpurun_out/r02_t47_bench_stream.json 2> gpurun_out/r02_t47_bench_stream.err
for f in 110m 600m stream; do python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_t47_bench_$f.json') if l.startswith('{')][-1]); print('$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline'].get('per_class_ms_per_step'), d.get('clocks'))"; tail -1 gpurun_out/r02_t47_bench_$f.err; done
