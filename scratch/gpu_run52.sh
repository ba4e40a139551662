#!/bin/bash
# final bench lines with the CPU baseline leg (default command lines)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_y_bench_n1.json 2> gpurun_out/r02_y_bench_n1.err
timeout 900 python bench.py --steps 128 --warmup 3 > gpurun_out/r02_y_bench_n1_128steps.json 2> gpurun_out/r02_y_bench_n1_128steps.err
timeout 1500 python bench.py --config 600m-16x30s > gpurun_out/r02_y_bench_600m_16x30s.json 2> gpurun_out/r02_y_bench_600m_16x30s.err
for f in n1 n1_128steps 600m_16x30s; do python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_y_bench_$f.json') if l.startswith('{')][-1]); print('$f', d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d.get('cpu_baseline',{}).get('value'))"; tail -1 gpurun_out/r02_y_bench_$f.err | cut -c1-200; done
