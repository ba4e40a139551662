#!/bin/bash
# N = 8: the 8192-clip job (16 steps x 64 clips x 8 GPUs), one all-gather
cd /root/repo
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 16 --warmup 3 > gpurun_out/r02_x_bench_n8.json 2> gpurun_out/r02_x_bench_n8.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_x_bench_n8.json') if l.startswith('{')][-1]); print('n8', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['config'].get('job_clips'))"; tail -2 gpurun_out/r02_x_bench_n8.err | cut -c1-200
