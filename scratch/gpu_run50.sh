#!/bin/bash
# smoke + N = 2 bench line (final code)
cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_x_smoke.log; tail -2 gpurun_out/r02_x_smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_x_bench_n2.json 2> gpurun_out/r02_x_bench_n2.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_x_bench_n2.json') if l.startswith('{')][-1]); print('n2', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"; tail -2 gpurun_out/r02_x_bench_n2.err
