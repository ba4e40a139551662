#!/bin/bash
# final validation: full GPU suite + smoke + default bench (with the chunked mel normalisation)
cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_zz_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_zz_pytest_gpu.log
tail -6 gpurun_out/r02_zz_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_zz_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_zz_smoke.log; tail -2 gpurun_out/r02_zz_smoke.log
timeout 900 python bench.py > gpurun_out/r02_zz_bench_n1.json 2> gpurun_out/r02_zz_bench_n1.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_zz_bench_n1.json') if l.startswith('{')][-1]); print('n1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_class_ms_per_step'], d['clocks'], d['gpu_launches'])"; tail -1 gpurun_out/r02_zz_bench_n1.err | cut -c1-200
