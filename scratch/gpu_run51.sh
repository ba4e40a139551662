#!/bin/bash
# conv1 on tcgen05: engine parity (subsampling output, layers, tokens), bench A/B
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "alternative_kernels and CONV1" > gpurun_out/r02_t51_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t51_tests.log
tail -15 gpurun_out/r02_t51_tests.log | cut -c1-250
for f in 0 1; do
  PK_CONV1_UMMA=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t51_bench_c$f.json 2> gpurun_out/r02_t51_bench_c$f.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_t51_bench_c$f.json') if l.startswith('{')][-1]); print('conv1_umma$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_class_ms_per_step'])"; tail -2 gpurun_out/r02_t51_bench_c$f.err
done
