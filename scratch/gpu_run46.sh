#!/bin/bash
# tcgen05 attention: selftests vs the fp32 kernel, timing vs mma.sync, engine parity, bench A/B
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "tcgen05_attention" > gpurun_out/r02_t46_selftest.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t46_selftest.log
tail -30 gpurun_out/r02_t46_selftest.log | cut -c1-220
PK_AU_DBG=1 PK_SELFTEST_TIME=1 timeout 300 python - > gpurun_out/r02_t46_time.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge; ge.load_package()
from parakeet_cpp_b200.engine import selftest_attention
for lens, tmax, mode in (([126], 126, 1), ([126], 126, 2), ([126], 126, 0), ([126] * 64, 126, 0)):
    try:
        print(len(lens), tmax, mode, selftest_attention(lens, tmax, mode), flush=True)
    except Exception as e:
        print('FAILED', len(lens), tmax, mode, e, flush=True)
PY
cat gpurun_out/r02_t46_time.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "alternative_kernels and ATTN" > gpurun_out/r02_t46_engine.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t46_engine.log
tail -12 gpurun_out/r02_t46_engine.log | cut -c1-220
for f in 0 1; do
  PK_ATTN_UMMA=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t46_bench_umma$f.json 2> gpurun_out/r02_t46_bench_umma$f.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_t46_bench_umma$f.json') if l.startswith('{')][-1]); print('umma$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_class_ms_per_step'])"; tail -2 gpurun_out/r02_t46_bench_umma$f.err
done
