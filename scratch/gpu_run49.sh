#!/bin/bash
# cluster / multicast GEMM: selftests, per-shape timing for cl = 1 / 2 / 4, engine parity, bench A/B
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "cluster_multicast or (alternative_kernels and CLUSTER) or tcgen05_gemm_matches" > gpurun_out/r02_t49_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t49_tests.log
tail -8 gpurun_out/r02_t49_tests.log | cut -c1-250
for cl in 1 2 4; do
PK_GEMM_CLUSTER=$cl PK_GEMM_TMA_OUT=1 PK_SELFTEST_TIME=1 timeout 300 python - > gpurun_out/r02_t49_time_cl$cl.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge; ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
for M, N, K, epi in ((8064, 2048, 512, 3), (8064, 1536, 512, 7), (8064, 1024, 512, 5), (6016, 4096, 1024, 3), (6016, 3072, 1024, 7), (6016, 2048, 1024, 5)):
    selftest_gemm(M, N, K, epi, 0)
PY
echo "== cluster $cl"; grep gemm_tc gpurun_out/r02_t49_time_cl$cl.log | cut -c1-120
done
for cl in 0 2 4; do
  PK_GEMM_CLUSTER=$cl timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t49_bench_cl$cl.json 2> gpurun_out/r02_t49_bench_cl$cl.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_t49_bench_cl$cl.json') if l.startswith('{')][-1]); print('cluster$cl', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_class_ms_per_step'])"; tail -2 gpurun_out/r02_t49_bench_cl$cl.err
done
