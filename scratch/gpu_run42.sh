#!/bin/bash
# fused GEMM+LayerNorm v3 (full-line stores, A multicast): selftests, timeline, bench A/B
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_gemm_layernorm or fused_layernorm_engine" > gpurun_out/r02_t42_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_t42_tests.log
tail -5 gpurun_out/r02_t42_tests.log
for mc in 1 0; do
PK_LN_MCAST=$mc PK_SELFTEST_TIME=1 PK_LN_DBG=1 timeout 300 python - > gpurun_out/r02_t42_time_mc$mc.log 2>&1 <<'PY'
import sys; sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge; ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm_ln
for M, K, mode in ((8064, 2048, 0), (8064, 512, 0), (8064, 2048, 1), (8064, 2560, 3)):
    print(M, K, mode, selftest_gemm_ln(M, K, mode, 0), flush=True)
PY
echo "== mcast $mc"; cat gpurun_out/r02_t42_time_mc$mc.log
done
for f in "0 1" "1 1" "1 0"; do
  set -- $f
  PK_FUSE_LN=$1 PK_LN_MCAST=$2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_t42_bench_fuse$1_mc$2.json 2> gpurun_out/r02_t42_bench_fuse$1_mc$2.err
  python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_t42_bench_fuse$1_mc$2.json') if l.startswith('{')][-1]); print('fuse$1 mc$2', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['per_class_ms_per_step'])"; tail -2 gpurun_out/r02_t42_bench_fuse$1_mc$2.err
done
