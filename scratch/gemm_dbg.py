"""Where does the tcgen05 GEMM's time go?  PK_GEMM_DBG bits: 1 = the epilogue releases the accumulator without draining it,
2 = no TMA loads (MMAs on stale smem), 4 = epilogue drains TMEM but does nothing else, 8 = everything but the global stores,
16 = probe only (effective SM clock of CTA 0)."""
import os, sys
sys.path.insert(0, '/root/repo')
os.environ['PK_SELFTEST_TIME'] = '1'
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI = dict(BIAS_F32=0, RELU_F32=1, RELU_ACT=2, SILU_ACT=3, RESID=4, GLU=5, BIAS_ACT=6, QKV=7)
cases = [(8064, 2048, 512, 'SILU_ACT'), (8064, 2048, 512, 'BIAS_F32'), (8064, 512, 2048, 'RESID'), (8064, 512, 512, 'RESID'), (6016, 4096, 1024, 'SILU_ACT')]
for two in sys.argv[1:] or ('0', '1'):
    for dbg in ('16', '1', '3', '4', '8', '17'):
        os.environ['PK_GEMM_2CTA'] = two
        os.environ['PK_GEMM_DBG'] = dbg
        print(f'--- 2CTA={two} DBG={dbg}', file=sys.stderr, flush=True)
        for (M, N, K, e) in cases:
            selftest_gemm(M, N, K, EPI[e], 0)
os.environ['PK_GEMM_DBG'] = '0'
