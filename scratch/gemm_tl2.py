import os, sys
sys.path.insert(0, '/root/repo')
os.environ['PK_SELFTEST_TIME'] = '1'
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI = dict(BIAS_F32=0, RELU_F32=1, RELU_ACT=2, SILU_ACT=3, RESID=4, GLU=5, BIAS_ACT=6, QKV=7)
for two, dbg in [a.split(':') for a in sys.argv[1:]]:
    os.environ['PK_GEMM_2CTA'] = two
    os.environ['PK_GEMM_DBG'] = dbg
    print(f'--- 2CTA={two} DBG={dbg}', file=sys.stderr, flush=True)
    for (M, N, K, e) in [(8064, 2048, 512, 'SILU_ACT'), (8064, 512, 2048, 'RESID')]:
        selftest_gemm(M, N, K, EPI[e], 0)
