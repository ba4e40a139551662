import os, sys
sys.path.insert(0, '/root/repo')
os.environ['PK_SELFTEST_TIME'] = '1'
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI = dict(BIAS_F32=0, RELU_F32=1, RELU_ACT=2, SILU_ACT=3, RESID=4, GLU=5, BIAS_ACT=6, QKV=7)
cases = [(8064, 2048, 512, 'SILU_ACT'), (8064, 1536, 512, 'QKV'), (8064, 1024, 512, 'GLU'), (8064, 512, 2560, 'BIAS_F32'), (80640, 256, 256, 'RELU_ACT'),
         (777, 2048, 512, 'SILU_ACT'), (300, 384, 128, 'QKV'), (300, 256, 256, 'RELU_F32'), (6016, 4096, 1024, 'SILU_ACT')]
for two in ('0', '1'):
    for tma in ('0', '1'):
        os.environ['PK_GEMM_2CTA'] = two
        os.environ['PK_GEMM_TMA_OUT'] = tma
        os.environ['PK_GEMM_DBG'] = '32' if two == '0' else '0'
        print(f'--- 2CTA={two} TMA_OUT={tma}', file=sys.stderr, flush=True)
        for (M, N, K, e) in cases:
            err, ref = selftest_gemm(M, N, K, EPI[e], 0)
            print(f'    err/ref {err / ref:.2e}', file=sys.stderr, flush=True)
