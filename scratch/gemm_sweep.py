"""GEMM shape sweep on the GPU: the 1-CTA (128 x BN) and the cta_group::2 (256 x 256) tcgen05 kernels, warm back-to-back
timing from pk_selftest_gemm (PK_SELFTEST_TIME=1 prints us + TFLOP/s to stderr)."""
import os, sys
sys.path.insert(0, '/root/repo')
os.environ['PK_SELFTEST_TIME'] = '1'
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI = dict(BIAS_F32=0, RELU_F32=1, RELU_ACT=2, SILU_ACT=3, RESID=4, GLU=5, BIAS_ACT=6, QKV=7)
cases = [(8064, 2048, 512, 'SILU_ACT'), (8064, 512, 2048, 'RESID'), (8064, 1536, 512, 'QKV'), (8064, 512, 512, 'RESID'),
         (8064, 1024, 512, 'GLU'), (8064, 512, 2560, 'BIAS_F32'), (80640, 256, 256, 'RELU_ACT'), (8064, 640, 512, 'BIAS_F32'),
         (6016, 4096, 1024, 'SILU_ACT'), (6016, 1024, 4096, 'RESID'), (6016, 3072, 1024, 'QKV'), (6016, 1024, 1024, 'RESID'),
         (6016, 2048, 1024, 'GLU')]
modes = sys.argv[1:] or ['0', '1']
for two in modes:
    os.environ['PK_GEMM_2CTA'] = two
    print(f'--- PK_GEMM_2CTA={two}', file=sys.stderr, flush=True)
    for (M, N, K, e) in cases:
        err, ref = selftest_gemm(M, N, K, EPI[e], 0)
        print(f'    err/ref {err / ref:.2e}', file=sys.stderr, flush=True)
