import os, sys
sys.path.insert(0, '/root/repo')
os.environ['PK_SELFTEST_TIME'] = '1'
os.environ['PK_GEMM_TMA_OUT'] = '1'
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI = dict(BIAS_F32=0, RELU_F32=1, RELU_ACT=2, SILU_ACT=3, RESID=4, GLU=5, BIAS_ACT=6, QKV=7)
os.environ['PK_GEMM_2CTA'] = '0'
for dbg in ('16', '1040', '2064'):
    os.environ['PK_GEMM_DBG'] = dbg
    print(f'--- DBG={dbg} (hint mode {int(dbg) >> 10}: 0 loads EVICT_LAST, 1 none, 2 + stores EVICT_FIRST)', file=sys.stderr, flush=True)
    for (M, N, K, e) in [(8064, 2048, 512, 'SILU_ACT'), (8064, 512, 2048, 'RESID'), (8064, 1536, 512, 'QKV'), (8064, 512, 512, 'RESID'), (8064, 1024, 512, 'GLU'), (6016, 4096, 1024, 'SILU_ACT')]:
        selftest_gemm(M, N, K, EPI[e], 0)
