"""compute-sanitizer target for the tcgen05 kernels that the tiny model does not reach: the attention (attention_umma.cu), the fused
GEMM + LayerNorm (gemm_tc_ln.cu: cluster, st.async statistics), the persistent GEMM and its cluster / multicast variant.
    compute-sanitizer --tool memcheck|racecheck python scratch/sanitize_tc.py"""
import os, sys
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
ge.load_package()
from parakeet_cpp_b200.engine import selftest_attention, selftest_gemm, selftest_gemm_ln
print('attention', selftest_attention([126, 40, 128], tmax=128))
print('gemm+ln mode 0', selftest_gemm_ln(300, 512, 0, 0))
print('gemm+ln mode 1', selftest_gemm_ln(300, 512, 1, 0))
print('gemm silu', selftest_gemm(300, 2048, 512, 3, 0))
os.environ['PK_GEMM_CLUSTER'] = '2'
print('gemm silu, 2-CTA clusters', selftest_gemm(300, 2048, 512, 3, 0))
print('done')
