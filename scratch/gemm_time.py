import sys, os
sys.path.insert(0,'/root/repo')
os.environ['PK_SELFTEST_TIME']='1'
import __graft_entry__ as ge
pkg=ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI=dict(BIAS_F32=0,RELU_F32=1,RELU_ACT=2,SILU_ACT=3,RESID=4,GLU=5,BIAS_ACT=6,QKV=7)
cases=[(8064,1536,512,'QKV'),(8064,2048,512,'SILU_ACT'),(8064,512,2048,'RESID'),(8064,1536,512,'BIAS_F32'),(8064,512,512,'RESID'),(8064,1024,512,'GLU'),(16128,2048,512,'SILU_ACT'),(16128,512,2048,'RESID'),(8064,2048,2048,'BIAS_F32')]
for math in (0,1):
    for (M,N,K,e) in cases:
        selftest_gemm(M,N,K,EPI[e],math)
