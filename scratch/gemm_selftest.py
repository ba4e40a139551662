import sys
sys.path.insert(0,'/root/repo')
import __graft_entry__ as ge
pkg=ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
EPI=dict(BIAS_F32=0,RELU_F32=1,RELU_ACT=2,SILU_ACT=3,RESID=4,GLU=5,BIAS_ACT=6)
cases=[(128,128,64,'BIAS_F32'),(128,128,256,'BIAS_F32'),(300,256,256,'RELU_F32'),(126,1025,512,'BIAS_F32'),(777,2048,512,'SILU_ACT'),
       (513,512,2048,'RESID'),(256,1024,512,'GLU'),(130,64,64,'BIAS_ACT'),(8064,512,2560,'BIAS_F32'),(8064,2048,512,'SILU_ACT')]
for math in (0,1):
    for (M,N,K,e) in cases:
        try:
            err,ref=selftest_gemm(M,N,K,EPI[e],math)
            print(f'math={math} M={M} N={N} K={K} {e:9s} err={err:.3e} ref={ref:.3e} rel={err/ref:.2e}', flush=True)
        except Exception as ex:
            print('FAIL', math,M,N,K,e,ex, flush=True); break
