import sys, os
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200.engine import selftest_gemm
M, N, K, e, math = (int(x) for x in sys.argv[1:6])
print(selftest_gemm(M, N, K, e, math))
