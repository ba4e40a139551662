#!/bin/bash
# the C++ multi-GPU example on 2 GPUs (tiny synthetic checkpoint) + the same job through the ctypes binding on one GPU for comparison
cd /root/repo
python - <<'PY'
import sys
sys.path.insert(0, '/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package()
from parakeet_cpp_b200 import synth
O = ge.load_oracle()
synth.save_safetensors('/tmp/tiny.safetensors', synth.make_weights(O.make_tiny_config(), seed=3))
PY
g++ -std=c++17 -O2 -Iinclude examples/sharded_transcribe.cpp -Lparakeet.cpp_b200 -lparakeet_b200 -Wl,-rpath,/root/repo/parakeet.cpp_b200 -lpthread -o /tmp/sharded_transcribe
/tmp/sharded_transcribe /tmp/tiny.safetensors 2 20 tiny > gpurun_out/r02_example_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_example_2gpu.log
/tmp/sharded_transcribe /tmp/tiny.safetensors 1 40 tiny > gpurun_out/r02_example_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_example_1gpu.log
cat gpurun_out/r02_example_2gpu.log gpurun_out/r02_example_1gpu.log
