#!/bin/bash
cd /root/repo
python scratch/gemm_tma.py > gpurun_out/r02_gemm_tma2.log 2>&1
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_o_110m_$tag.json 2> gpurun_out/r02_o_110m_$tag.err
  env "$@" timeout 600 python bench.py --config eou-120m-stream --streams 1 --no-cpu-baseline > gpurun_out/r02_o_stream1_$tag.json 2> gpurun_out/r02_o_stream1_$tag.err
  for w in 110m stream1; do echo "$tag $w $(python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/r02_o_${w}_$tag.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['e2e']['value'], d['roofline'].get('per_class_ms_per_step',{}).get('gemm'))" 2>&1 | tail -1)"; done
}
run pdl0 PK_PDL=0
run latetrig_pdl1 PK_PDL=1 PK_LIB=/root/repo/parakeet.cpp_b200/libparakeet_b200_latetrig.so
run notrig_pdl1 PK_PDL=1 PK_LIB=/root/repo/parakeet.cpp_b200/libparakeet_b200_notrig.so
run pdl0_b PK_PDL=0
grep -A3 "TMA_OUT=1" gpurun_out/r02_gemm_tma2.log | head -30
