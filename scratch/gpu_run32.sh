#!/bin/bash
cd /root/repo
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_n_110m_$tag.json 2> gpurun_out/r02_n_110m_$tag.err
  env "$@" timeout 600 python bench.py --config eou-120m-stream --streams 1 --no-cpu-baseline > gpurun_out/r02_n_stream1_$tag.json 2> gpurun_out/r02_n_stream1_$tag.err
  for w in 110m stream1; do echo "$tag $w $(python -c "import json,sys; d=json.loads([l for l in open('gpurun_out/r02_n_${w}_$tag.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['e2e']['value'])" 2>&1 | tail -1)"; done
}
run nograph_pdl0 PK_GRAPH=0 PK_PDL=0
run nograph_pdl1 PK_GRAPH=0 PK_PDL=1
run graph_notrig_pdl1 PK_PDL=1 PK_LIB=/root/repo/parakeet.cpp_b200/libparakeet_b200_notrig.so
run graph_pdl0 PK_PDL=0
run graph_pdl1 PK_PDL=1
