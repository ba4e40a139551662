#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
for pdl in 1 0; do
PK_PDL=$pdl timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_m_bench_110m_pdl$pdl.json 2> gpurun_out/r02_m_bench_110m_pdl$pdl.err
PK_PDL=$pdl timeout 900 python bench.py --config eou-120m-stream --no-cpu-baseline > gpurun_out/r02_m_bench_stream64_pdl$pdl.json 2> gpurun_out/r02_m_bench_stream64_pdl$pdl.err
PK_PDL=$pdl timeout 900 python bench.py --config eou-120m-stream --streams 1 --no-cpu-baseline > gpurun_out/r02_m_bench_stream1_pdl$pdl.json 2> gpurun_out/r02_m_bench_stream1_pdl$pdl.err
done
PK_PDL=1 timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_m_bench_600m_pdl1.json 2> gpurun_out/r02_m_bench_600m_pdl1.err
tail -4 gpurun_out/r02_pytest_gpu.log; for f in 110m_pdl1 110m_pdl0 stream64_pdl1 stream64_pdl0 stream1_pdl1 stream1_pdl0 600m_pdl1; do echo $f; head -c 250 gpurun_out/r02_m_bench_$f.json; echo; tail -1 gpurun_out/r02_m_bench_$f.err; done
