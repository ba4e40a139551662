#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "600m or tdt or boost or stream or transcribe or full_batch" > gpurun_out/r02_pytest_tdt.log 2>&1
python scratch/tdt_phases2.py > gpurun_out/r02_tdt_phases3.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k_bench_110m.json 2> gpurun_out/r02_k_bench_110m.err
tail -4 gpurun_out/r02_pytest_tdt.log; cat gpurun_out/r02_tdt_phases3.log; head -c 300 gpurun_out/r02_k_bench_110m.json; tail -2 gpurun_out/r02_k_bench_110m.err
