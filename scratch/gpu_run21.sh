#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -k "600m or tdt or boost or stream" > gpurun_out/r02_pytest_tdt.log 2>&1
python scratch/tdt_phases2.py 600m-16x30s > gpurun_out/r02_tdt_phases2.log 2>&1
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02_h_bench_600m.json 2> gpurun_out/r02_h_bench_600m.err
tail -4 gpurun_out/r02_pytest_tdt.log; cat gpurun_out/r02_tdt_phases2.log; head -c 300 gpurun_out/r02_h_bench_600m.json; tail -2 gpurun_out/r02_h_bench_600m.err
