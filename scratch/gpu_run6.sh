#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q -k "gemm" > gpurun_out/r02_pytest_gemm.log 2>&1
python scratch/gemm_tl.py 32 96 > gpurun_out/r02_gemm_tl2.log 2>&1
PK_GEMM_DBG=0 python scratch/gemm_sweep.py 0 1 > gpurun_out/r02_gemm_sweep3.log 2>&1
tail -5 gpurun_out/r02_pytest_gemm.log; grep -v "tile  7" gpurun_out/r02_gemm_tl2.log | cut -c1-160; grep -v "err/ref" gpurun_out/r02_gemm_sweep3.log
