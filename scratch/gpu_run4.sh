#!/bin/bash
cd /root/repo
python scratch/gemm_dbg.py 0 1 > gpurun_out/r02_gemm_dbg3.log 2>&1
PK_GEMM_DBG=0 python scratch/gemm_sweep.py 0 1 > gpurun_out/r02_gemm_sweep2.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "gemm or encoder_110m or full_batch" > gpurun_out/r02_pytest_gemm.log 2>&1
cat gpurun_out/r02_gemm_dbg3.log gpurun_out/r02_gemm_sweep2.log; tail -3 gpurun_out/r02_pytest_gemm.log
