#!/bin/bash
cd /root/repo
export PK_SELFTEST_TIME=0
for DBG in 0 1; do
PK_GEMM_DBG=$DBG PK_GEMM_TMA_OUT=1 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/r02_gemm_fc1_dbg$DBG -f python scratch/gemm_one.py 8064 2048 512 3 0 > gpurun_out/r02_ncu_gemm_dbg$DBG.log 2>&1
done
ls -la gpurun_out/r02_gemm_fc1_dbg*
