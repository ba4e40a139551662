#!/bin/bash
# compute-sanitizer on the tcgen05 kernels the tiny model does not reach
cd /root/repo
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python scratch/sanitize_tc.py > gpurun_out/r02_z_sanitizer_tc_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_z_sanitizer_tc_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck python scratch/sanitize_tc.py > gpurun_out/r02_z_sanitizer_tc_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_z_sanitizer_tc_racecheck.txt
tail -12 gpurun_out/r02_z_sanitizer_tc_memcheck.txt | cut -c1-200; tail -12 gpurun_out/r02_z_sanitizer_tc_racecheck.txt | cut -c1-200
