#!/bin/bash
# compute-sanitizer on the tiny configuration with the final code (new mel normalisation kernels included): offline CTC / TDT, boosted, resampled, streaming
cd /root/repo
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python scratch/sanitize_tiny.py > gpurun_out/r02_zz_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_zz_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python scratch/sanitize_tiny.py > gpurun_out/r02_zz_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_zz_sanitizer_racecheck.txt
tail -6 gpurun_out/r02_zz_sanitizer_memcheck.txt | cut -c1-200; tail -8 gpurun_out/r02_zz_sanitizer_racecheck.txt | cut -c1-200
