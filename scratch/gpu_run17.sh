#!/bin/bash
# evidence pass: launch list, ncu of the dominant GEMM inside the real step, sanitizer logs, final bench lines
cd /root/repo
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scratch/sanitize_tiny.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python scratch/sanitize_tiny.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 2000 --csv --log-file gpurun_out/r02_f_launches.csv python scratch/one_step.py ctc 2 > gpurun_out/r02_f_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -s 150 -c 12 -f -o gpurun_out/r02_f_gemm_in_step python scratch/one_step.py ctc 2 > gpurun_out/r02_f_gemm_ncu.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_f_bench_110m.json 2> gpurun_out/r02_f_bench_110m.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_f_bench_reference_arm.json 2> gpurun_out/r02_f_bench_reference_arm.err
timeout 900 python bench.py --config 600m-16x30s --steps 8 --warmup 3 > gpurun_out/r02_f_bench_600m.json 2> gpurun_out/r02_f_bench_600m.err
tail -5 gpurun_out/r02_sanitizer_memcheck.log; tail -8 gpurun_out/r02_sanitizer_racecheck.log; tail -2 gpurun_out/r02_f_launches.log; tail -2 gpurun_out/r02_f_gemm_ncu.log; head -c 400 gpurun_out/r02_f_bench_110m.json; echo; head -c 400 gpurun_out/r02_f_bench_600m.json; echo; cat gpurun_out/r02_f_bench_reference_arm.json | head -c 600
