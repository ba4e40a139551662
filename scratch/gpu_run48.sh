#!/bin/bash
# evidence: ncu --set full of the tcgen05 attention and of the fused GEMM+LayerNorm kernel inside a step, launch list, sustained-clock bench
cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:relpos_attention_umma -s 2 -c 2 -o gpurun_out/r02_v_attn_umma -f python scratch/one_step.py ctc 1 > gpurun_out/r02_v_ncu_attn.log 2>&1
PK_FUSE_LN=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_ln -s 2 -c 4 -o gpurun_out/r02_v_gemm_ln -f python scratch/one_step.py ctc 1 > gpurun_out/r02_v_ncu_gemm_ln.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r02_v_launches.csv python scratch/one_step.py ctc 1 > gpurun_out/r02_v_launches.log 2>&1
ls -la gpurun_out/r02_v_*
timeout 900 python bench.py --steps 128 --warmup 3 --no-cpu-baseline > gpurun_out/r02_v_bench_110m_128steps.json 2> gpurun_out/r02_v_bench_110m_128steps.err
python -c "import json; d=json.loads([l for l in open('gpurun_out/r02_v_bench_110m_128steps.json') if l.startswith('{')][-1]); print('128 steps', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'], d['wall_s'])"; tail -2 gpurun_out/r02_v_bench_110m_128steps.err
