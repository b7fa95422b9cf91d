#!/bin/bash
# Round-2 ncu evidence.  Run under gpurun on ONE GPU; every ncu pass follows a plain run of the same command that
# exited 0 (a number printed under ncu is never a bench value).
#  1. bench.py itself, plain                         -> gpurun_out/r2_bench_plain.json
#  2. launch list of bench.py (graph nodes per launch) -> gpurun_out/r2_launches_bench.csv
#  3. ncu --set full of the four chain launches of one CDAE update, of the multi-layer weight-gradient launch, and of
#     the fused encoder sampler                      -> gpurun_out/r2_full_*.ncu-rep
set -x
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_r2_l.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv
python scripts/prof_cdae.py > gpurun_out/r2_pc_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:chain -c 4 -o gpurun_out/r2_full_chain \
    python scripts/prof_cdae.py > gpurun_out/ncu_r2_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn16 -c 3 -o gpurun_out/r2_full_tn \
    python scripts/prof_cdae.py > gpurun_out/ncu_r2_tn.log 2>&1
python scripts/prof_step.py 2 > gpurun_out/r2_ps_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:enc_sample -c 1 -o gpurun_out/r2_full_enc \
    python scripts/prof_step.py 2 > gpurun_out/ncu_r2_enc.log 2>&1
ls -la gpurun_out/*.ncu-rep
