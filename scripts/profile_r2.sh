#!/bin/bash
# Round-2 ncu evidence.  Run under gpurun on ONE GPU; every ncu pass follows a plain run of the same command that
# exited 0 (a number printed under ncu is never a bench value).  Outputs stay below gpurun's 64 MiB merge limit: the
# full-section captures are exported to raw CSV on the box and only the chain capture keeps its .ncu-rep.
#  1. bench.py itself, plain                           -> gpurun_out/r2_bench_plain.json
#  2. launch list of bench.py (graph nodes per launch) -> gpurun_out/r2_launches_bench.csv
#  3. ncu --set full: the four chain launches of one CDAE update, the weight-gradient launches, the fused encoder
#     sampler                                          -> gpurun_out/r2_full_*.raw.csv (+ r2_full_chain.ncu-rep)
set -x
python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_r2_l.log 2>&1
wc -l gpurun_out/r2_launches_bench.csv
python scripts/prof_cdae.py > gpurun_out/r2_pc_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:chain -c 4 -o gpurun_out/r2_full_chain \
    python scripts/prof_cdae.py > gpurun_out/ncu_r2_chain.log 2>&1
ncu -i gpurun_out/r2_full_chain.ncu-rep --page raw --csv > gpurun_out/r2_full_chain.raw.csv
ncu --set full --clock-control none -k regex:gemm_tn16 -c 3 -o /tmp/r2_full_tn \
    python scripts/prof_cdae.py > gpurun_out/ncu_r2_tn.log 2>&1
ncu -i /tmp/r2_full_tn.ncu-rep --page raw --csv > gpurun_out/r2_full_tn.raw.csv
python scripts/prof_step.py 2 > gpurun_out/r2_ps_plain.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:enc_sample -c 1 -o /tmp/r2_full_enc \
    python scripts/prof_step.py 2 > gpurun_out/ncu_r2_enc.log 2>&1
ncu -i /tmp/r2_full_enc.ncu-rep --page raw --csv > gpurun_out/r2_full_enc.raw.csv
du -sh gpurun_out
