#!/bin/bash
# Round-1 (chain kernels) ncu evidence.  Run under gpurun on ONE GPU, after the plain commands exited 0.
#  1. launch list of one full train step (same workload as bench.py: scripts/prof_step.py)
#  2. ncu --set full of the five fused-chain launches of one CDAE update (3xTF32 U-chain, 3xTF32 p-chain, score,
#     tangent, adjoint) and of one [H,H] weight-gradient contraction
set -x
python scripts/prof_step.py 3 > gpurun_out/ps_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r1c_launches_step.csv \
    python scripts/prof_step.py 3 > gpurun_out/ncu_r1c_l.log 2>&1
python scripts/prof_cdae.py > gpurun_out/pc_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 5 -o gpurun_out/r1c_full_chain \
    python scripts/prof_cdae.py > gpurun_out/ncu_r1c_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tn_kernel -s 8 -c 1 -o gpurun_out/r1c_full_tn \
    python scripts/prof_cdae.py > gpurun_out/ncu_r1c_tn.log 2>&1
ls -la gpurun_out/*.ncu-rep
