#!/bin/bash
# ncu evidence for round 1 (run under gpurun from the repo root)
set -x
python scripts/prof_cdae.py > gpurun_out/prof_plain.log 2>&1 || exit 1
# every launch of the second train call with its duration
ncu --metrics gpu__time_duration.sum --clock-control none -s 80 -c 90 --csv --log-file gpurun_out/launches_cdae.csv python scripts/prof_cdae.py > gpurun_out/ncu_l.log 2>&1
# full sections for one representative of each GEMM flavour (indices inside the 2nd call: 66 gemm launches per call)
i=0
for skip in 73 83 97 107 112; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:gemm_ -s $skip -c 1 -o gpurun_out/prof_r1_$i python scripts/prof_cdae.py > gpurun_out/ncu_f$i.log 2>&1
done
ls -la gpurun_out
