#!/bin/bash
# ncu full-section captures of single GEMM launches from the native self-test binary (round 1, after the lean epilogue)
set -x
./build/gemm_selftest one 2 768 1 1 > gpurun_out/one_a.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 5 -c 1 -o gpurun_out/prof_r1b_3x_cg2 ./build/gemm_selftest one 2 768 1 1 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 5 -c 1 -o gpurun_out/prof_r1b_mulsig ./build/gemm_selftest one 3 256 0 -1 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 5 -c 1 -o gpurun_out/prof_r1b_tangent ./build/gemm_selftest one 5 256 0 -1 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
