#!/bin/bash
# Final round-1 evidence, same command as the driver's bench: launch list of `bench.py` itself under ncu
# (graph-replayed kernels are profiled per node), after the plain command exited 0.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1d_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1d.log 2>&1
tail -2 gpurun_out/ncu_r1d.log
wc -l gpurun_out/r1d_launches_bench.csv
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
