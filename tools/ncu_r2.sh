#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): launch list of the bench command, then one full capture of the
# curve kernel and of the small-batch kernel.  Outputs under gpurun_out/; summaries are written to profiles/ afterwards
# (tools/ncu_summary.py).  Numbers printed by a run under ncu are never bench values.
set -x
export SV_BENCH_QUICK=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 > gpurun_out/r2_ncu_bench_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_main -s 6 -c 1 -f -o gpurun_out/r2_k_main \
    python bench.py --steps 4 --warmup 3 > gpurun_out/r2_ncu_kmain_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_small -s 20 -c 1 -f -o gpurun_out/r2_k_small \
    python tools/latency_table.py > gpurun_out/r2_ncu_ksmall_stdout.log 2>&1
ls -la gpurun_out/*.ncu-rep
