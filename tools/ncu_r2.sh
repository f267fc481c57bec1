#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): launch list of the bench command, then one full capture of the
# curve kernel and of the small-batch kernel.  The .ncu-rep files are too large to travel back (gpurun_out/ is capped at
# 64 MiB), so the raw page, the per-kernel summary and the executed-opcode mix are exported on the box and the reports
# deleted.  Numbers printed by a run under ncu are never bench values.
set -x
export SV_BENCH_QUICK=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 > gpurun_out/r2_ncu_bench_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_main -s 6 -c 1 -f -o /tmp/r2_k_main \
    python bench.py --steps 4 --warmup 3 > gpurun_out/r2_ncu_kmain_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_small -s 20 -c 1 -f -o /tmp/r2_k_small \
    python tools/latency_table.py > gpurun_out/r2_ncu_ksmall_stdout.log 2>&1
for k in k_main k_small; do
  ncu -i /tmp/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_${k}_ncu_raw.csv
  ncu -i /tmp/r2_$k.ncu-rep --page source --csv > /tmp/r2_${k}_src.csv
  python tools/ncu_opmix.py /tmp/r2_${k}_src.csv > gpurun_out/r2_${k}_dynamic_opmix.txt
  python tools/ncu_summary.py gpurun_out/r2_${k}_ncu_raw.csv > gpurun_out/r2_${k}_ncu_summary.md
done
ls -la gpurun_out/ | tail -20
