#!/bin/bash
set -x
export SV_BENCH_QUICK=1
ncu --set full --clock-control none -k regex:k_main -s 6 -c 1 -f -o /tmp/r2c_k_main python bench.py --steps 4 --warmup 3 > gpurun_out/r2c_ncu_kmain_stdout.log 2>&1
ncu -i /tmp/r2c_k_main.ncu-rep --page raw --csv > gpurun_out/r2c_k_main_ncu_raw.csv
python tools/ncu_summary.py gpurun_out/r2c_k_main_ncu_raw.csv > gpurun_out/r2c_k_main_ncu_summary.md
grep -E "dram__bytes|gpu__time_duration|lts__t" gpurun_out/r2c_k_main_ncu_summary.md
