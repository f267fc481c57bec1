#!/bin/bash
# second capture of round 2: curve kernel with the per-slot slab access-policy window, and the bucket kernel of the BIP-340
# batch path.  Same export-on-the-box scheme as tools/ncu_r2.sh.
set -x
export SV_BENCH_QUICK=1
ncu --set full --clock-control none --import-source on -k regex:k_main -s 6 -c 1 -f -o /tmp/r2b_k_main \
    python bench.py --steps 4 --warmup 3 > gpurun_out/r2b_ncu_kmain_stdout.log 2>&1
SV_L2_POLICY=0 ncu --set full --clock-control none -k regex:k_main -s 6 -c 1 -f -o /tmp/r2b_k_main_nopolicy \
    python bench.py --steps 4 --warmup 3 > gpurun_out/r2b_ncu_kmain_nopolicy_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sb_window -s 1 -c 1 -f -o /tmp/r2b_k_sb_window \
    python tools/measure_n3.py > gpurun_out/r2b_ncu_sb_stdout.log 2>&1
for k in k_main k_main_nopolicy k_sb_window; do
  ncu -i /tmp/r2b_$k.ncu-rep --page raw --csv > gpurun_out/r2b_${k}_ncu_raw.csv
  python tools/ncu_summary.py gpurun_out/r2b_${k}_ncu_raw.csv > gpurun_out/r2b_${k}_ncu_summary.md
done
ncu -i /tmp/r2b_k_sb_window.ncu-rep --page source --csv > /tmp/r2b_sb_src.csv && python tools/ncu_opmix.py /tmp/r2b_sb_src.csv > gpurun_out/r2b_k_sb_window_dynamic_opmix.txt
grep -E "dram__bytes|gpu__time_duration" gpurun_out/r2b_*_summary.md
