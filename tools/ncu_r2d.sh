#!/bin/bash
# round 2, final kernel set: launch list of a short bench run + one full capture of the curve kernel (no-sqrt flow)
set -x
export SV_BENCH_QUICK=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 4 --warmup 3 > gpurun_out/r2d_launches_stdout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_main -s 6 -c 1 -f -o /tmp/r2d_k_main python bench.py --steps 4 --warmup 3 > gpurun_out/r2d_ncu_kmain_stdout.log 2>&1
ncu -i /tmp/r2d_k_main.ncu-rep --page raw --csv > gpurun_out/r2d_k_main_ncu_raw.csv
python tools/ncu_summary.py gpurun_out/r2d_k_main_ncu_raw.csv > gpurun_out/r2d_k_main_ncu_summary.md
ncu -i /tmp/r2d_k_main.ncu-rep --page source --csv > /tmp/r2d_src.csv
python tools/ncu_opmix.py /tmp/r2d_src.csv > gpurun_out/r2d_k_main_opmix.md 2>&1
grep -E "dram__bytes|gpu__time_duration|lts__t|fmaheavy|inst_executed.sum " gpurun_out/r2d_k_main_ncu_summary.md
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r2d_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u in ("nsecond", "ns") else (v / 1e3 if u in ("usecond", "us") else v)
    k = r[ki].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:60s} x{c:3d}  total {t:9.3f} ms  avg {t/c:8.3f} ms")
PY
