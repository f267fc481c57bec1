#!/bin/bash
# round-2 closing measurements on one B200: tests, smoke, every bench config, latency table (outputs under gpurun_out/final_*)
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > $O/final_pytest_gpu.log; cat $O/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/final_smoke.log 2>&1; tail -2 $O/final_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/final_bench_reference.json 2> $O/final_bench_reference.err
python bench.py > $O/final_bench.json 2> $O/final_bench.err
python bench.py --config c3 --steps 20 --warmup 3 > $O/final_bench_c3.json 2> $O/final_c3.err
python bench.py --config c4 --steps 10 --warmup 3 > $O/final_bench_c4_dedup.json 2> $O/final_c4.err
SV_BENCH_NODEDUP=1 python bench.py --config c4 --steps 10 --warmup 3 > $O/final_bench_c4_nodedup.json 2>> $O/final_c4.err
python bench.py --config c5 --steps 2 --warmup 1 > $O/final_bench_c5_n1.json 2> $O/final_c5.err
python tools/latency_table.py > $O/final_latency.json 2> $O/final_latency.err
python - <<'PY'
import json
for f in ("final_bench_reference","final_bench","final_bench_c3","final_bench_c4_dedup","final_bench_c4_nodedup","final_bench_c5_n1"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().split("\n")[-1])
        print(f, d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("failed_checks"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
