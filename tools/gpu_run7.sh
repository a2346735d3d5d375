#!/bin/bash
# one GPU: the default bench line (pipelined e2e, bits/base, same-config CPU baseline), the reference arm, the new tests
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python bench.py > $O/s7_bench.json 2> $O/s7_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > $O/s7_ref.json 2> $O/s7_ref.err; echo "ref rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/s7_pytest.log 2>&1; echo "pytest rc=$?" >> $O/s7_pytest.log; tail -5 $O/s7_pytest.log
python - <<P
import json
try:
    d=json.loads(open("$O/s7_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()})
    print("e2e", d["e2e"]); print("bits", d.get("bits_per_base")); print("cpu", d.get("cpu_baseline")); print("roof", d["roofline"]["frac"], d["allocator"], d.get("ingest"))
except Exception as e:
    print("ERR", e); print(open("$O/s7_bench.err").read()[-2500:])
print(open("$O/s7_ref.json").read()[:600])
P
