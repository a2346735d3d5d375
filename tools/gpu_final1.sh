#!/bin/bash
# one GPU, final build: whole GPU test suite, the default bench line, (optionally) the ncu capture of the walk
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/f1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/f1_pytest.log; tail -4 $O/f1_pytest.log
timeout 1200 python bench.py --steps 20 --warmup 5 > $O/f1_bench.json 2> $O/f1_bench.err; echo "bench rc=$?"
python - <<P
import json
try:
    d=json.loads(open("$O/f1_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["phases_ms"].items()}, d["allocator"], d["roofline"]["frac"], d["roofline"]["traffic"])
    e=d["e2e"]; print("  e2e", round(e["value"],1), round(e["ms_per_step"],1), e["what"], "single", round(e["single_job"]["value"],1), {k:round(v,1) for k,v in e["single_job"]["host_wall_ms"].items()})
    print("  bits", d.get("bits_per_base",{}).get("gpu_over_ref_t1"), d.get("bits_per_base",{}).get("gpu_over_ref_tN"), "cpu", d.get("cpu_baseline"))
except Exception as e:
    print("ERR", e); print(open("$O/f1_bench.err").read()[-1500:])
P
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -c 1 -o $O/f1_walk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --ingest-reads 0 > $O/f1_w.log 2>&1; echo "walk capture rc=$?"
fi
