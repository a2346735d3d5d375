#!/bin/bash
# one GPU: ncu --set full of the largest kernels behind the walk (pool probe, radix scatter, bit-sliced consensus)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pool_probe_kernel|consensus_sliced_kernel" -s 2 -c 2 -o $O/f3_s2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --ingest-reads 0 > $O/f3_a.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_scatter_kernel -s 25 -c 4 -o $O/f3_rs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --ingest-reads 0 > $O/f3_b.log 2>&1; echo "rc=$?"
ls -la $O/f3_*.ncu-rep
