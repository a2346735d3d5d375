#!/bin/bash
# one GPU: ncu --set full of the 250-bp walk kernel, walk_kernel<8, 32> (configs[3] shape at a fifth of its size)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o $O/f4_walk8 -f python tools/run_shape.py 2e7 250 2e8 0 1 > $O/f4.log 2>&1; echo "rc=$?"
tail -2 $O/f4.log | cut -c1-700
