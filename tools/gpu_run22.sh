#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
HARCGPU_LAPS=1 timeout 300 python tools/run_shape.py 3e6 100 repeats:20 1 1 > $O/s22_rep.txt 2> $O/s22_rep.err; echo "rep rc=$?"; cat $O/s22_rep.txt
