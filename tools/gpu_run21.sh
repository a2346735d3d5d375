#!/bin/bash
# one GPU: repeat-rich genome against a uniform one of the same size (robustness of the walk on big bins), then the region
# profile of the walk (-DWALK_PROF build, made on the box only)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python tools/run_shape.py 3e6 100 repeats:20 1 1 > $O/s21_rep.txt 2> $O/s21_rep.err; echo "rep rc=$?"; cat $O/s21_rep.txt
timeout 600 python tools/run_shape.py 3e6 100 8638673 1 1 > $O/s21_uni.txt 2> $O/s21_uni.err; echo "uni rc=$?"; cat $O/s21_uni.txt
timeout 600 python tools/run_shape.py 2e7 100 repeats:100 1 1 > $O/s21_rep100.txt 2> $O/s21_rep100.err; echo "rep100 rc=$?"; cat $O/s21_rep100.txt
HARC_CUFLAGS=-DWALK_PROF python harc_b200/build.py -f > /dev/null 2>&1
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --ingest-reads 0 > $O/s21_prof.json 2> $O/s21_prof.err; echo "prof rc=$?"
grep WALK_PROF $O/s21_prof.err | tail -2
