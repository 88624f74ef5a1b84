#!/bin/bash
# ncu launch list of bench.py itself on the final code (durations only; one pass, no replay)
out=gpurun_out/f9
mkdir -p $out
A="--steps 2 --warmup 3 --no-configs --no-cpu-baseline --no-quality"
timeout 16 python bench.py $A > $out/bench_plain.json 2> $out/bench_plain.err; rc=$?; echo "plain rc=$rc" | tee $out/rc.txt
[ $rc -eq 0 ] && { timeout 22 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_tf32.csv \
    python bench.py $A > $out/bench_under_ncu.json 2> $out/bench_under_ncu.err; echo "ncu rc=$?" | tee -a $out/rc.txt; }
wc -l $out/*.csv
