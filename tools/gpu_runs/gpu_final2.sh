#!/bin/bash
# final single-GPU evidence of round 2 (split tf32 plan): default bench in every mode, CPU arm, stage times, ncu lists + captures
out=gpurun_out/f2
mkdir -p $out
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --mode bf16 --steps 50 --warmup 10 --no-configs > $out/bench_bf16.json 2> $out/bench_bf16.err; echo "bench_bf16 rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --mode fp16 --steps 50 --warmup 10 --no-configs > $out/bench_fp16.json 2> $out/bench_fp16.err; echo "bench_fp16 rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench_ref rc=$?" | tee -a $out/rc.txt
for m in tf32 fp16 bf16; do timeout 200 python tools/stage_times.py $m > $out/stages_$m.txt 2>&1; done
paste $out/stages_tf32.txt $out/stages_bf16.txt
python - <<PY
import json
for f in ("bench","bench_bf16","bench_fp16","bench_ref"):
    try:
        d=json.loads(open("$out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1), "sync", round(d.get("e2e_synchronous",{}).get("value",0),1), "frac", d.get("roofline",{}).get("frac"))
    except Exception as e: print(f,"parse failed",e)
PY
o=$out/ncu
mkdir -p $o
for m in tf32 bf16; do
  python tools/profile_step.py --mode $m > $o/plain_$m.log 2>&1 || { echo "plain $m failed"; continue; }
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -s 44 -c 44 --csv --log-file $o/forward_$m.csv python tools/profile_step.py --mode $m > $o/ncu_fwd_$m.log 2>&1
done
cap() {  # name mode regex skip
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o $o/$1 python tools/profile_step.py --mode $2 > $o/$1.log 2>&1
  ncu -i $o/$1.ncu-rep --page raw --csv > $o/$1_raw.csv 2>/dev/null
}
cap pair_mrf1k11_tf32split tf32 tc_pair_kernel 49
cap pair_mrf3k11_tf32split tf32 tc_pair_kernel 67
cap up_ups1_tf32split tf32 tc_up_kernel 5
ls $o
