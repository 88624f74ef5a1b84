#!/bin/bash
# final single-GPU evidence run of round 2: full GPU test suite, default bench in every mode, CPU arm, ncu of the S2D pair
out=gpurun_out/f1
mkdir -p $out gpurun_out/ncu_r2c
timeout 2400 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -6 $out/pytest.log
cp gpurun_out/parity_r2.jsonl $out/parity_r2.jsonl 2>/dev/null
cp gpurun_out/ragged_r2.json $out/ragged_r2.json 2>/dev/null
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
o=gpurun_out/ncu_r2c
python tools/profile_step.py --mode bf16 > $o/plain_bf16.log 2>&1 && {
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 44 -c 44 --csv --log-file $o/forward_bf16.csv python tools/profile_step.py --mode bf16 > $o/ncu_fwd_bf16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_pair_kernel -s 67 -c 1 -o $o/pair_mrf3k11_bf16_s2d python tools/profile_step.py --mode bf16 > $o/cap.log 2>&1
ncu -i $o/pair_mrf3k11_bf16_s2d.ncu-rep --page raw --csv > $o/pair_mrf3k11_bf16_s2d_raw.csv 2>/dev/null
}
python tools/profile_step.py --mode tf32 > $o/plain_tf32.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 44 -c 44 --csv --log-file $o/forward_tf32.csv python tools/profile_step.py --mode tf32 > $o/ncu_fwd_tf32.log 2>&1
ls $o
