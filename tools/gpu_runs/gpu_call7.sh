#!/bin/bash
out=gpurun_out/c7
mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -8 $out/pytest.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
tail -3 $out/bench.err
timeout 600 python bench.py --mode bf16 --steps 50 --warmup 10 --no-configs > $out/bench_bf16.json 2> $out/bench_bf16.err; echo "bench_bf16 rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --mode fp16 --steps 50 --warmup 10 --no-configs > $out/bench_fp16.json 2> $out/bench_fp16.err; echo "bench_fp16 rc=$?" | tee -a $out/rc.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; echo "bench_ref rc=$?" | tee -a $out/rc.txt
python - <<PY
import json
for f in ("bench","bench_bf16","bench_fp16","bench_ref"):
    try:
        d=json.loads(open("$out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value",round(d["value"],1),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1), "sync", round(d.get("e2e_synchronous",{}).get("value",0),1), "frac", d.get("roofline",{}).get("frac"))
    except Exception as e: print(f,"parse failed",e)
PY
