#!/bin/bash
# N-GPU bench line only (final code)
N=${1:-2}
out=gpurun_out/c23_n$N
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee $out/rc.txt
python - <<PY
import json
try:
    d=json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
    print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"n",d["n_gpus"])
    c3=d["config3"]; print("config3",{m:(c3[m]["ms_per_step"],c3[m]["value"],c3[m]["e2e"]["value"]) for m in ("bf16","fp16")})
    c4=d["config4"]; print("config4",{m:(v["ms_per_step"],v["value"],v["chunks_equal_unchunked"]) for m,v in c4.items() if isinstance(v,dict)})
except Exception as e: print("parse failed",e)
PY
