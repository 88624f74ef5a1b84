#!/bin/bash
# usage: tools/scale_run.sh N   -- N-GPU weak-scaling lines: tf32 default (16 utt/GPU) and bf16 with 32 utt/GPU (BASELINE configs[2] at N=8)
N=${1:-8}
i=0
for args in "--mode tf32" "--mode bf16 --batch 32"; do
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+i)) \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-quality $args 2>&1 | tail -1 > gpurun_out/scale${N}_$i.json
done
cat gpurun_out/scale${N}_*.json | python -c '
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d["dtype"], d["n_gpus"], round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), d["config"]["workload"], d["clocks"])'
