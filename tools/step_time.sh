#!/bin/bash
# usage: tools/step_time.sh "<ENV=V ...>" [modes] [repeats]  -- device-timed ms/step of the bench workload, several runs
envs="$1"; modes="${2:-tf32 bf16}"; reps="${3:-3}"
for m in $modes; do
  out=""
  for r in $(seq $reps); do
    v=$(env $envs python bench.py --mode $m --steps 30 --warmup 5 --no-cpu-baseline --no-quality 2>/dev/null | python -c "import json,sys; print('%.3f' % json.loads(sys.stdin.read())['ms_per_step'])")
    out="$out $v"
  done
  echo "$envs $m:$out"
done
