#!/bin/bash
# usage: tools/bench_stages.sh "<ENV=V ...>" [modes]   -- per-stage MRF times of the bench workload under env overrides
envs="$1"; modes="${2:-tf32 bf16}"
for m in $modes; do
  env $envs python bench.py --mode $m --steps 10 --warmup 3 --no-cpu-baseline --no-quality | python -c "
import json,sys
d=json.loads(sys.stdin.read())
ps={p['kernel']:p['ms'] for p in d['roofline']['per_stage']}
st=lambda i: sum(v for k,v in ps.items() if k.startswith('mrf%d'%i))
print('$envs', d['dtype'], 'ms/step %.3f'%d['ms_per_step'], ' '.join('mrf%d %.3f'%(i,st(i)) for i in range(4)), ' '.join('%s %.3f'%(k,v) for k,v in ps.items() if k.startswith('ups') or k.startswith('mrf3') or k.startswith('mrf2')))
"
done
