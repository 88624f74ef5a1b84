#!/bin/bash
out=gpurun_out/f4
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee $out/rc.txt
grep smoke $out/smoke.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "c_abi_error or (golden and tf32) or fp32_plane or unusual" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -3 $out/pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > $out/bench_quick.json 2> $out/bench_quick.err; echo "bench rc=$?" | tee -a $out/rc.txt
python -c "
import json; d=json.loads(open('$out/bench_quick.json').read().strip().splitlines()[-1]); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
