#!/bin/bash
# One gpurun call: smoke, GPU tests, short bench, per-stage times.  Logs under gpurun_out/<tag>/.
tag=${1:-call}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/smi.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt
timeout 1500 python -m pytest tests -m gpu -q -k "${PYTEST_K:-not variants}" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -40 $out/pytest.log
timeout 900 python bench.py --steps ${STEPS:-30} --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
tail -5 $out/bench.err
for m in tf32 fp16 bf16; do timeout 200 python tools/stage_times.py $m > $out/stages_$m.txt 2>&1; done
cat $out/stages_bf16.txt
