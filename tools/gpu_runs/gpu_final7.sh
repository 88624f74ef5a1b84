#!/bin/bash
# last GPU call of round 2: smoke() and the GPU suite minus its three long cases, on the final tree, with per-test durations
out=gpurun_out/f7
mkdir -p $out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee $out/rc.txt
grep smoke $out/smoke.log
timeout 160 python -m pytest tests -m gpu -x -q --durations=25 -p no:cacheprovider \
    -k "not variants and not config3_full and not long_form and not arithmetic_model and not unusual" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -32 $out/pytest.log
