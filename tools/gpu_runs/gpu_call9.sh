#!/bin/bash
out=gpurun_out/c9
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/rc.txt; tail -4 $out/smoke.log
HFG_LIB_PATH=$T HFG_TC_VERBOSE=1 timeout 200 python tools/variant_hash.py bf16 2>&1 | grep -E "pair\] N=(32|64) " | sort | uniq -c | head -12
timeout 1500 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or saturated or stage_bound or config2 or ragged_batch_against or chunking or concurrent or independence" > $out/pytest_par.log 2>&1; echo "pytest_par rc=$?" | tee -a $out/rc.txt
tail -12 $out/pytest_par.log
for v in 1 0; do for m in bf16 tf32; do HFG_LIB_PATH=$T HFG_TC_S2D=$v timeout 200 python tools/stage_times.py $m > $out/stages_${m}_s2d$v.txt 2>&1; done; done
for m in bf16 tf32; do echo "== $m: s2d 1 / 0"; paste $out/stages_${m}_s2d1.txt $out/stages_${m}_s2d0.txt; done
cp gpurun_out/parity_r2.jsonl $out/ 2>/dev/null
