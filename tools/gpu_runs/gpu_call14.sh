#!/bin/bash
# split tf32 plan (fp16 hi + lo planes): parity, A/B against the fp32-plane path, stage times
out=gpurun_out/c14
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -x -k "tf32" > $out/pytest_tf32.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
grep -E "passed|failed|max-abs|Error|assert" $out/pytest_tf32.log | tail -40
for v in 1 0 1 0; do HFG_LIB_PATH=$T HFG_TC_TF32_MIXED=$v timeout 200 python tools/stage_times.py tf32 > $out/stages_tf32_mixed$v.txt 2>&1; grep total $out/stages_tf32_mixed$v.txt | sed "s/^/tf32 mixed=$v /"; done
paste $out/stages_tf32_mixed0.txt $out/stages_tf32_mixed1.txt
HFG_LIB_PATH=$T HFG_TC_VERBOSE=1 timeout 200 python tools/stage_times.py tf32 2>&1 | grep "^\[" | sort | uniq -c | sort -rn | head -40 > $out/geom.txt; cat $out/geom.txt
