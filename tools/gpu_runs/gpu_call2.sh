#!/bin/bash
out=gpurun_out/c2
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
timeout 1800 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -15 $out/pytest.log
# tf32: fp16 intermediate on / off (tuning library), per-stage times
for v in 1 0; do HFG_LIB_PATH=$T HFG_TC_TF32_H16=$v HFG_TC_VERBOSE=0 timeout 200 python tools/stage_times.py tf32 > $out/stages_tf32_h16_$v.txt 2>&1; done
paste $out/stages_tf32_h16_1.txt $out/stages_tf32_h16_0.txt
HFG_LIB_PATH=$T HFG_TC_SUM_PLANES=0 timeout 200 python tools/stage_times.py bf16 > $out/stages_bf16_acc.txt 2>&1
timeout 200 python tools/stage_times.py bf16 > $out/stages_bf16.txt 2>&1
paste $out/stages_bf16.txt $out/stages_bf16_acc.txt
timeout 900 python bench.py --steps 30 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/rc.txt
tail -3 $out/bench.err
bash tools/ncu_r2.sh > $out/ncu.log 2>&1; echo "ncu rc=$?" | tee -a $out/rc.txt
tail -20 $out/ncu.log
