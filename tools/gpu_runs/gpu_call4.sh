#!/bin/bash
out=gpurun_out/c4
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
timeout 600 python -m pytest tests/test_ar_decoder.py -m gpu -q -s > $out/pytest_ard.log 2>&1; echo "pytest_ard rc=$?" | tee -a $out/rc.txt
grep -E "passed|failed|decoder|Error" $out/pytest_ard.log | head
for ns in 0 64 256 1024; do
  for m in bf16 tf32; do
    HFG_LIB_PATH=$T HFG_TC_EPI_SLEEP_NS=$ns timeout 200 python tools/stage_times.py $m > $out/stages_${m}_sleep$ns.txt 2>&1
  done
done
for m in bf16 tf32; do echo "== $m: sleep 0 / 64 / 256 / 1024"; paste $out/stages_${m}_sleep0.txt $out/stages_${m}_sleep64.txt $out/stages_${m}_sleep256.txt $out/stages_${m}_sleep1024.txt | cut -c1-200; done
# batch 32 per GPU (config 3 at 8 GPUs)
for ns in 0 256; do HFG_LIB_PATH=$T HFG_TC_EPI_SLEEP_NS=$ns timeout 200 python tools/stage_times.py bf16 32 > $out/stages_bf16_b32_sleep$ns.txt 2>&1; done
paste $out/stages_bf16_b32_sleep0.txt $out/stages_bf16_b32_sleep256.txt
