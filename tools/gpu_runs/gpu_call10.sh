#!/bin/bash
out=gpurun_out/c10
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for v in 0 1; do
  HFG_LIB_PATH=$T HFG_TC_S2D=$v timeout 300 python tools/tune_layers.py --which 2 --stages 2,3 --mode bf16 > $out/layers_bf16_s2d$v.txt 2>&1
done
paste $out/layers_bf16_s2d0.txt $out/layers_bf16_s2d1.txt | cut -c1-220
HFG_LIB_PATH=$T HFG_TC_S2D=1 HFG_TC_PAIR_CTAS=1 timeout 300 python tools/tune_layers.py --which 2 --stages 2 --mode bf16 > $out/layers_bf16_s2d1_ctas1.txt 2>&1
HFG_LIB_PATH=$T HFG_TC_S2D=1 HFG_TC_PAIR_OCC2=0 timeout 300 python tools/tune_layers.py --which 2 --stages 2 --mode bf16 > $out/layers_bf16_s2d1_occ1.txt 2>&1
echo "== stage 2 s2d: ctas=1 | occ2=0"; paste $out/layers_bf16_s2d1_ctas1.txt $out/layers_bf16_s2d1_occ1.txt | cut -c1-220
