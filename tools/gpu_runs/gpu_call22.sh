#!/bin/bash
# twelve epilogue warps for the one-CTA-per-SM pair variants (HFG_TC_PAIR_EW=12): bit identity, per-layer A/B, stage times
out=gpurun_out/c22
mkdir -p $out
export HFG_LIB_PATH=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for m in tf32 bf16 fp16; do for g in 8 12; do echo -n "$m ew=$g "; HFG_TC_PAIR_EW=$g timeout 200 python tools/variant_hash.py $m 2>&1 | grep -E "HASH|rror" | tail -1; done; done
for m in fp16 tf32; do for g in 8 12; do
  HFG_TC_PAIR_EW=$g timeout 300 python tools/tune_layers.py --which 2 --mode $m --stages 0,1 --pairs 0 > $out/layers_${m}_ew$g.txt 2>&1
done; paste <(cut -c1-75 $out/layers_${m}_ew8.txt) <(cut -c46-80 $out/layers_${m}_ew12.txt); done
for m in tf32 bf16; do for g in 8 12; do HFG_TC_PAIR_EW=$g timeout 200 python tools/stage_times.py $m > $out/stages_${m}_ew$g.txt 2>&1; done; paste $out/stages_${m}_ew8.txt $out/stages_${m}_ew12.txt; done
