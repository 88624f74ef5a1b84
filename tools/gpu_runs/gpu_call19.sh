#!/bin/bash
# tile as two independent halves (HFG_TC_PAIR_GROUPS): bit identity, per-layer A/B, stage times
out=gpurun_out/c19
mkdir -p $out
export HFG_LIB_PATH=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for m in tf32 bf16; do for g in 0 1 2; do echo -n "$m groups=$g "; HFG_TC_PAIR_GROUPS=$g timeout 200 python tools/variant_hash.py $m 2>&1 | grep -E "HASH|rror" | tail -1; done; done
for m in fp16 tf32; do for g in 0 2; do
  HFG_TC_PAIR_GROUPS=$g timeout 300 python tools/tune_layers.py --which 2 --mode $m --pairs 0 > $out/layers_${m}_g$g.txt 2>&1
done; paste <(cut -c1-75 $out/layers_${m}_g0.txt) <(cut -c46-80 $out/layers_${m}_g2.txt); done
for m in tf32 bf16; do for g in 0 1; do HFG_TC_PAIR_GROUPS=$g timeout 200 python tools/stage_times.py $m > $out/stages_${m}_g$g.txt 2>&1; done; paste $out/stages_${m}_g0.txt $out/stages_${m}_g1.txt; done
