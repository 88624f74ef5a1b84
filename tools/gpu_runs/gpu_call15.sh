#!/bin/bash
# per-layer A/B of the fused pair: fp16 mode vs the hi + lo variant of the tf32 mode; phase timelines
out=gpurun_out/c15
mkdir -p $out
export HFG_LIB_PATH=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for m in fp16 tf32; do timeout 300 python tools/tune_layers.py --which 2 --mode $m --pairs 0 > $out/layers_$m.txt 2>&1; done
paste <(cut -c1-70 $out/layers_fp16.txt) <(cut -c46-80 $out/layers_tf32.txt)
for m in fp16 tf32; do
  rm -f $out/tl_$m.txt
  HFG_TC_TIMELINE=$out/tl_$m.txt timeout 300 python tools/tune_layers.py --which 2 --mode $m --stages 1,3 --resblocks 0,2 --pairs 0 > /dev/null 2>&1
  python tools/pair_timeline.py $out/tl_$m.txt > $out/timeline_$m.txt 2>&1
  grep -E "^stage|steady" $out/timeline_$m.txt | sed "s/^/$m /"
done
