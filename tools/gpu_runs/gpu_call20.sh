#!/bin/bash
out=gpurun_out/c20
mkdir -p $out
export HFG_LIB_PATH=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for m in tf32 bf16; do for g in 0 1; do echo -n "$m contig=$g "; HFG_TC_UP_CONTIG=$g timeout 200 python tools/variant_hash.py $m 2>&1 | grep -E "HASH|rror" | tail -1; done; done
for m in tf32 bf16; do for g in 0 1 0 1; do HFG_TC_UP_CONTIG=$g timeout 200 python tools/stage_times.py $m > $out/stages_${m}_c$g.txt 2>&1; grep -E "ups|total" $out/stages_${m}_c$g.txt | tr '\n' ' ' | sed "s/^/$m contig=$g /"; echo; done; done
