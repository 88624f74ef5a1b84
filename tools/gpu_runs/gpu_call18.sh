#!/bin/bash
# FHADD hi/lo join + split, 32-byte paired-row stores in the space-to-depth epilogue: parity subset + stage times
out=gpurun_out/c18
mkdir -p $out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "golden or config2 or variants or ragged_batch_against or chunking or fp32_plane" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/rc.txt
tail -3 $out/pytest.log
for m in tf32 bf16 fp16; do timeout 200 python tools/stage_times.py $m > $out/stages_$m.txt 2>&1; done
paste $out/stages_tf32.txt $out/stages_bf16.txt $out/stages_fp16.txt
export HFG_LIB_PATH=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for m in fp16 tf32; do timeout 300 python tools/tune_layers.py --which 2 --mode $m --stages 1,3 --pairs 0 > $out/layers_$m.txt 2>&1; done
paste <(cut -c1-75 $out/layers_fp16.txt) <(cut -c46-80 $out/layers_tf32.txt)
