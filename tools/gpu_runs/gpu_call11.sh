#!/bin/bash
out=gpurun_out/c11
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
timeout 900 python -m pytest tests -m gpu -q -k "saturates or c_abi_errors or ar_decoder or variants" > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/rc.txt
tail -6 $out/pytest.log
# C = 64: big tiles, one CTA per SM, with / without the space-to-depth conv2
for v in 0 2; do HFG_LIB_PATH=$T HFG_TC_S2D=$v HFG_TC_PAIR_MT=4 HFG_TC_PAIR_OCC2=0 timeout 300 python tools/tune_layers.py --which 2 --stages 2 --mode bf16 > $out/layers_c64_mt4_s2d$v.txt 2>&1; done
paste $out/layers_c64_mt4_s2d0.txt $out/layers_c64_mt4_s2d2.txt | cut -c1-220
for v in 0 2; do HFG_LIB_PATH=$T HFG_TC_S2D=$v HFG_TC_PAIR_MT=4 HFG_TC_PAIR_OCC2=0 HFG_TC_PAIR_CTAS=1 timeout 300 python tools/tune_layers.py --which 2 --stages 2 --mode bf16 > $out/layers_c64_mt4_ctas1_s2d$v.txt 2>&1; done
paste $out/layers_c64_mt4_ctas1_s2d0.txt $out/layers_c64_mt4_ctas1_s2d2.txt | cut -c1-220
