#!/bin/bash
out=gpurun_out/c12
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for v in 0 1; do HFG_LIB_PATH=$T HFG_TC_PDL=$v timeout 200 python tools/variant_hash.py bf16 2>&1 | grep HASH; done
for v in 0 1; do HFG_LIB_PATH=$T HFG_TC_PDL=$v timeout 200 python tools/variant_hash.py tf32 2>&1 | grep HASH; done
for v in 0 1 0 1; do for m in bf16 tf32; do HFG_LIB_PATH=$T HFG_TC_PDL=$v timeout 200 python tools/stage_times.py $m > $out/stages_${m}_pdl$v.txt 2>&1; grep total $out/stages_${m}_pdl$v.txt | sed "s/^/$m pdl=$v /"; done; done
paste $out/stages_bf16_pdl0.txt $out/stages_bf16_pdl1.txt
# host-buffer path (graph capture with programmatic edges) must still work and match
HFG_LIB_PATH=$T HFG_TC_PDL=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "host_path or generate_stream or host_buffer or chunking or ragged_batch_against" > $out/pytest_pdl.log 2>&1; echo "pytest_pdl rc=$?"; tail -3 $out/pytest_pdl.log
