#!/bin/bash
# (1) N = 128 pairs with two CTAs per SM (HFG_TC_PAIR_OCC2=1) vs the default, fp16 and the tf32 hi + lo variant
# (2) ncu captures of the hi + lo pair kernel (source view) -- after the plain run exited 0
out=gpurun_out/c16
mkdir -p $out
T=tts-sambert_hifigan_b200/lib/libhfg_b200_tuning.so
for m in fp16 tf32; do for o in 0 1; do
  HFG_LIB_PATH=$T HFG_TC_PAIR_OCC2=$o timeout 300 python tools/tune_layers.py --which 2 --mode $m --stages 1 --pairs 0,2 > $out/occ${o}_$m.txt 2>&1
done; paste <(cut -c1-75 $out/occ0_$m.txt) <(cut -c46-80 $out/occ1_$m.txt); done
python tools/profile_step.py --mode tf32 > $out/plain_tf32.log 2>&1 || { echo "plain run failed"; tail -5 $out/plain_tf32.log; exit 1; }
cap() {  # name mode regex skip
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o $out/$1 python tools/profile_step.py --mode $2 > $out/$1.log 2>&1
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
}
cap pair_mrf1k3_tf32lo tf32 tc_pair_kernel 45
cap pair_mrf1k11_tf32lo tf32 tc_pair_kernel 49
cap pair_mrf3k11_tf32lo tf32 tc_pair_kernel 67
ls -la $out
