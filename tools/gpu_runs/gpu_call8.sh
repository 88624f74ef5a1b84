#!/bin/bash
out=gpurun_out/c8
mkdir -p $out gpurun_out/ncu_r2b
timeout 600 python -m pytest tests/test_ar_decoder.py -m gpu -q -s > $out/pytest_ard.log 2>&1; echo "pytest_ard rc=$?" | tee -a $out/rc.txt
grep -E "passed|failed|decoder|Error|assert" $out/pytest_ard.log | head
o=gpurun_out/ncu_r2b
for m in bf16 tf32; do
  python tools/profile_step.py --mode $m > $o/plain_$m.log 2>&1 || { echo "plain run failed for $m"; exit 1; }
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -s 44 -c 44 --csv --log-file $o/forward_$m.csv python tools/profile_step.py --mode $m > $o/ncu_fwd_$m.log 2>&1
done
cap() { ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o $o/$1 python tools/profile_step.py --mode $2 > $o/$1.log 2>&1; ncu -i $o/$1.ncu-rep --page raw --csv > $o/$1_raw.csv 2>/dev/null; }
cap post_bf16 bf16 tc_conv_post 1
cap post_tf32 tf32 tc_conv_post 1
cap pair_mrf0k11_tf32 tf32 tc_pair_kernel 40
ls $o | head -30
