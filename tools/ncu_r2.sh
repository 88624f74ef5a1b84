#!/bin/bash
# ncu evidence for round 2 (one GPU; run only after the same commands exited 0 without ncu).
#   launch lists of one short bench-shaped run per mode, and --set full captures of the kernels VERDICT names.
out=gpurun_out/ncu_r2
mkdir -p $out
for m in bf16 tf32; do
  python tools/profile_step.py --mode $m > $out/plain_$m.log 2>&1 || { echo "plain run failed for $m"; exit 1; }
done
# every launch of the second forward (44 launches bf16 / tf32 now fused: 44): light metric set
for m in bf16 tf32; do
  n=$(grep -o "[0-9]*)\?$" $out/plain_$m.log | tail -1)
  ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
      --clock-control none -s ${n:-44} -c ${n:-44} --csv --log-file $out/forward_$m.csv python tools/profile_step.py --mode $m > $out/ncu_fwd_$m.log 2>&1
done
cap() {  # name mode regex skip
  ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o $out/$1 python tools/profile_step.py --mode $2 > $out/$1.log 2>&1
  ncu -i $out/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
}
cap post_bf16 bf16 tc_conv_post 1
cap up_ups1_bf16 bf16 tc_up_kernel 5
cap up_ups3_bf16 bf16 tc_up_kernel 7
cap pair_mrf3k3_bf16 bf16 tc_pair_kernel 63
cap pair_mrf3k11_bf16 bf16 tc_pair_kernel 67
cap pair_mrf1k11_bf16 bf16 tc_pair_kernel 49
cap pair_mrf0k11_tf32 tf32 tc_pair_kernel 40
cap pair_mrf1k11_tf32 tf32 tc_pair_kernel 49
ls -la $out
