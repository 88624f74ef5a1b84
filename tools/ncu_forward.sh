#!/bin/bash
# usage: tools/ncu_forward.sh MODE LAUNCHES  -- per-kernel ncu metrics of ONE forward (the second of two) of the bench workload
mode=$1; n=$2
python tools/profile_step.py --mode $mode > /dev/null || exit 1      # must pass plain first
ncu --clock-control none -s $n -c $n --csv --log-file gpurun_out/sol_$mode.csv \
    --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread \
    python tools/profile_step.py --mode $mode > gpurun_out/ncu_sol_$mode.log 2>&1
