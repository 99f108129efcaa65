#!/bin/bash
# r01 profiling recipe (B200_PROFILING.md): launch lists (device time of every launch) and one --set full capture per hot
# kernel.  Every ncu run follows a plain run of the same command that exited 0.  Two parts (gpurun brings back <= 64 MiB):
#   bash tools/profile_r01.sh a   -> launch lists, accum_1d, elbo_chains
#   bash tools/profile_r01.sh b   -> accum_2d_cols, td_factor, td_selinv, predict_2d_cols
#   bash tools/profile_r01.sh c   -> inputs in no particular order: launch lists + partition / unit kernels; predict_1d
set -x
B1="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-2d"
B2="python bench.py --workload 2d --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
full() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3 -f ${@:4} > gpurun_out/ncu_$3.log 2>&1; }
if [ "$1" = "a" ]; then
  $B1 > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_1d.csv $B1 > gpurun_out/ncu_l1.log 2>&1
  $B2 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_2d.csv $B2 > gpurun_out/ncu_l2.log 2>&1
  full accum_1d_kernel 3 accum_1d $B1
  full elbo_chains_kernel 3 chains_1d $B1
elif [ "$1" = "c" ]; then
  R1="python tools/binned_1d_only.py"
  R2="python tools/binned_2d_only.py"
  fullns() { ncu --set full --clock-control none -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3 -f ${@:4} > gpurun_out/ncu_$3.log 2>&1; }
  $R1 > gpurun_out/plain_r1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_binned_1d.csv $R1 > gpurun_out/ncu_lr1.log 2>&1
  $R2 > gpurun_out/plain_r2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_binned_2d.csv $R2 > gpurun_out/ncu_lr2.log 2>&1
  fullns accum_1d_units_kernel 1 accum_1d_units $R1
  fullns accum_2d_units_kernel 1 accum_2d_units $R2
  $B1 > gpurun_out/plain1.log 2>&1 && fullns predict_1d_kernel 1 predict_1d $B1
else
  $B2 > gpurun_out/plain2.log 2>&1 || exit 1
  full accum_2d_cols_kernel 3 accum_2d_cols $B2
  full td_factor_kernel 3 td_factor $B2
  full td_selinv_kernel 3 td_selinv $B2
  full predict_2d_cols_kernel 1 predict_2d_cols $B2
fi
ls -la gpurun_out/*.ncu-rep
