#!/bin/bash
# r02 profiling recipe (B200_PROFILING.md): launch lists (device time of every launch) and one --set full capture per hot
# kernel.  Every ncu run follows a plain run of the same command that exited 0.  Two parts (gpurun brings back <= 64 MiB):
#   bash tools/profile_r02.sh a   -> launch lists (1-D bench, 2-D bench); accum_1d, elbo_chains, predict_1d
#   bash tools/profile_r02.sh b   -> accum_2d_cols, nd_factor (leaf level and a middle level), nd_selinv, predict_2d_cols
#   bash tools/profile_r02.sh c   -> only the two streaming 2-D kernels (accum_2d_cols, predict_2d_cols)
set -x
B1="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-2d"
B2="python bench.py --workload 2d --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
full() { ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3 -f ${@:4} > gpurun_out/ncu_$3.log 2>&1; }
if [ "$1" = "a" ]; then
  $B1 > gpurun_out/r2_plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_1d.csv $B1 > gpurun_out/ncu_l1.log 2>&1
  $B2 > gpurun_out/r2_plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches_2d.csv $B2 > gpurun_out/ncu_l2.log 2>&1
  full accum_1d_kernel 3 accum_1d $B1
  full elbo_chains_cluster_kernel 6 chain_kuu_1d $B1        # two launches per bound: the Kuu chain (side stream; even), the P chains (odd)
  full elbo_chains_cluster_kernel 7 chains_p_1d $B1
  full predict_1d_kernel 1 predict_1d $B1
elif [ "$1" = "c" ]; then
  $B2 > gpurun_out/r2_plain2.log 2>&1 || exit 1
  full accum_2d_cols_kernel 3 accum_2d_cols $B2
  full predict_2d_cols_kernel 1 predict_2d_cols $B2
else
  $B2 > gpurun_out/r2_plain2.log 2>&1 || exit 1
  full accum_2d_cols_kernel 3 accum_2d_cols $B2
  full nd_factor_kernel 27 nd_factor_leaves $B2          # 9 levels per factorisation: launch 27 = the leaf level of the 4th
  full nd_factor_kernel 31 nd_factor_level4 $B2          # 16 fronts of 13 x 13 tiles
  full nd_factor_kernel 35 nd_factor_root $B2            # one 10 x 10 front: the pure chain
  full nd_selinv_kernel 28 nd_selinv_level1 $B2
  full predict_2d_cols_kernel 1 predict_2d_cols $B2
fi
# captures over 40 MB (the chain kernels: 190 KB of SASS with per-instruction counters) stay on the box: raw page exported as CSV
for f in gpurun_out/prof_*.ncu-rep; do
  if [ $(stat -c %s $f) -gt 40000000 ]; then
    ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null
    ncu -i $f --page source --csv > ${f%.ncu-rep}.source.csv 2>/dev/null
    rm -f $f
  fi
done
ls -la gpurun_out/
