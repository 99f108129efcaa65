set -x
B1="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-2d"
B2="python bench.py --workload 2d --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$B1 > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_1d.csv $B1 > gpurun_out/ncu_l1.log 2>&1
$B2 > gpurun_out/plain2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/launches_2d.csv $B2 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:accum_1d_kernel -s 3 -c 1 -o gpurun_out/prof_accum_1d -f $B1 > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elbo_chains_kernel -s 3 -c 1 -o gpurun_out/prof_chains_1d -f $B1 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:accum_2d_kernel -s 3 -c 1 -o gpurun_out/prof_accum_2d -f $B2 > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bb_syrk_kernel -s 300 -c 1 -o gpurun_out/prof_bb_syrk -f $B2 > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bb_potrf_kernel -s 300 -c 1 -o gpurun_out/prof_bb_potrf -f $B2 > gpurun_out/ncu_f5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bb_sel_symm_kernel -s 300 -c 1 -o gpurun_out/prof_bb_sel_symm -f $B2 > gpurun_out/ncu_f6.log 2>&1
ls -la gpurun_out/*.ncu-rep
