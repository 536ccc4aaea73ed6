#!/bin/bash
# round-2 ncu evidence of the final build (run under gpurun on one B200): launch list of the bench command + one --set full
# capture per kernel class of the dense pass, the gather and the top training kernels.  usage: scripts/profile_r2.sh OUTDIR
OUT=${1:-gpurun_out}
mkdir -p $OUT
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 2000 --csv --log-file $OUT/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > $OUT/r2_bench_under_ncu.json 2> /dev/null
$NCU --set full --import-source on -k regex:gemm_tc_pair -c 5 -f -o $OUT/r2_prof_fc python scripts/one_pass.py 256 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:gemm_tc_persistent -c 1 -f -o $OUT/r2_prof_out python scripts/one_pass.py 256 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"conv_sweep|conv1_wide" -c 5 -f -o $OUT/r2_prof_sweep python scripts/one_pass.py 256 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:gather_patches -c 1 -f -o $OUT/r2_prof_gather python scripts/exp_gather.py base: > /dev/null 2>&1
$NCU --set full --import-source on -k regex:"wgrad_tc|tbn_bwd_dx|tbn_act" -s 6 -c 6 -f -o $OUT/r2_prof_train python scripts/one_train_step.py 1024 1 > /dev/null 2>&1
ls -la $OUT/r2_prof_*.ncu-rep $OUT/r2_launches.csv
