#!/bin/bash
# Round-2 ncu evidence (one GPU): launch lists of the bench's legs and full captures of the dominant kernels.
# Every program is first run WITHOUT ncu (exit 0 required); numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
NCU="ncu --clock-control none"
python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > $OUT/r2_ncu_plain_bench.json 2> $OUT/r2_ncu_plain_bench.log || exit 1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > $OUT/r2_ncu_list.log 2>&1
$NCU --set full --import-source on -k regex:hmc_random_tc_kernel -s 1 -c 1 -f -o $OUT/r2_tc_full python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > $OUT/r2_ncu_tc.log 2>&1
python profiles/run_diag_once.py 16384 1000 > $OUT/r2_diag_plain.txt 2>&1 || exit 1
$NCU --set full --import-source on -k regex:"diag_variogram_f32x2|diag_moments_f32x4" -s 4 -c 3 -f -o $OUT/r2_diag_full python profiles/run_diag_once.py 16384 1000 > $OUT/r2_ncu_diag.log 2>&1
BIGD_NITER=1 BIGD_PREC=fp16x2 BIGD_CHAINS=32768 python profiles/run_bigd_once.py > $OUT/r2_bigd_plain.txt 2>&1 || exit 1
BIGD_NITER=1 BIGD_PREC=fp16x2 BIGD_CHAINS=32768 $NCU --set full --import-source on -k regex:bigd_gemm_step -s 20 -c 1 -f -o $OUT/r2_bigd_full python profiles/run_bigd_once.py > $OUT/r2_ncu_bigd.log 2>&1
python profiles/run_nuts_once.py 18944 2 > $OUT/r2_nuts_plain.txt 2>&1 || exit 1
$NCU --set full --import-source on -k regex:hmc_nuts_tc_kernel -c 1 -f -o $OUT/r2_nuts_full python profiles/run_nuts_once.py 18944 2 > $OUT/r2_ncu_nuts.log 2>&1
ls -la $OUT/*.ncu-rep
