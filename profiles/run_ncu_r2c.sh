#!/bin/bash
# Round-2 final ncu evidence (one GPU): launch list of the bench's legs and a full capture of the headline kernel.
# The program is first run WITHOUT ncu (exit 0 required); numbers printed under ncu are never bench values.
set -u
OUT=gpurun_out
NCU="ncu --clock-control none"
python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > $OUT/r2c_ncu_plain_bench.json 2> $OUT/r2c_ncu_plain_bench.log || exit 1
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/r2c_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-secondary > $OUT/r2c_ncu_list.log 2>&1
$NCU --set full --import-source on -k regex:hmc_random_tc_kernel -s 1 -c 1 -f -o $OUT/r2c_tc_full python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > $OUT/r2c_ncu_tc.log 2>&1
ls -la $OUT/*.ncu-rep
