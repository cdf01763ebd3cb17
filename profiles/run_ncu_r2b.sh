#!/bin/bash
# Round-2 (second half) ncu evidence: the all-lags FFT variogram kernel.  Plain run first, numbers under ncu are not bench values.
set -u
OUT=gpurun_out
NCU="ncu --clock-control none"
python profiles/run_diag_once.py 65536 800 > $OUT/r2b_diag_plain.txt 2>&1 || exit 1
HMC_B200_DIAG_FFT_FREE=1 python profiles/run_diag_once.py 65536 800 > $OUT/r2b_diag_plain_free.txt 2>&1 || exit 1
$NCU --set full --import-source on -k regex:"diag_fft_power" -s 1 -c 1 -f -o $OUT/r2b_fft_full python profiles/run_diag_once.py 65536 800 > $OUT/r2b_ncu_fft.log 2>&1
ls -la $OUT/*.ncu-rep
