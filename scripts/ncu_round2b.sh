#!/bin/bash
# ncu --set full of a LARGE CONV-variant GEMM launch (combined encoder) and of the tcgen05 attention kernel, VidOR-shaped batch.
set -u
OUT=gpurun_out
CMD="python bench.py --workload vidor --videos 12 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --vidor-videos 0"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)3, \(int\)128, \(bool\)0, \(int\)1, \(bool\)1' --launch-skip 9 -c 2 -f -o $OUT/r2_prof_gemm_conv_big $CMD > $OUT/r2_prof_gemm_conv_big.log 2>&1
ncu -i $OUT/r2_prof_gemm_conv_big.ncu-rep --page details > $OUT/r2_ncu_full_gemm_conv_big.txt 2>/dev/null
ncu -i $OUT/r2_prof_gemm_conv_big.ncu-rep --page source --csv > $OUT/r2_src_gemm_conv_big.csv 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:mha16_tc_kernel' --launch-skip 1 -c 1 -f -o $OUT/r2_prof_mha16_tc $CMD > $OUT/r2_prof_mha16_tc.log 2>&1
ncu -i $OUT/r2_prof_mha16_tc.ncu-rep --page details > $OUT/r2_ncu_full_mha16_tc.txt 2>/dev/null
ncu -i $OUT/r2_prof_mha16_tc.ncu-rep --page source --csv > $OUT/r2_src_mha16_tc.csv 2>/dev/null
ls -la $OUT/*.ncu-rep | tail -3
