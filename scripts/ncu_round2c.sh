#!/bin/bash
# Round-2 evidence: launch lists of one bench step (VidVRD default mode, VidVRD bf16 mode, VidOR incl. grounding) and ncu --set full
# captures of the two largest GEMM launches in both modes (conv taps: M = 481k, N = 1536, K = 1024).
set -u
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --vidor-videos 0 --no-graph --no-pipeline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r2_launches_vidvrd200.csv $B > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r2_launches_vidvrd200_bf16.csv $B --precision bf16 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r2_launches_vidor50.csv $B --workload vidor --videos 50 > /dev/null 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:gemm_tc_kernel' -c 400 --csv --log-file $OUT/r2_gemm_traffic.csv $B > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)3, \(int\)256, \(bool\)0, \(int\)3' --launch-skip 3 -c 2 -f -o $OUT/r2_prof_gemm_default $B > /dev/null 2>&1
ncu -i $OUT/r2_prof_gemm_default.ncu-rep --page details > $OUT/r2_ncu_full_gemm_default.txt 2>/dev/null
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)1, \(int\)256, \(bool\)0, \(int\)3' --launch-skip 2 -c 3 -f -o $OUT/r2_prof_gemm_bf16 $B --precision bf16 > /dev/null 2>&1
ncu -i $OUT/r2_prof_gemm_bf16.ncu-rep --page details > $OUT/r2_ncu_full_gemm_bf16.txt 2>/dev/null
ls -la $OUT/r2_launches_*.csv $OUT/r2_gemm_traffic.csv $OUT/r2_ncu_full_gemm_*.txt
