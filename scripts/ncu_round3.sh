#!/bin/bash
# Round-2 (second session) evidence: launch lists of one bench step after the fused decoder attention (VidVRD default mode, fp16x3 mode,
# VidOR incl. grounding) and ncu --set full captures of the conv-tap GEMM in the fp16x3 mode and of the fused attention kernel.
set -u
OUT=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --vidor-videos 0 --no-graph --no-pipeline --modes="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r3_launches_vidvrd200.csv $B > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/r3_launches_vidvrd200_fp16x3.csv $B --precision fp16x3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/r3_launches_vidor50.csv $B --workload vidor --videos 50 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:gemm_tc_kernel<\(int\)5, \(int\)256, \(bool\)0, \(int\)3' --launch-skip 3 -c 2 -f -o $OUT/r3_prof_gemm_fp16x3 $B --precision fp16x3 > /dev/null 2>&1
ncu -i $OUT/r3_prof_gemm_fp16x3.ncu-rep --page details > $OUT/r3_ncu_full_gemm_fp16x3.txt 2>/dev/null
ls -la $OUT/r3_launches_*.csv $OUT/r3_ncu_full_gemm_*.txt
