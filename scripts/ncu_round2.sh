#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 item 4 asks for (run on the GPU box: gpurun -- bash scripts/ncu_round2.sh).
# A small VidOR-shaped batch (12 videos incl. grounding + eval) provides every kernel in bench context.
set -u
OUT=gpurun_out
CMD="python bench.py --workload vidor --videos 12 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --vidor-videos 0"
cap() {  # name, regex, count
  timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:$2" -c "$3" -f -o $OUT/r2_prof_$1 $CMD > $OUT/r2_prof_$1.log 2>&1
  ncu -i $OUT/r2_prof_$1.ncu-rep --page details > $OUT/r2_ncu_full_$1.txt 2>/dev/null
  ls -la $OUT/r2_prof_$1.ncu-rep
}
cap mha16 'mha_kernel<\(int\)16>' 3
cap relmatch 'rel_ov_kernel|greedy_match_kernel|rank_kernel' 6
cap gemm_conv 'gemm_tc_kernel<\(int\)3, \(int\)128, \(bool\)0, \(int\)1, \(bool\)1' 4
cap cq 'cq_attention_kernel|grounding_post_kernel' 2
