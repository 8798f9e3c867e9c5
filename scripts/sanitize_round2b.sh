#!/bin/bash
# compute-sanitizer over the kernels added in the second session of round 2: mode-5 GEMM family (in-place fp16 split behind named barriers),
# mha64_tc_kernel, role_attention_hid_kernel.  Output: gpurun_out/r3_sanitizer.txt
OUT=gpurun_out/r3_sanitizer.txt
T1='tests/test_gpu_gemm.py -k "test_gemm_modes and 5- or test_gemm_epilogue_options and 5 or test_cta_pair_multicast_equals_single_cta and 5 or test_fused_dwconv_gemm_equals_dwconv_then_gemm and 5"'
T2='tests/test_gpu_attention.py -k tc64'
T3='tests/test_gpu_bigc.py -k "role_fold and (mid128 or vidvrd)"'
{
echo "compute-sanitizer over the second-session kernels (B200): mode-5 GEMM tests, vsg_mha_tc64 tests, role-fold tests"
for tool in memcheck synccheck; do
  echo "== $tool =="
  for T in "$T1" "$T2" "$T3"; do
    eval timeout 900 compute-sanitizer --tool $tool python -m pytest $T -x -q -m gpu 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|error" | head -8
  done
done
echo "== racecheck =="
for T in "$T1" "$T2" "$T3"; do
  eval timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest $T -x -q -m gpu > /tmp/race.log 2>&1
  grep -E "passed|failed|RACECHECK SUMMARY" /tmp/race.log | head -4
  grep -E "hazard detected" /tmp/race.log | sed 's/at __shared__.*//' | sort | uniq -c | sort -rn | head -6
  grep -E "(Read|Write) Thread" /tmp/race.log | sed -E 's/Thread \([0-9,]+\)/Thread/; s/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -8
done
} > $OUT 2>&1
cat $OUT
