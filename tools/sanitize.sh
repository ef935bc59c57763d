#!/bin/bash
# compute-sanitizer over the C++ drop-in test programs (reference practice: /root/reference/ci/test_cub.sh:57-76,
# cub/test/run_test.cmake:85-143): memcheck, racecheck, synccheck, initcheck.  Sizes are capped (B200RS_TEST_MAX_N) because
# racecheck is ~100x slower than native.  usage (GPU box): bash tools/sanitize.sh <out-dir>
OUT=${1:-gpurun_out/sanitizer}
mkdir -p $OUT
for tool in memcheck racecheck synccheck initcheck; do
  for prog in test_cub_shim test_thrust_shim; do
    B200RS_TEST_MAX_N=100000 timeout 1200 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 30 \
      tools/bin/$prog > $OUT/${tool}_${prog}.log 2>&1
    echo "$tool $prog exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|all checks passed' $OUT/${tool}_${prog}.log | tr '\n' ' ')"
  done
done
