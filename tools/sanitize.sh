#!/usr/bin/env bash
# tools/sanitize.sh -- compute-sanitizer over the GPU suite's decode (stream kernel: mbarrier ring, self-resetting
# arrival counters, cross-CTA merge), split-K projection (L2 workspace, spin-waits) and persistent step kernel
# (grid barriers) cases.  Run on a GPU box; summaries land in $OUT (default gpurun_out/sanitizer).
#   memcheck  : out-of-bounds / misaligned global, shared and local accesses, leaks of device memory
#   racecheck : shared-memory hazards between the warps of a CTA
#   synccheck : invalid use of bar.sync / mbarrier / __syncwarp
set -uo pipefail
OUT="${OUT:-gpurun_out/sanitizer}"
mkdir -p "$OUT"
SEL_DECODE='test_decode_matches_oracle and (ragged-bs16-shuffled or xl-25heads or hs128-bs32) or test_decode_append_fused and 12-64-16-0 or test_decode_tile_shapes_and_splits and 3-3-7'
SEL_GEMM='test_qkv_append_split_k or test_matmul_bias_auto_split_k or test_split_k_with_two_streams_live'
SEL_MEGA='test_model_persistent_step_kernel and (4-128-16-8-70 or 2-64-4-3-1) or test_model_decode_steps_match_oracle and auto and 3-2-64-131-16-0'
run() {   # $1 tool, $2 name, $3 file, $4 -k expression
  local log="$OUT/$1_$2.log"
  timeout "${TMO:-900}" compute-sanitizer --tool "$1" --error-exitcode 97 --print-limit 20 ${EXTRA:-} \
      python -m pytest "$3" -m gpu -x -q -k "$4" > "$log" 2>&1
  local rc=$?
  {
    echo "== $1 / $2: exit $rc"
    grep -E "passed|failed|deselected" "$log" | tail -1
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY" "$log" | tail -3
  } | tee -a "$OUT/summary.txt"
}
: > "$OUT/summary.txt"
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  run "$tool" decode tests/test_gpu_parity.py "$SEL_DECODE"
  run "$tool" gemm   tests/test_gpu_qkv.py    "$SEL_GEMM"
  run "$tool" mega   tests/test_gpu_model.py  "$SEL_MEGA"
done
cat "$OUT/summary.txt"
