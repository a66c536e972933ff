#!/usr/bin/env bash
# ncu --set full of the 3xTF32 prefill kernel at the two shapes quoted in profiles/r02_prefill.md (one B200)
set -uo pipefail
OUT=gpurun_out
A="python tools/prefill_bench.py --shape 124m --B 16 --T 2048 --path 4 --iters 2 --warmup 1"
B="python tools/prefill_bench.py --shape long --B 2 --T 2048 --before 30000 --path 4 --iters 2 --warmup 1"
$A > $OUT/r02_plain_tc3_124m.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pa_prefill_tc3 -s 1 -c 1 -f -o $OUT/r02_prefill_tc3_124m $A > $OUT/r02_ncu_tc3_124m.log 2>&1
echo "tc3 124m: rc=$?"
$B > $OUT/r02_plain_tc3_long.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pa_prefill_tc3 -s 1 -c 1 -f -o $OUT/r02_prefill_tc3_long $B > $OUT/r02_ncu_tc3_long.log 2>&1
echo "tc3 long: rc=$?"
