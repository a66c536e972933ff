#!/usr/bin/env bash
# tools/profile_r02.sh -- the round-2 ncu evidence (run under gpurun on ONE B200): each command first plain (must exit 0),
# then the launch list / the --set full capture of the dominant kernel.  Outputs under gpurun_out/.
set -uo pipefail
OUT=gpurun_out
COMMON="--configs none --no-cpu-baseline --no-clock-hold --no-verify --no-e2e --no-prefill"
CMD2="python bench.py --steps 3 --warmup 3 $COMMON"
CMD4="python bench.py --workload cfg4 --steps 1 --warmup 3 $COMMON --no-model"
CMD5="python bench.py --workload cfg5 --steps 1 --warmup 3 $COMMON --no-model --no-tc-prefill"
$CMD2 > $OUT/r02_plain_cfg2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r02_launches_cfg2.csv $CMD2 > $OUT/r02_ncu_cfg2.log 2>&1
echo "launch list cfg2: rc=$?"
$CMD4 > $OUT/r02_plain_cfg4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pa_decode_stream -s 60 -c 1 -f -o $OUT/r02_decode_cfg4 $CMD4 > $OUT/r02_ncu_cfg4.log 2>&1
echo "full cfg4: rc=$?"
$CMD5 > $OUT/r02_plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pa_decode_stream -s 2 -c 1 -f -o $OUT/r02_decode_cfg5 $CMD5 > $OUT/r02_ncu_cfg5.log 2>&1
echo "full cfg5: rc=$?"
ls -la $OUT/*.ncu-rep 2>/dev/null
