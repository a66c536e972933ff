#!/usr/bin/env bash
# tools/sass_census.sh -- per-kernel SASS opcode census of libpaged_attn.so (run anywhere: cuobjdump needs no GPU).
# Counts, per kernel, the mnemonics that prove the Blackwell paths (B200_PROFILING.md): UBLKCP (cp.async.bulk),
# UTMALDG / UTMASTG (TMA tensor loads / stores), UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st),
# UTCBAR (tcgen05.commit), SYNCS (mbarrier), ELECT, LDGSTS (cp.async), plus HMMA/IMMA (legacy mma.sync: expected 0),
# FFMA and MUFU.EX2 as the SIMT workload markers, registers and the total instruction count.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
LIB="${1:-$HERE/../llm.c-paged_b200/libpaged_attn.so}"
echo "# SASS census of $(basename "$LIB") ($(date -u +%Y-%m-%dT%H:%MZ)), cuobjdump $(cuobjdump --version | tail -1 | sed 's/.*release //')"
echo "# columns: instructions UBLKCP UTMALDG UTMASTG UTCHMMA LDTM STTM UTCBAR SYNCS ELECT LDGSTS HMMA FFMA MUFU.EX2 | kernel"
cuobjdump -sass "$LIB" | awk '
  function flush() {
    if (name != "") printf "%7d %6d %7d %7d %7d %5d %5d %6d %6d %5d %6d %5d %6d %6d | %s\n", n, c["UBLKCP"], c["UTMALDG"], c["UTMASTG"], c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTCBAR"], c["SYNCS"], c["ELECT"], c["LDGSTS"], c["HMMA"], c["FFMA"], c["EX2"], name
  }
  /Function :/ { flush(); name = $3; n = 0; delete c; next }
  /^ +\/\*[0-9a-f]+\*\// {
    n++
    if ($0 ~ /UBLKCP/) c["UBLKCP"]++
    if ($0 ~ /UTMALDG/) c["UTMALDG"]++
    if ($0 ~ /UTMASTG/) c["UTMASTG"]++
    if ($0 ~ /UTCHMMA|UTCQMMA|UTCOMMA/) c["UTCHMMA"]++
    if ($0 ~ /LDTM/) c["LDTM"]++
    if ($0 ~ /STTM/) c["STTM"]++
    if ($0 ~ /UTCBAR/) c["UTCBAR"]++
    if ($0 ~ /SYNCS/) c["SYNCS"]++
    if ($0 ~ /ELECT/) c["ELECT"]++
    if ($0 ~ /LDGSTS/) c["LDGSTS"]++
    if ($0 ~ /[ ;]HMMA|[ ;]IMMA/) c["HMMA"]++
    if ($0 ~ /FFMA/) c["FFMA"]++
    if ($0 ~ /MUFU\.EX2/) c["EX2"]++
  }
  END { flush() }' | while IFS='|' read -r counts name; do
    printf "%s| %s\n" "$counts" "$(echo "$name" | c++filt | cut -c1-110)"
  done | sort -t'|' -k2
