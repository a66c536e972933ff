#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own paged path (block_manager.c + paged_infer.c) from the
# sources where they lie under $REF (default /root/reference) into shared objects
# under oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).
# No reference source is copied into the repo: the patched text only ever exists on
# the pipe into gcc.  The ONLY edit is the geometry macros, which are unconditional
# #defines (block_manager.c:4-6, paged_infer.c:16-18) and so cannot be set with -D.
#
# Each geometry is built twice:
#   fast   = the reference Makefile's flags (-O3 -Ofast -Wno-unused-result -fopenmp -DOMP,
#            Makefile:2,33) with gcc because clang (Makefile:1) is absent
#   strict = -O2 -fno-fast-math -ffp-contract=off  (IEEE evaluation order as written)
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_ref.sh: $REF absent (GPU box?) -- keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT"

# block_size:max_blocks:max_prompts
VARIANTS="${VARIANTS:-32:100:100 16:100:100 16:4352:256 16:12:8 8:64:16 2:64:8 4:24:6}"
FAST="-O3 -Ofast -Wno-unused-result -fopenmp -DOMP"
STRICT="-O2 -fno-fast-math -ffp-contract=off -Wno-unused-result -fopenmp -DOMP"
COMMON="-x c - -shared -fPIC -fvisibility=hidden -Wl,-Bsymbolic -w -lm"

patch_macros() { # $1 bs $2 mb $3 mp
  sed -E -e "s/^#define MAX_PROMPTS .*/#define MAX_PROMPTS $3/" \
         -e "s/^#define MAX_BLOCKS .*/#define MAX_BLOCKS $2/" \
         -e "s/^#define BLOCK_SIZE .*/#define BLOCK_SIZE $1/"
}
gen() { # emits the translation unit on stdout
  patch_macros "$1" "$2" "$3" < "$REF/block_manager.c"
  echo    # block_manager.c has no trailing newline
  echo '#define main ref_paged_infer_main'
  patch_macros "$1" "$2" "$3" < "$REF/paged_infer.c" | sed -e '/#include "block_manager.c"/d'
  echo
  echo '#undef main'
  cat "$HERE/ref_wrap.c"
}
for v in $VARIANTS; do
  IFS=: read -r bs mb mp <<< "$v"
  for flavor in fast strict; do
    if [ "$flavor" = fast ]; then fl="$FAST"; else fl="$STRICT"; fi
    so="$OUT/libref_bs${bs}_mb${mb}_mp${mp}_${flavor}.so"
    if [ ! -f "$so" ] || [ "$HERE/ref_wrap.c" -nt "$so" ] || [ "$HERE/build_ref.sh" -nt "$so" ]; then
      gen "$bs" "$mb" "$mp" | gcc $fl $COMMON -o "$so"
      echo "built $so"
    fi
  done
done
# contiguous attention_forward from the trainer (differential oracle)
for flavor in fast strict; do
  if [ "$flavor" = fast ]; then fl="$FAST"; else fl="$STRICT"; fi
  so="$OUT/libref_train_${flavor}.so"
  if [ ! -f "$so" ] || [ "$HERE/ref_train_wrap.c" -nt "$so" ]; then
    { cat "$REF/train_gpt2.c"; echo; cat "$HERE/ref_train_wrap.c"; } | gcc -DTESTING $fl $COMMON -o "$so"
    echo "built $so"
  fi
done
