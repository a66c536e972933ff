#!/usr/bin/env bash
# oracle/build_dropin.sh -- TEST INFRASTRUCTURE ONLY.
# Proves the drop-in on the reference's OWN translation unit: builds the
# reference's paged_infer.c twice from where it lies under $REF (default /root/reference),
#   oracle/_ref/paged_infer_ref      as shipped: #include "block_manager.c", its own CPU add_to_cache /
#                                    attention_paged (paged_infer.c:10,16-18,163-240,505-573)
#   oracle/_ref/paged_infer_patched  INTEGRATION.md's patch applied ON THE PIPE INTO gcc and linked against
#                                    libpaged_attn.so: the include becomes "paged_attn.h", the three geometry
#                                    macros and the bodies of attention_paged / add_to_cache go away
# (no reference source is copied into the repo).  main (:953-1101), gpt2_forward with its call site
# (:696-716), the checkpoint / token / tokenizer readers and every other op compile UNCHANGED.
# One line is added to BOTH builds so the integer state can be compared: the reference's own
# print_state(block_manager, 0) just before main prints "Finished!".
# tests/test_gpu_dropin.py runs both binaries on the same synthetic L=1 checkpoint (written with
# pa_checkpoint_write; with L=1 the fork's `l < 1` layer loop, :659, is the whole model) and asserts identical
# generated token ids and block tables.
set -euo pipefail
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/.." && pwd)"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "build_dropin.sh: $REF absent (GPU box?) -- keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT"
FLAGS="-O3 -Ofast -Wno-unused-result -fopenmp -DOMP -w"     # the reference Makefile's flags (Makefile:2,33), gcc for clang

instrument() {   # the reference's own debug dump of prompt 0, once, at the end of main
  sed -e 's/^\( *\)printf("Finished!\\n");/\1print_state(block_manager, 0);\n\1printf("Finished!\\n");/'
}
drop_function() {   # $1 = name: removes `void name(...) { ... }` (definition starts at column 0, ends at the first `}` at column 0)
  awk -v pat="^void $1\\\\(" '
    skipping { if ($0 ~ /^}/) skipping = 0; next }
    $0 ~ pat { skipping = 1; next }
    { print }'
}
patch_tu() {
  sed -e 's|^#include "block_manager.c"|#include "paged_attn.h"|' \
      -e '/^#define MAX_PROMPTS /d' -e '/^#define MAX_BLOCKS /d' -e '/^#define BLOCK_SIZE /d' |
    drop_function attention_paged | drop_function add_to_cache
}

ref_bin="$OUT/paged_infer_ref"
pat_bin="$OUT/paged_infer_patched"
if [ ! -f "$ref_bin" ] || [ "$HERE/build_dropin.sh" -nt "$ref_bin" ]; then
  ( cd "$REF" && instrument < paged_infer.c | gcc $FLAGS -x c - -I"$REF" -lm -o "$ref_bin" )
  echo "built $ref_bin"
fi
lib="$ROOT/llm.c-paged_b200/libpaged_attn.so"
if [ ! -f "$pat_bin" ] || [ "$HERE/build_dropin.sh" -nt "$pat_bin" ] || [ "$ROOT/include/paged_attn.h" -nt "$pat_bin" ] || [ "$lib" -nt "$pat_bin" ]; then
  [ -f "$lib" ] || { echo "build_dropin.sh: build libpaged_attn.so first" >&2; exit 1; }
  instrument < "$REF/paged_infer.c" | patch_tu |
    gcc $FLAGS -x c - -I"$ROOT/include" -L"$ROOT/llm.c-paged_b200" -lpaged_attn \
        -Wl,-rpath,'$ORIGIN/../../llm.c-paged_b200' -lm -o "$pat_bin"
  echo "built $pat_bin"
fi
# ---- the reference's own unit test of the block manager (block_manager_test.c), as shipped and against the library:
# the include becomes "paged_attn.h" (with PA_COMPAT_MACROS for its loops over BLOCK_SIZE) and the final free(manager)
# becomes destroy_block_manager(manager) (the manager owns a device pool); run with PA_COMPAT_HOST_PAGES=1 because the
# test writes and reads pages through KVBlock.keys / .values from the host.
bt_ref="$OUT/block_manager_test_ref"
bt_pat="$OUT/block_manager_test_patched"
if [ ! -f "$bt_ref" ] || [ "$HERE/build_dropin.sh" -nt "$bt_ref" ]; then
  ( cd "$REF" && gcc -O2 -w -x c block_manager_test.c -I"$REF" -o "$bt_ref" )
  echo "built $bt_ref"
fi
if [ ! -f "$bt_pat" ] || [ "$HERE/build_dropin.sh" -nt "$bt_pat" ] || [ "$ROOT/include/paged_attn.h" -nt "$bt_pat" ] || [ "$lib" -nt "$bt_pat" ]; then
  sed -e 's|^#include "block_manager.c"|#define PA_COMPAT_MACROS\n#include <stdlib.h>\n#include "paged_attn.h"|' \
      -e 's|^\( *\)free(manager);|\1destroy_block_manager(manager);|' "$REF/block_manager_test.c" |
    gcc -O2 -w -x c - -I"$ROOT/include" -L"$ROOT/llm.c-paged_b200" -lpaged_attn \
        -Wl,-rpath,'$ORIGIN/../../llm.c-paged_b200' -o "$bt_pat"
  echo "built $bt_pat"
fi

# what the patch removed / kept, for the record (stderr): the patched unit must not define the two functions
n=$(instrument < "$REF/paged_infer.c" | patch_tu | grep -c '^void attention_paged(\|^void add_to_cache(' || true)
[ "$n" = "0" ] || { echo "build_dropin.sh: the patch left $n definitions behind" >&2; exit 1; }
