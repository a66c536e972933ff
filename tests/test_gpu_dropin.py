"""The drop-in proven on the reference's OWN translation unit.

oracle/build_dropin.sh compiles /root/reference/paged_infer.c twice: as shipped (its CPU add_to_cache /
attention_paged, `#include "block_manager.c"`), and with INTEGRATION.md's patch applied on the pipe into gcc
and linked against libpaged_attn.so.  Both run the reference's `main` (paged_infer.c:953-1101) unchanged on
the same synthetic files in the reference's formats: an L=1 checkpoint (with one layer the fork's `l < 1`
loop, :659, is the whole model), an int32 token stream and a tokenizer whose pieces spell the token id.
Generated token ids and the final block-manager state (print_state) must be identical."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import __graft_entry__ as ge

REF_BIN = os.path.join(ge.ROOT, "oracle", "_ref", "paged_infer_ref")
PATCHED_BIN = os.path.join(ge.ROOT, "oracle", "_ref", "paged_infer_patched")


def write_inputs(d, seed=2024):
    pa = ge.load_binding()
    lib = pa.load()
    maxT, V, L, NH, Cc = 64, 50257, 1, 2, 128        # V as GPT-2's: main fills the tail of gen_tokens with GPT2_EOT = 50256
    cfg = pa.PaModelConfig(maxT, V, L, NH, Cc)
    n = lib.pa_model_param_count(C.byref(cfg))
    rng = np.random.default_rng(seed)
    params = (rng.standard_normal(n) * 0.08).astype(np.float32)
    pa.check(lib.pa_checkpoint_write(os.path.join(d, "gpt2_124M.bin").encode(), C.byref(cfg), params.ctypes.data), "checkpoint")
    os.makedirs(os.path.join(d, "data"), exist_ok=True)
    ids = rng.integers(0, V - 1, size=4096).astype(np.int32)
    pa.check(lib.pa_tokens_write(os.path.join(d, "data", "tiny_shakespeare_val.bin").encode(), pa.iptr(ids), len(ids)), "tokens")
    pieces = [f"<{i}>".encode() for i in range(V)]
    arr = (C.c_char_p * V)(*pieces)
    lens = (C.c_ubyte * V)(*[len(p) for p in pieces])
    pa.check(lib.pa_tokenizer_write(os.path.join(d, "gpt2_tokenizer.bin").encode(), arr, lens, V), "tokenizer")
    return ids


def run_main(exe, cwd, env=None):
    r = subprocess.run([exe], cwd=cwd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, f"{exe}: rc {r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}"
    gen = r.stdout.split("Start sliding the window", 1)[1].split("\n---\n", 1)[0]
    tokens = [int(t) for t in re.findall(r"<(\d+)>", gen)]
    # print_state is also called once before the run ("State before running"): the LAST dump is the final state
    final = re.findall(r"Block manager llru \d+\nPrompt 0 block count: \d+\n(?:Block \d+: filled \d+, llru \d+\n)*", r.stdout)
    return tokens, (final[-1] if final else ""), r.stdout


def test_reference_main_runs_unmodified_on_synthetic_files(tmp_path):
    """CPU: the as-shipped build of the reference's main reads the files pa_formats wrote and generates 18 tokens
    (the sliding window from t = 32 to 49: one first pass of 32 rows, 17 single-row appends) with the cache in
    pages 0 (32 rows) and 1 (17 rows)."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/paged_infer_ref not built (no /root/reference at build time)")
    write_inputs(str(tmp_path))
    tokens, final, out = run_main(REF_BIN, str(tmp_path))
    assert len(tokens) == 18, out[-2000:]
    assert "Prompt 0 block count: 2" in final and "Block 0: filled 32" in final and "Block 1: filled 17" in final, final


@pytest.mark.gpu
def test_patched_reference_main_generates_the_same_tokens_and_block_tables(tmp_path):
    if not (os.path.exists(REF_BIN) and os.path.exists(PATCHED_BIN)):
        pytest.skip("oracle/_ref/paged_infer_{ref,patched} not built (no /root/reference at build time)")
    write_inputs(str(tmp_path))
    want_tokens, want_state, _ = run_main(REF_BIN, str(tmp_path))
    got_tokens, got_state, out = run_main(PATCHED_BIN, str(tmp_path))
    assert len(want_tokens) == 18
    assert got_tokens == want_tokens, out[-2000:]
    assert got_state == want_state and "Block 1: filled 17" in got_state, (got_state, want_state)
    # another geometry through the environment: block 16 -> pages 0,1 full + page 2..3 (the reference cannot: its first
    # append of T = 32 rows must fit ONE page, paged_infer.c:542-545 -- the library crosses page boundaries)
    env = dict(os.environ, PA_BLOCK_SIZE="16")
    got16, state16, _ = run_main(PATCHED_BIN, str(tmp_path), env=env)
    assert got16 == want_tokens
    assert "Prompt 0 block count: 4" in state16 and "Block 3: filled 1" in state16, state16


BT_REF = os.path.join(ge.ROOT, "oracle", "_ref", "block_manager_test_ref")
BT_PATCHED = os.path.join(ge.ROOT, "oracle", "_ref", "block_manager_test_patched")


def test_reference_block_manager_test_as_shipped():
    if not os.path.exists(BT_REF):
        pytest.skip("oracle/_ref/block_manager_test_ref not built (no /root/reference at build time)")
    r = subprocess.run([BT_REF], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "All tests passed!" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_block_manager_test_against_the_library():
    """The reference's own block_manager_test.c (request_block, host writes and read-backs through KVBlock.keys /
    .values, `filled`, free_blocks_for_prompt) compiled against paged_attn.h and linked with libpaged_attn.so: the
    pool of a create_block_manager() manager lives in managed memory under PA_COMPAT_HOST_PAGES=1, so the reference's
    host-side page accesses work on the very pointers the kernels use."""
    if not os.path.exists(BT_PATCHED):
        pytest.skip("oracle/_ref/block_manager_test_patched not built (no /root/reference at build time)")
    env = dict(os.environ, PA_COMPAT_HOST_PAGES="1")
    r = subprocess.run([BT_PATCHED], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and "All tests passed!" in r.stdout, r.stdout + r.stderr
