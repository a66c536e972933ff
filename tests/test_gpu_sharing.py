"""-m gpu: the device side of the allocator extensions.  A forked sequence and a
sequence built on a cached prefix must attend exactly as an independently filled sequence with the
same tokens would (oracle: one private copy of everything)."""
import numpy as np
import pytest

import oracle_api as oa
from gpu_common import Scenario, assert_close, pa

pytestmark = pytest.mark.gpu


def _mirror_seq(sc, s):
    """(Re)build oracle sequence s from what the device pool holds for engine sequence s."""
    eng, orc, bs = sc.eng, sc.orc, sc.bs
    orc.free_blocks_for_prompt(s)
    n = eng.seq_len(s)
    tbl = eng.table(s)
    slots = np.array([tbl[p // bs] * bs + p % bs for p in range(n)], dtype=np.int32)
    k, v = eng.read_pool_rows(sc.layer, slots)
    for t0 in range(0, n, bs):
        idx = orc.request_block(s)
        ok, ov = orc.page_arrays(idx)
        m = min(bs, n - t0)
        ok[:m], ov[:m] = k[t0:t0 + m], v[t0:t0 + m]
        orc.set_filled(idx, m)


@pytest.mark.parametrize("ctx0", [37, 32, 5])
def test_forked_sequences_decode_like_private_copies(ctx0):
    NH, hs, bs = 2, 64, 16
    Cc = NH * hs
    sc = Scenario(NH, hs, bs, [ctx0, 0, 20], seed=401, extra_blocks=16)
    try:
        eng, orc = sc.eng, sc.orc
        assert eng.seq_fork(0, 1) == 0, pa.last_error()
        assert eng.seq_len(1) == ctx0
        shared = ctx0 // bs
        assert list(eng.table(1))[:shared] == list(eng.table(0))[:shared]
        if ctx0 % bs:
            assert eng.table(1)[-1] != eng.table(0)[-1]
        _mirror_seq(sc, 1)
        for step in range(3):                                   # the copies diverge: different new tokens each step
            qkv = oa.normal((3, 3 * Cc), seed=410 + step)
            assert eng.step_begin([0, 1, 2], [1, 1, 1]) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            d, o = pa.DevBuf.from_numpy(qkv), pa.DevBuf(3 * Cc * 4)
            pa.check(eng.decode_append(0, d.ptr, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc, o.ptr, Cc), "decode_append")
            eng.sync()
            for s in range(3):
                orc.add_to_cache(qkv[s][None, None, :], 1, 1, 1, prompt=s)
            want = orc.decode_batch([0, 1, 2], NH, qkv[:, :Cc])
            assert_close(o.download((3, Cc)), want, f"fork, step {step}")
        # shared pages were never written: sequence 0 still reads its original prefix
        assert eng.lib.pa_page_refcount(eng.h, int(eng.table(0)[0])) == (2 if shared else 1)
    finally:
        sc.close()


def test_prefix_cache_hit_gives_the_same_attention_as_a_full_prefill():
    NH, hs, bs = 2, 64, 16
    Cc = NH * hs
    n0 = 50
    sc = Scenario(NH, hs, bs, [n0, 0], seed=421, extra_blocks=16, max_batch_tokens=64)
    try:
        eng, orc = sc.eng, sc.orc
        toks0 = np.arange(1000, 1000 + n0, dtype=np.int32)
        assert eng.prefix_insert(0, toks0) == 3                 # 48 tokens = 3 full pages registered
        # a second prompt shares the first 40 tokens: 2 pages (32 tokens) come from the cache
        toks1 = np.concatenate([toks0[:40], np.arange(7000, 7015, dtype=np.int32)])
        hit = eng.prefix_match(1, toks1)
        assert hit == 32 and list(eng.table(1)) == list(eng.table(0))[:2]
        rest = len(toks1) - hit
        qkv = oa.normal((rest, 3 * Cc), seed=422)
        # what a model would produce for tokens 32..39 is the same K/V as sequence 0 holds there; the
        # synthetic test simply reuses sequence 0's rows for those positions
        tbl0 = eng.table(0)
        slots = np.array([tbl0[p // bs] * bs + p % bs for p in range(32, 40)], dtype=np.int32)
        k0, v0 = eng.read_pool_rows(0, slots)
        qkv[:8, Cc:2 * Cc], qkv[:8, 2 * Cc:] = k0, v0
        assert eng.step_begin([1], [rest]) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        d, o = pa.DevBuf.from_numpy(qkv), pa.DevBuf(rest * Cc * 4)
        pa.check(eng.append(0, d.ptr + Cc * 4, d.ptr + 2 * Cc * 4, 3 * Cc), "append")
        pa.check(eng.prefill(0, d.ptr, 3 * Cc, o.ptr, Cc), "prefill")
        eng.sync()
        got = o.download((rest, Cc))
        # oracle: sequence 1 with a PRIVATE copy of all 55 tokens
        _mirror_seq(sc, 1)
        want = orc.attend_rows([1], [0], [hit + 1], [rest], NH, qkv[:, :Cc])
        assert_close(got, want, "prefill on top of a cached prefix")
        # first 40 positions of both sequences hold identical K/V
        s1 = np.array([eng.table(1)[p // bs] * bs + p % bs for p in range(40)], dtype=np.int32)
        s0 = np.array([tbl0[p // bs] * bs + p % bs for p in range(40)], dtype=np.int32)
        ka, _ = eng.read_pool_rows(0, s1)
        kb, _ = eng.read_pool_rows(0, s0)
        assert np.array_equal(ka, kb)
    finally:
        sc.close()


def test_swapped_out_sequence_attends_identically_after_swap_in():
    """Eviction with swapping on: the victim's K/V (both layers) go to host memory and come back into
    different pages; decode over it equals the never-evicted oracle."""
    NH, hs, bs, L = 2, 64, 16, 2
    Cc = NH * hs
    ctx = [40, 33, 48]
    eng = pa.PagedAttn(bs, 9, 3, NH, hs, n_layers=L, device=0, max_batch_tokens=64)      # 9 pages: exactly full
    orcs = [oa.OrcManager(Cc, bs, 64, 3) for _ in range(L)]
    lib = eng.lib
    try:
        assert lib.pa_set_evict_swap(eng.h, 1) == 0
        for s, n in enumerate(ctx):
            assert eng.step_begin([s], [n]) == 0, pa.last_error()
            slots = eng.slot_mapping().copy()
            for l in range(L):
                kv = oa.normal((n, 2, Cc), seed=900 + 10 * s + l)
                eng.write_pool_rows(l, slots, kv[:, 0], kv[:, 1])
                for t in range(n):
                    row = np.concatenate([np.zeros(Cc, np.float32), kv[t, 0], kv[t, 1]])
                    orcs[l].add_to_cache(row[None, None, :], 1, 1, 1, prompt=s)
        # sequence 1 grows into a new page: the pool is full, the LRU prompt (0) is swapped out
        assert eng.step_begin([1], [16]) == 0, pa.last_error()
        assert eng.seq_len(0) == 0 and lib.pa_seq_swapped_tokens(eng.h, 0) == 40
        old_table = None
        # decode over sequence 0 again: it is swapped back in (into other pages); the next victim of the
        # reference's page-LRU is swapped out in turn
        q = oa.normal((1, Cc), seed=950)
        for l in range(L):
            assert eng.step_begin_readonly([0]) == 0, pa.last_error()
            if old_table is None:
                old_table = list(eng.table(0))
                assert eng.seq_len(0) == 40 and lib.pa_seq_swapped_tokens(eng.h, 0) == 0
                assert lib.pa_seq_swapped_tokens(eng.h, 1) + lib.pa_seq_swapped_tokens(eng.h, 2) > 0
            pa.check(eng.upload(), "upload")
            dq, do = pa.DevBuf.from_numpy(q), pa.DevBuf(Cc * 4)
            pa.check(eng.decode(l, dq.ptr, Cc, do.ptr, Cc), "decode")
            eng.sync()
            want = orcs[l].decode_batch([0], NH, q)
            assert_close(do.download((1, Cc)), want, f"layer {l} after swap-in")
    finally:
        eng.close()
        for o in orcs:
            o.close()
