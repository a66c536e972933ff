"""allocator extensions (sequence fork with shared pages, prefix cache by hashing).
Integer logic on host-only handles (CPU); the device side is covered by tests/test_gpu_sharing.py.
With the extensions unused the block manager must still reproduce the reference trace: that is
what tests/test_block_manager.py and tests/test_golden.py keep checking."""
import numpy as np
import pytest

import __graft_entry__ as ge

pa = ge.load_binding()


def make(bs=4, blocks=12, seqs=6):
    return pa.PagedAttn(bs, blocks, seqs, 2, 8, device=pa.PA_HOST_ONLY, max_batch_tokens=64)


def refs(eng, n):
    return [eng.lib.pa_page_refcount(eng.h, i) for i in range(n)]


def test_fork_shares_full_pages_and_copies_the_partial_one():
    eng = make()
    try:
        assert eng.step_begin([0], [10]) == 0               # pages 0,1 full, page 2 holds 2 rows
        assert list(eng.table(0)) == [0, 1, 2]
        assert eng.seq_fork(0, 1) == 0
        assert list(eng.table(1)) == [0, 1, 3]               # full pages shared, the partial one copied to a new page
        assert refs(eng, 4) == [2, 2, 1, 1]
        assert eng.seq_len(1) == 10
        # both append independently: only private pages change
        assert eng.step_begin([0, 1], [3, 1]) == 0
        assert list(eng.table(0)) == [0, 1, 2, 4] and list(eng.table(1)) == [0, 1, 3]
        assert eng.seq_len(0) == 13 and eng.seq_len(1) == 11
        # freeing one holder keeps the shared pages alive, with a valid owner for the LRU
        assert eng.seq_free(0) == 0
        assert refs(eng, 5) == [1, 1, 0, 1, 0]
        assert eng.mgr.contents.blocks[0].prompt_id == 1 and eng.mgr.contents.blocks[2].prompt_id == -1
        assert eng.seq_free(1) == 0
        assert refs(eng, 5) == [0, 0, 0, 0, 0]
        # fork at a page boundary shares everything
        assert eng.step_begin([2], [8]) == 0
        assert eng.seq_fork(2, 3) == 0
        assert list(eng.table(3)) == list(eng.table(2)) and refs(eng, 2) == [2, 2]
        assert eng.step_begin([2, 3], [1, 1]) == 0           # each gets its own new page
        assert eng.table(2)[2] != eng.table(3)[2]
        # truncating into a shared page is refused; rolling back own appends is fine
        assert eng.seq_truncate(3, 6) == pa.PA_ERR_UNSUPPORTED
        assert eng.seq_truncate(3, 8) == 0 and list(eng.table(3)) == list(eng.table(2))[:2]
        # errors
        assert eng.seq_fork(2, 3) == pa.PA_ERR_INVALID       # destination not empty
        assert eng.seq_fork(2, 2) == pa.PA_ERR_INVALID
    finally:
        eng.close()


def test_eviction_with_shared_pages_frees_only_unshared_ones():
    eng = make(bs=4, blocks=6, seqs=4)
    try:
        assert eng.step_begin([0], [8]) == 0                  # pages 0,1
        assert eng.seq_fork(0, 1) == 0                         # shared
        assert eng.step_begin([1], [4]) == 0                   # page 2 private to 1
        assert eng.step_begin([2], [12]) == 0                  # pages 3,4,5: pool full
        # the next allocation evicts the prompt owning the LRU page = prompt 0 (pages 0,1 were allocated
        # first); they are shared with prompt 1, so that frees nothing and the allocator goes on to the
        # next victim -- prompt 1, the pages' new owner -- until a page is free
        assert eng.step_begin([3], [4]) == 0
        assert eng.seq_len(0) == 0 and eng.seq_len(1) == 0 and eng.seq_len(2) == 12
        assert list(eng.table(3)) == [0]
        assert refs(eng, 6) == [1, 0, 0, 1, 1, 1]
    finally:
        eng.close()


def test_prefix_cache_insert_match_and_eviction():
    eng = make(bs=4, blocks=8, seqs=6)
    try:
        toks = np.arange(100, 114, dtype=np.int32)            # 14 tokens: 3 full pages + 2
        assert eng.step_begin([0], [14]) == 0
        assert eng.prefix_insert(0, toks) == 3
        assert eng.lib.pa_prefix_cached_pages(eng.h) == 3
        assert refs(eng, 4) == [2, 2, 2, 1]
        assert eng.prefix_insert(0, toks) == 0                 # idempotent
        # a new prompt with the same first 9 tokens reuses 2 pages (8 tokens), never the whole prompt
        other = np.concatenate([toks[:9], [7, 7, 7]]).astype(np.int32)
        assert eng.prefix_match(1, other) == 8
        assert list(eng.table(1)) == list(eng.table(0))[:2]
        assert eng.prefix_match(2, toks[:8]) == 4              # 8 tokens: only 4 may be reused (one page left to compute)
        assert eng.prefix_match(3, np.array([1, 2, 3, 4, 5], np.int32)) == 0
        assert eng.prefix_match(1, toks) == pa.PA_ERR_INVALID  # not empty
        # the cache keeps pages alive after every sequence is gone
        for s in (0, 1, 2):
            assert eng.seq_free(s) == 0
        assert refs(eng, 4) == [1, 1, 1, 0]
        assert eng.mgr.contents.blocks[0].prompt_id == -2
        assert eng.prefix_match(4, toks) == 12
        assert eng.seq_free(4) == 0
        # under pressure cached-only pages are dropped first, least recently used first
        assert eng.step_begin([5], [4 * 5]) == 0               # 5 free pages exactly: nothing evicted yet
        assert eng.lib.pa_prefix_cached_pages(eng.h) == 3
        assert eng.step_begin([5], [4]) == 0                   # needs one more: a cached page goes
        assert eng.lib.pa_prefix_cached_pages(eng.h) == 2
        assert eng.seq_len(5) == 24
    finally:
        eng.close()


def test_extensions_off_leave_refcounts_binary():
    eng = make(bs=4, blocks=5, seqs=3)
    try:
        rng = np.random.default_rng(0)
        for _ in range(200):
            s = int(rng.integers(0, 3))
            if rng.random() < 0.2:
                eng.seq_free(s)
            else:
                eng.step_begin([s], [int(rng.integers(1, 6))])
            used = 0
            for i in range(5):
                r = eng.lib.pa_page_refcount(eng.h, i)
                assert r in (0, 1)
                assert (r == 1) == (eng.mgr.contents.blocks[i].prompt_id != -1)
                used += r
            assert used == sum(len(eng.table(p)) for p in range(3))
    finally:
        eng.close()


def test_swap_out_on_eviction_and_swap_in_on_next_use():
    """With pa_set_evict_swap the allocator's LRU eviction keeps a host copy; the sequence returns with
    its length intact the next time a step names it (integer logic; the copies are device-side)."""
    eng = make(bs=4, blocks=6, seqs=4)
    lib = eng.lib
    try:
        assert lib.pa_set_evict_swap(eng.h, 1) == 0
        assert eng.step_begin([0], [8]) == 0        # pages 0,1
        assert eng.step_begin([1], [8]) == 0        # pages 2,3
        assert eng.step_begin([2], [8]) == 0        # pages 4,5: full
        assert eng.step_begin([1], [1]) == 0        # needs a page: the LRU prompt (0) is swapped out, not lost
        assert eng.seq_len(0) == 0 and lib.pa_seq_swapped_tokens(eng.h, 0) == 8
        assert eng.seq_len(1) == 9
        # naming sequence 0 again brings it back; whoever the reference's page-LRU picks as the next victim
        # (the prompt owning the oldest page: 1, whose first pages are older than 2's) is swapped out, not lost
        assert eng.step_begin([0], [1]) == 0
        assert eng.seq_len(0) == 9 and lib.pa_seq_swapped_tokens(eng.h, 0) == 0
        want = {0: 9, 1: 9, 2: 8}
        for s_, n in want.items():                    # nothing is ever lost: resident + swapped = everything appended
            assert eng.seq_len(s_) + lib.pa_seq_swapped_tokens(eng.h, s_) == n
            assert eng.seq_len(s_) == 0 or lib.pa_seq_swapped_tokens(eng.h, s_) == 0
        assert lib.pa_seq_swapped_tokens(eng.h, 1) == 9
        # explicit swap in / out
        assert lib.pa_seq_swap_in(eng.h, 1) == 0
        assert eng.seq_len(1) == 9
        for s_, n in want.items():
            assert eng.seq_len(s_) + lib.pa_seq_swapped_tokens(eng.h, s_) == n
        resident = [s_ for s_ in want if eng.seq_len(s_)]
        assert lib.pa_seq_swap_out(eng.h, resident[0]) == 0
        assert eng.seq_len(resident[0]) == 0 and lib.pa_seq_swapped_tokens(eng.h, resident[0]) == want[resident[0]]
        assert lib.pa_seq_swap_in(eng.h, resident[0]) == 0
        # switched off: the reference's behaviour (the victim is dropped)
        assert lib.pa_set_evict_swap(eng.h, 0) == 0
        before = {s_: lib.pa_seq_swapped_tokens(eng.h, s_) for s_ in want}
        assert eng.step_begin([3], [8]) == 0
        assert {s_: lib.pa_seq_swapped_tokens(eng.h, s_) for s_ in want} == before      # no new host copies
        assert sum(eng.seq_len(s_) + before[s_] for s_ in want) < sum(want.values())     # somebody was dropped
    finally:
        eng.close()
