"""A step never evicts its own sequences (pa_step.c): the whole-prompt LRU eviction inside
request_block (block_manager.c:104-113) may take any sequence OUTSIDE the step being built; when only
sequences of the step are left the call fails with PA_ERR_NO_BLOCKS and leaves no trace.  Host-only
handles (integer logic); the reference itself has no multi-sequence step to compare with -- the
single-prompt behaviour (tests/test_block_manager.py) is unchanged because nothing is pinned there."""
import numpy as np

import __graft_entry__ as ge

pa = ge.load_binding()


def make(bs=4, blocks=4, seqs=4, tokens=64):
    return pa.PagedAttn(bs, blocks, seqs, 2, 8, device=pa.PA_HOST_ONLY, max_batch_tokens=tokens)


def state(eng, seqs):
    return [(eng.seq_len(s), list(eng.table(s))) for s in range(seqs)]


def test_exhausted_pool_does_not_evict_a_sequence_of_the_step():
    # the advisor's reproduction: bs=4, 4 pages; seq0 holds 6 tokens (pages 0,1), seq1 holds 8 (pages 2,3)
    eng = make()
    try:
        assert eng.step_begin([0], [6]) == 0 and eng.step_begin([1], [8]) == 0
        before = state(eng, 2)
        # seq1 needs a new page; the only candidates are seq0 (in the step) and seq1 itself
        rc = eng.step_begin([0, 1], [1, 1])
        assert rc == pa.PA_ERR_NO_BLOCKS and "No blocks available" in pa.last_error()
        assert state(eng, 2) == before                      # seq0's already placed token was taken back
        assert eng.step_rollback() == pa.PA_ERR_INVALID     # there is no step
        # the same request with an outsider to evict goes through and evicts the outsider only
        eng.seq_free(0)
        assert eng.step_begin([2], [6]) == 0                # pages 0,1 now belong to seq2 (older than seq1's last touch?)
        assert eng.step_begin([1], [0]) == 0                # no-op step (touches nothing)
        rc = eng.step_begin([1], [1])
        assert rc == 0
        assert eng.seq_len(2) == 0 and eng.seq_len(1) == 9  # seq2 was the LRU outsider
        ctx = eng.context_lens()
        assert list(ctx) == [9]
    finally:
        eng.close()


def test_victim_is_the_lru_sequence_outside_the_step():
    eng = make(bs=4, blocks=6, seqs=4)
    try:
        assert eng.step_begin([0], [8]) == 0    # pages 0,1  (oldest)
        assert eng.step_begin([1], [8]) == 0    # pages 2,3
        assert eng.step_begin([2], [8]) == 0    # pages 4,5
        # a step naming the two OLDEST sequences: the pool is full, the victim must be seq2 (the newest!)
        assert eng.step_begin([0, 1], [1, 1]) == 0
        assert eng.seq_len(2) == 0
        assert eng.seq_len(0) == 9 and eng.seq_len(1) == 9
        slots = eng.slot_mapping()
        t0, t1 = list(eng.table(0)), list(eng.table(1))
        assert list(slots) == [t0[2] * 4, t1[2] * 4]
        assert sorted(t0 + t1) == [0, 1, 2, 3, 4, 5]
        # nothing stays pinned after the step: a later single-sequence request evicts by plain LRU again
        m = eng.mgr.contents
        assert not any(m.pinned[i] for i in range(4))
    finally:
        eng.close()


def test_failed_step_rolls_back_every_sequence_it_touched():
    eng = make(bs=4, blocks=5, seqs=4)
    try:
        assert eng.step_begin([0], [7]) == 0    # pages 0,1
        assert eng.step_begin([1], [8]) == 0    # pages 2,3
        before = state(eng, 3)
        free_before = sum(1 for i in range(5) if eng.mgr.contents.blocks[i].prompt_id == -1)
        # seq0 takes 1 (fits page 1), seq2 takes the free page, seq1 then finds nothing: all three are in the step
        rc = eng.step_begin([0, 2, 1], [1, 3, 1])
        assert rc == pa.PA_ERR_NO_BLOCKS
        assert state(eng, 3) == before
        assert sum(1 for i in range(5) if eng.mgr.contents.blocks[i].prompt_id == -1) == free_before
        assert eng.mgr.contents.blocks[1].filled == 3
    finally:
        eng.close()


def test_per_sequence_cap_failure_also_rolls_back():
    eng = pa.PagedAttn(4, 16, 2, 2, 8, device=pa.PA_HOST_ONLY, max_batch_tokens=64, max_blocks_per_seq=2)
    try:
        assert eng.step_begin([0], [5]) == 0
        assert eng.step_begin([1], [2]) == 0
        before = state(eng, 2)
        assert eng.step_begin([1, 0], [3, 4]) == pa.PA_ERR_NO_BLOCKS      # seq0 would need a third page
        assert state(eng, 2) == before
    finally:
        eng.close()


def test_a_sequence_may_appear_once_per_step():
    eng = make(blocks=8)
    try:
        assert eng.step_begin([0, 0], [1, 1]) == pa.PA_ERR_INVALID
        assert "twice" in pa.last_error()
        assert eng.seq_len(0) == 0
        assert eng.step_begin([0, 1], [1, 1]) == 0           # and nothing stayed pinned
        assert eng.step_begin_readonly([1, 1]) == pa.PA_ERR_INVALID
    finally:
        eng.close()


def test_fork_never_evicts_its_source():
    eng = make(bs=4, blocks=3, seqs=3)
    try:
        assert eng.step_begin([0], [10]) == 0                # pages 0,1,2 -- the pool is full, last page partial
        before = state(eng, 1)
        assert eng.seq_fork(0, 1) == pa.PA_ERR_NO_BLOCKS     # the copy of the partial page would need seq0's own pages
        assert state(eng, 1) == before and eng.seq_len(1) == 0
        assert [eng.lib.pa_page_refcount(eng.h, i) for i in range(3)] == [1, 1, 1]
    finally:
        eng.close()


def test_swapped_in_sequence_is_not_evicted_by_a_later_row_of_the_same_step():
    eng = make(bs=4, blocks=4, seqs=4)
    try:
        assert eng.lib.pa_set_evict_swap(eng.h, 1) == 0
        assert eng.step_begin([0], [8]) == 0                 # pages 0,1
        assert eng.lib.pa_seq_swap_out(eng.h, 0) == 0        # host copy, pages free
        assert eng.step_begin([1], [8]) == 0                 # pages 0,1 now seq1's
        assert eng.step_begin([2], [3]) == 0                 # page 2 (one row left)
        # the step brings seq0 back (needs 2 pages: evicts -- swaps out -- seq1, the LRU outsider) and lets seq2 grow
        assert eng.step_begin([0, 2], [1, 1]) == 0
        assert eng.seq_len(0) == 9 and eng.seq_len(2) == 4 and eng.seq_len(1) == 0
        assert eng.lib.pa_seq_swapped_tokens(eng.h, 1) == 8
        assert list(eng.context_lens()) == [9, 4]
    finally:
        eng.close()


def test_step_validate_catches_tables_that_point_outside_the_pool():
    """pa_step_validate (run by pa_step_upload under PA_VALIDATE_STEP=1, which the suite sets): every page index and
    slot the kernels will use lies inside the pool, windows and prefix sums are consistent."""
    import ctypes as C
    eng = make(bs=4, blocks=8, seqs=3)
    try:
        assert eng.step_begin([0, 1], [6, 3]) == 0
        assert eng.lib.pa_step_validate(eng.h) == 0
        n, stride = C.c_int(), C.c_int()
        tbl = eng.lib.pa_step_block_table(eng.h, C.byref(n), C.byref(stride))
        old = tbl[1]
        tbl[1] = 8                                   # one past the last page
        assert eng.lib.pa_step_validate(eng.h) == pa.PA_ERR_INVALID and "outside the pool" in pa.last_error()
        tbl[1] = old
        nt = C.c_int()
        slots = eng.lib.pa_step_slot_mapping(eng.h, C.byref(nt))
        old = slots[0]
        slots[0] = 8 * 4
        assert eng.lib.pa_step_validate(eng.h) == pa.PA_ERR_INVALID and "slot" in pa.last_error()
        slots[0] = old
        ctx = eng.lib.pa_step_context_lens(eng.h, C.byref(n))
        old = ctx[0]
        ctx[0] = 4 * 4 * 4                           # more tokens than the table row has pages for
        assert eng.lib.pa_step_validate(eng.h) == pa.PA_ERR_INVALID
        ctx[0] = old
        assert eng.lib.pa_step_validate(eng.h) == 0
    finally:
        eng.close()
