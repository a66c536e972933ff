"""Integer parity of the product's host side (pa_block_manager.c, pa_step.c) -- block tables,
LRU counters, eviction order, slot mappings -- against the compiled reference (oracle/_ref) and
the oracle restatement, bit-exact.  CPU only: host-only handles, no compute."""
import numpy as np
import pytest

import __graft_entry__ as ge
import oracle_api as oa
from trace_driver import make_trace, run_trace

pa = ge.load_binding()

GEOMS = [(16, 12, 8), (4, 24, 6), (32, 100, 100), (8, 64, 16)]


def host_engine(bs, mb, mp, NH=2, hs=2, **kw):
    return pa.PagedAttn(bs, mb, mp, NH, hs, device=pa.PA_HOST_ONLY, **kw)


@pytest.mark.parametrize("bs,mb,mp", GEOMS)
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_trace_matches_reference_and_oracle(bs, mb, mp, seed):
    ops = make_trace(seed, 800, mp, bs, n_active=min(mp, 10))
    eng = host_engine(bs, mb, mp)
    orc = oa.OrcManager(4, bs, mb, mp, alloc_data=False)
    ref = oa.RefManager(4, bs, mb, mp) if oa.have_ref(bs, mb, mp) else None
    try:
        a = run_trace(pa.ManagerAdapter(eng.mgr), ops, mp, mb, snap_every=5)
        b = run_trace(orc, ops, mp, mb, snap_every=5)
        assert a == b
        if ref is not None:
            assert a == run_trace(ref, ops, mp, mb, snap_every=5)
    finally:
        eng.close(); orc.close()
        if ref is not None:
            ref.close()


def test_survey_scripted_trace_on_product():
    eng = host_engine(32, 100, 100)
    m = pa.ManagerAdapter(eng.mgr)
    try:
        for _ in range(3):
            for p in (0, 1, 2):
                m.request_block(p)
        assert [m.table(p) for p in (0, 1, 2)] == [[0, 3, 6], [1, 4, 7], [2, 5, 8]]
        assert m.epoch() == 9
        m.free_blocks_for_prompt(1)
        for _ in range(4):
            m.request_block(3)
        assert m.table(3) == [1, 4, 7, 9]
        got = [m.request_block(4) for _ in range(95)]
        assert got[90:] == [0, 3, 6, 2, 5]
        assert m.request_block(100) == -1 and m.request_block(-1) == -1
    finally:
        eng.close()


@pytest.mark.parametrize("bs,mb,mp", [(16, 64, 8), (4, 40, 6), (32, 100, 100)])
def test_step_tables_match_oracle(bs, mb, mp):
    """pa_step_begin = per-token page choice of add_to_cache; slot = table[pos/bs]*bs + pos%bs."""
    rng = np.random.default_rng(bs)
    eng = host_engine(bs, mb, mp, max_batch_tokens=4096)
    orc = oa.OrcManager(4, bs, mb, mp, alloc_data=False)
    try:
        nseq = min(mp, 6)
        for step in range(40):
            k = int(rng.integers(1, nseq + 1))
            seqs = rng.choice(nseq, size=k, replace=False).astype(np.int32)
            n_new = (rng.integers(1, 4, size=k) if step else rng.integers(1, 3 * bs, size=k)).astype(np.int32)
            budget_pages = sum((orc.context_len(int(s)) + int(n) + bs - 1) // bs for s, n in zip(seqs, n_new))
            others = sum(len(orc.table(p)) for p in range(nseq) if p not in seqs)
            if budget_pages + others > mb:
                victim = int(seqs[0])
                eng.seq_free(victim); orc.free_blocks_for_prompt(victim)
                continue
            assert eng.step_begin(seqs, n_new) == 0, pa.last_error()
            want_slots = []
            for s, n in zip(seqs, n_new):
                left = int(n)
                while left > 0:                  # one page choice per page touched (extension:
                    idx = orc.choose_page(int(s))  # the reference cannot cross a page, :542-545)
                    f = orc.block_info(idx)[0]
                    take = min(left, bs - f)
                    want_slots.extend(idx * bs + f + r for r in range(take))
                    orc.set_filled(idx, f + take)
                    left -= take
            assert eng.slot_mapping().tolist() == want_slots
            assert eng.context_lens().tolist() == [orc.context_len(int(s)) for s in seqs]
            tbl = eng.step_block_table()
            for i, s in enumerate(seqs):
                t = orc.table(int(s))
                assert tbl[i, :len(t)].tolist() == t
                assert eng.table(int(s)) == t
            m = pa.ManagerAdapter(eng.mgr)
            assert m.epoch() == orc.epoch()
            for s in seqs:                       # implied slot of every cached position
                L = orc.context_len(int(s))
                for pos in (0, L // 2, L - 1):
                    assert orc.slot(int(s), pos) == eng.table(int(s))[pos // bs] * bs + pos % bs
    finally:
        eng.close(); orc.close()


def test_multi_token_step_equals_reference_add_to_cache_when_it_fits():
    """n_new tokens that fit the current page = ONE add_to_cache(n_tail=n) of the reference: one
    page choice / LRU stamp, filled += n (paged_infer.c:518-529,570)."""
    bs, mb, mp = 32, 100, 100
    if not oa.have_ref(bs, mb, mp):
        pytest.skip("reference build absent")
    eng = host_engine(bs, mb, mp, max_batch_tokens=64)
    ref = oa.RefManager(4, bs, mb, mp)
    try:
        T = 32
        qkv = np.zeros((1, T, 12), dtype=np.float32)
        m = pa.ManagerAdapter(eng.mgr)
        for step, n in enumerate([32, 1, 1, 5, 1, 20, 1, 1, 1, 1]):
            ref.add_to_cache(qkv, 1, T, n)
            assert eng.step_begin([0], [n]) == 0
            assert ref.table(0) == eng.table(0), step
            assert ref.epoch() == m.epoch(), step
            assert [ref.block_info(i) for i in ref.table(0)] == [m.block_info(i) for i in eng.table(0)]
    finally:
        eng.close(); ref.close()


def test_eviction_drops_whole_prompt_and_context():
    bs, mb, mp = 4, 6, 4
    eng = host_engine(bs, mb, mp, max_batch_tokens=64)
    try:
        assert eng.step_begin([0, 1], [8, 8]) == 0          # 2 + 2 pages
        assert eng.step_begin([2], [8]) == 0                # 2 pages -> pool full
        assert eng.seq_len(0) == 8
        assert eng.step_begin([3], [1]) == 0                # evicts prompt 0 (oldest lru_counter)
        assert eng.seq_len(0) == 0 and eng.table(0) == []
        assert eng.table(3) == [0]
        assert eng.seq_len(1) == 8 and eng.seq_len(2) == 8
    finally:
        eng.close()


def test_truncate_adopt_and_errors():
    bs = 16
    eng = host_engine(bs, 32, 4, max_batch_tokens=256)
    try:
        assert eng.step_begin([0], [40]) == 0
        assert eng.table(0) == [0, 1, 2]
        assert eng.seq_truncate(0, 17) == 0 and eng.table(0) == [0, 1] and eng.seq_len(0) == 17
        assert eng.seq_truncate(0, 16) == 0 and eng.table(0) == [0] and eng.seq_len(0) == 16
        assert eng.seq_truncate(0, 99) == pa.PA_ERR_INVALID
        assert eng.step_begin([0], [20]) == 0 and eng.seq_len(0) == 36 and eng.table(0) == [0, 1, 2]
        assert eng.step_rollback() == 0 and eng.seq_len(0) == 16 and eng.table(0) == [0]
        assert eng.seq_adopt(1, [9, 3, 7], 33) == 0
        assert eng.table(1) == [9, 3, 7] and eng.seq_len(1) == 33
        assert eng.seq_adopt(2, [3], 5) == pa.PA_ERR_INVALID          # page in use
        assert eng.seq_adopt(2, [11, 12], 16) == pa.PA_ERR_INVALID    # 16 tokens need exactly 1 page
        assert eng.step_begin([7], [1]) == pa.PA_ERR_INVALID          # Invalid prompt ID
        assert eng.step_begin([0], [10_000]) == pa.PA_ERR_INVALID     # over max_batch_tokens
        # a per-sequence page cap fails like the reference's "No blocks available."
        capped = host_engine(bs, 8, 2, max_blocks_per_seq=2, max_batch_tokens=256)
        assert capped.step_begin([0], [33]) == pa.PA_ERR_NO_BLOCKS
        capped.close()
    finally:
        eng.close()
