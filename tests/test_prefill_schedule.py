"""The work list of the persistent tensor-core prefill kernel (pa_prefill_tc3.cu: make_schedule, exported as
pa_prefill_schedule) is pure host code: checked here without a GPU.  Every (sequence, q tile of 128 rows, head)
unit must appear exactly once, the columns (one CTA each) must carry about the same number of key tiles, and the
q tiles of one (sequence, head) must sit next to each other inside a length class (L2 sharing of K/V)."""
import ctypes as C

import numpy as np
import pytest

import __graft_entry__ as ge

pa = ge.load_binding()


def _schedule(eng, key_tile, max_ctas=148):
    lib = eng.lib
    n_ctas, rows = C.c_int(0), C.c_int(0)
    pa.check(lib.pa_prefill_schedule(eng.h, key_tile, max_ctas, None, 0, C.byref(n_ctas), C.byref(rows)), "sizes")
    units = np.full((rows.value * n_ctas.value, 2), -7, dtype=np.int32)
    pa.check(lib.pa_prefill_schedule(eng.h, key_tile, max_ctas, units.ctypes.data_as(C.POINTER(C.c_int)), len(units),
                                     C.byref(n_ctas), C.byref(rows)), "schedule")
    return units.reshape(rows.value, n_ctas.value, 2)


def _key_tiles(n_new, before, kv_start, qt, BN):
    """key tiles unit (sequence, q tile qt) walks: the kernel's geo()"""
    j0 = qt * 128
    rows = min(128, n_new - j0)
    kv_end = before + n_new
    lim_last = kv_end - (n_new - 1 - (j0 + rows - 1))
    k_begin = (kv_start // BN) * BN
    return (lim_last - k_begin + BN - 1) // BN if lim_last > k_begin else 0


def _check(NH, hs, bs, before, n_new, kv_start=None, max_ctas=148, min_balance=None):
    B = len(n_new)
    BN = 64 if hs == 64 else 32
    pages = sum((b + n + bs - 1) // bs + 1 for b, n in zip(before, n_new)) + 8
    eng = pa.PagedAttn(bs, pages, B, NH, hs, device=pa.PA_HOST_ONLY, max_batch_tokens=sum(n_new) + sum(before) + 8)
    try:
        ids = list(range(B))
        if any(before):
            assert eng.step_begin(ids, before) == 0, pa.last_error()
        assert eng.step_begin(ids, n_new) == 0, pa.last_error()
        if kv_start is not None:
            assert eng.step_set_kv_start(kv_start) == 0, pa.last_error()
        ks = kv_start or [0] * B
        sched = _schedule(eng, BN, max_ctas)
        rows, n_ctas, _ = sched.shape
        want = {(s, qt, h) for s in range(B) for qt in range((n_new[s] + 127) // 128) for h in range(NH)}
        assert n_ctas == min(len(want), max_ctas) and rows == (len(want) + n_ctas - 1) // n_ctas
        seen = set()
        load = np.zeros(n_ctas)
        for r in range(rows):
            for c in range(n_ctas):
                s, y = int(sched[r, c, 0]), int(sched[r, c, 1])
                if s < 0:
                    assert r == rows - 1, "padding only in the last row"
                    continue
                u = (s, y & 0xffff, y >> 16)
                assert u in want and u not in seen, u
                seen.add(u)
                load[c] += _key_tiles(n_new[s], before[s], ks[s], u[1], BN) + 1.5
        assert seen == want
        balance = load.mean() / load.max()
        if min_balance is not None:
            assert balance >= min_balance, f"balance {balance:.3f}"
        return sched, balance
    finally:
        eng.close()


def test_every_unit_once_small_and_ragged():
    _check(3, 64, 16, [0, 100, 0, 17, 300], [300, 129, 128, 1, 257])
    _check(2, 128, 16, [0, 77, 0], [200, 65, 64])
    _check(1, 64, 8, [0], [1])                                           # one unit, one CTA
    _check(4, 64, 16, [0, 0], [300, 450], kv_start=[200, 129])           # q tiles without a visible key are still listed
    rng = np.random.default_rng(3)
    for _ in range(10):
        B = int(rng.integers(1, 30))
        _check(int(rng.integers(1, 13)), int(rng.choice([64, 128])), 16, [int(x) for x in rng.integers(0, 400, B)],
               [int(x) for x in rng.integers(1, 900, B)], max_ctas=int(rng.choice([1, 7, 148, 160])))


@pytest.mark.parametrize("name,NH,hs,before,n_new,floor", [
    ("16 x 2048", 12, 64, [0] * 16, [2048] * 16, 0.97),
    ("16 x 1024", 12, 64, [0] * 16, [1024] * 16, 0.95),
    ("2 x 2048 on 30000 cached, head_dim 128", 32, 128, [30000] * 2, [2048] * 2, 0.97),
    ("8 x 4096, head_dim 128", 32, 128, [0] * 8, [4096] * 8, 0.97),
    ("ragged", 12, 64, [0] * 24, [int(x) for x in np.random.default_rng(0).integers(1, 4096, 24)], 0.95),
])
def test_columns_balance(name, NH, hs, before, n_new, floor):
    _, balance = _check(NH, hs, 16, before, n_new, min_balance=floor)
    print(f"{name}: balance {balance:.3f}")


def test_q_tiles_of_a_sequence_and_head_are_neighbours():
    """In deal order (boustrophedon undone) a (sequence, head)'s q tiles of one length class are consecutive."""
    sched, _ = _check(12, 64, 16, [0] * 16, [2048] * 16)
    rows, n_ctas, _ = sched.shape
    order = []
    for r in range(rows):
        cols = range(n_ctas) if r % 2 == 0 else range(n_ctas - 1, -1, -1)
        order += [(int(sched[r, c, 0]), int(sched[r, c, 1]) & 0xffff, int(sched[r, c, 1]) >> 16) for c in cols if sched[r, c, 0] >= 0]
    runs = same = 0
    for a, b in zip(order, order[1:]):
        runs += 1
        same += (a[0], a[2]) == (b[0], b[2])
    # 16 q tiles in 8 classes: runs of two neighbours -> at least ~half of the adjacent pairs share (sequence, head)
    assert same / runs >= 0.45, same / runs
    # and the first class listed is the longest q tiles
    assert order[0][1] == 15


def test_bad_arguments():
    eng = pa.PagedAttn(16, 64, 2, 2, 64, device=pa.PA_HOST_ONLY, max_batch_tokens=256)
    try:
        n, r = C.c_int(0), C.c_int(0)
        assert eng.lib.pa_prefill_schedule(eng.h, 64, 148, None, 0, C.byref(n), C.byref(r)) != 0          # no step yet
        assert eng.step_begin([0, 1], [130, 5]) == 0
        assert eng.lib.pa_prefill_schedule(eng.h, 48, 148, None, 0, C.byref(n), C.byref(r)) != 0          # key tile 32 or 64
        assert eng.lib.pa_prefill_schedule(eng.h, 64, 148, None, 0, C.byref(n), C.byref(r)) == 0
        assert (n.value, r.value) == (6, 1)                                                              # 3 q tiles x 2 heads
        small = np.zeros((2, 2), dtype=np.int32)
        assert eng.lib.pa_prefill_schedule(eng.h, 64, 148, small.ctypes.data_as(C.POINTER(C.c_int)), 2, C.byref(n), C.byref(r)) != 0
    finally:
        eng.close()
