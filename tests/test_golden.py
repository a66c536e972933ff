"""Committed golden vectors (tests/golden/, generated from the compiled reference by
tests/golden/make_golden.py).  CPU tests pin the oracle and the product's block manager to
them; the gpu-marked tests pin the CUDA path.  None of this reads /root/reference."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import __graft_entry__ as ge
import oracle_api as oa
from trace_driver import run_trace

pa = ge.load_binding()
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _tuplify(x):
    if isinstance(x, list):
        return [_tuplify(v) for v in x]
    return x


def _norm(log):
    """JSON turns tuples into lists; normalise both sides."""
    return json.loads(json.dumps(log))


def test_golden_trace_oracle_and_product():
    g = json.load(open(os.path.join(GOLD, "golden_trace.json")))
    bs, mb, mp = g["geom"]
    ops = [tuple(e["op"]) for e in g["log"]]
    orc = oa.OrcManager(4, bs, mb, mp, alloc_data=False)
    eng = pa.PagedAttn(bs, mb, mp, 2, 2, device=pa.PA_HOST_ONLY)
    try:
        assert _norm(run_trace(orc, ops, mp, mb, snap_every=10)) == g["log"]
        assert _norm(run_trace(pa.ManagerAdapter(eng.mgr), ops, mp, mb, snap_every=10)) == g["log"]
    finally:
        orc.close(); eng.close()


def _fill(mgr, kv, bs):
    for t0 in range(0, kv.shape[0], bs):
        idx = mgr.request_block(0)
        k, v = mgr.page_arrays(idx)
        n = min(bs, kv.shape[0] - t0)
        k[:n], v[:n] = kv[t0:t0 + n, 0], kv[t0:t0 + n, 1]
        mgr.set_filled(idx, n)


@pytest.mark.parametrize("name", ["golden_attn_bs16.npz", "golden_attn_bs2.npz"])
def test_golden_attention_oracle_bit_exact(name):
    g = np.load(os.path.join(GOLD, name))
    bs, mb, mp, T, Cc, NH, offset = [int(x) for x in g["geom"]]
    orc = oa.OrcManager(Cc, bs, mb, mp)
    try:
        _fill(orc, g["kv"], bs)
        _, out, pre, att = orc.attend(0, g["inp"], 1, T, NH, offset, want_scratch=True)
        assert np.array_equal(out.view(np.uint32), g["out"].view(np.uint32))
        assert np.array_equal(att.view(np.uint32), g["att"].view(np.uint32))
    finally:
        orc.close()


def test_golden_window_oracle_bit_exact():
    g = np.load(os.path.join(GOLD, "golden_window_bs32.npz"))
    bs, mb, mp, T, Cc, NH = [int(x) for x in g["geom"]]
    orc = oa.OrcManager(Cc, bs, mb, mp)
    try:
        for step in range(19):
            window = np.ascontiguousarray(g["stream"][step:step + T][None])
            orc.add_to_cache(window, 1, T, T if step == 0 else 1)
            _, out = orc.attend(0, window, 1, T, NH, step)
            assert np.array_equal(out[0].view(np.uint32), g["outs"][step].view(np.uint32)), step
            tbl = orc.table(0)
            assert tbl + [-1] * (2 - len(tbl)) == g["tables"][step].tolist()
            assert orc.epoch() == int(g["epochs"][step])
        assert [orc.block_info(i)[0] for i in orc.table(0)] == g["filled"].tolist()
    finally:
        orc.close()


# ----------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_golden_window_cuda_compat_path():
    from gpu_common import assert_close
    lib = pa.load()
    g = np.load(os.path.join(GOLD, "golden_window_bs32.npz"))
    bs, mb, mp, T, Cc, NH = [int(x) for x in g["geom"]]
    lib.pa_set_default_geometry(bs, mb, mp)
    m = lib.create_block_manager(Cc)
    assert m
    try:
        ad = pa.ManagerAdapter(m)
        for step in range(19):
            window = np.ascontiguousarray(g["stream"][step:step + T][None])
            lib.add_to_cache(m, window.ctypes.data, 1, T, Cc, T if step == 0 else 1)
            nb = C.c_int()
            kv = lib.collect_kv_blocks(m, 0, C.byref(nb))
            out = np.zeros((1, T, Cc), dtype=np.float32)
            lib.attention_paged(out.ctypes.data, None, None, window.ctypes.data, kv[0], kv[1], 1, T, Cc, NH, step)
            assert_close(out[0], g["outs"][step], f"golden window step {step}")
            tbl = ad.table(0)
            assert tbl + [-1] * (2 - len(tbl)) == g["tables"][step].tolist()       # bit-exact block table
            assert m.contents.lru_epoch == int(g["epochs"][step])
        assert [ad.block_info(i)[0] for i in ad.table(0)] == g["filled"].tolist()
    finally:
        lib.destroy_block_manager(m)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["golden_attn_bs16.npz", "golden_attn_bs2.npz"])
def test_golden_attention_cuda(name):
    """Full-window rows through pa_prefill-style raw step (compat attention_paged) and the decode
    kernels for the last row."""
    from gpu_common import assert_close
    lib = pa.load()
    g = np.load(os.path.join(GOLD, name))
    bs, mb, mp, T, Cc, NH, offset = [int(x) for x in g["geom"]]
    lib.pa_set_default_geometry(bs, mb, mp)
    m = lib.create_block_manager(Cc)
    assert m
    try:
        kv = g["kv"]
        for t0 in range(0, kv.shape[0], bs):
            blk = lib.request_block(m, 0)
            n = min(bs, kv.shape[0] - t0)
            k_rows = np.ascontiguousarray(kv[t0:t0 + n, 0]); v_rows = np.ascontiguousarray(kv[t0:t0 + n, 1])
            pa.check(lib.pa_memcpy_h2d(blk.contents.keys, k_rows.ctypes.data, k_rows.nbytes, None), "h2d")
            pa.check(lib.pa_memcpy_h2d(blk.contents.values, v_rows.ctypes.data, v_rows.nbytes, None), "h2d")
            blk.contents.filled = n
        nb = C.c_int()
        kvp = lib.collect_kv_blocks(m, 0, C.byref(nb))
        out = np.zeros((1, T, Cc), dtype=np.float32)
        inp = np.ascontiguousarray(g["inp"])
        lib.attention_paged(out.ctypes.data, None, None, inp.ctypes.data, kvp[0], kvp[1], 1, T, Cc, NH, offset)
        assert_close(out, g["out"], name)
    finally:
        lib.destroy_block_manager(m)
        lib.pa_set_default_geometry(32, 100, 100)
