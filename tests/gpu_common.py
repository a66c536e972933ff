"""Shared helpers for the -m gpu parity tests: build a device pool + block tables through the
C ABI, mirror the same logical K/V content into the CPU oracle, compare."""
import numpy as np

import __graft_entry__ as ge
import oracle_api as oa

pa = ge.load_binding()

# Tolerance for fp32 attention outputs (north_star: "max relative error 1e-5"):
#   max|a-b| / max|ref| <= 1e-5   and   allclose(rtol=1e-5, atol=1e-6)
REL_TOL = 1e-5
RTOL, ATOL = 1e-5, 1e-6
# Tolerance of the opt-in tensor-core prefill (PA_TUNE_PREFILL_PATH=3): Q, K, V and P are rounded
# to TF32 (10-bit mantissa, relative step 2^-10 ~ 1e-3) before the fp32-accumulated products and
# exp is ex2.approx, so for N(0,1) inputs the outputs carry ~1e-3 of the largest |reference|:
#   max|a-b| / max|ref| <= 5e-3
TC_REL_TOL = 5e-3


def assert_close(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    scale = max(np.abs(want).max(), 1e-30)
    err = np.abs(got - want).max() / scale
    assert err <= REL_TOL, f"{what}: max|a-b|/max|ref| = {err:.3e} > {REL_TOL}"
    bad = np.abs(got - want) > ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} elements outside allclose(rtol={RTOL}, atol={ATOL}); worst {np.abs(got - want).max():.3e}"
    return err


# Tolerance of the fp32-accurate tensor-core prefill (PA_TUNE_PREFILL_PATH=4, and the automatic choice for large
# steps): the north-star bar max|a-b| / max|ref| <= 1e-5 unchanged; element-wise allclose(rtol=1e-5, atol=3e-6):
# the 3xTF32 split drops terms of 2^-21 relative per product and the tensor core's fp32 accumulate truncates, which
# leaves ~1e-6 of the largest |reference| on elements near zero (measured worst case 2.4e-6 absolute on values ~4).
TC3_ATOL = 3e-6


def assert_close_tc3(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    assert err <= REL_TOL, f"{what}: max|a-b|/max|ref| = {err:.3e} > {REL_TOL}"
    bad = np.abs(got - want) > TC3_ATOL + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()} elements outside allclose(rtol={RTOL}, atol={TC3_ATOL}); worst {np.abs(got - want).max():.3e}"
    return err


def assert_close_tc(got, want, what=""):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    assert err <= TC_REL_TOL, f"{what}: max|a-b|/max|ref| = {err:.3e} > {TC_REL_TOL} (TF32 tolerance)"
    return err


def assert_close_illconditioned(got, want32, want64, what="", floor=1e-2):
    """Inputs in the reference test's range U[0,100) (test_paged_attn.c:201) give logits of order
    1e4 whose fp32 ulp is ~1e-3, so ANY evaluation order moves the softmax weights by ~1e-3 and
    the fp32 reference is itself only defined to about that.  Bar: no further from the fp64
    truth than 4x the fp32 reference is, with the reference test's own absolute tolerance
    (1e-2 on values in [0,100), test_paged_attn.c:8) as the floor."""
    got = np.asarray(got, dtype=np.float64)
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    ref_err = np.abs(np.asarray(want32, dtype=np.float64) - want64).max()
    err = np.abs(got - want64).max()
    assert err <= max(4 * ref_err, floor), f"{what}: |got-truth|={err:.3e}, |ref32-truth|={ref_err:.3e}"
    return err


class Scenario:
    """A batch of sequences with given context lengths on one engine + the oracle twin."""

    def __init__(self, NH, hs, bs, ctx, n_layers=1, layer=0, seed=1234, shuffle=False, extra_blocks=8,
                 dist="normal", device=0, max_batch_tokens=0):
        self.NH, self.hs, self.bs, self.C = NH, hs, bs, NH * hs
        self.ctx = list(ctx)
        self.B = len(ctx)
        self.layer = layer
        pages = [(c + bs - 1) // bs for c in ctx]
        self.max_blocks = sum(pages) + extra_blocks + self.B
        self.eng = pa.PagedAttn(bs, self.max_blocks, self.B, NH, hs, n_layers=n_layers, device=device,
                                max_batch_tokens=max(max(ctx), self.B, max_batch_tokens) + 8)
        self.orc = oa.OrcManager(self.C, bs, self.max_blocks, self.B)
        rng = np.random.default_rng(seed)
        # whole pool content in one upload
        shape = (self.max_blocks * bs, self.C)
        if dist == "normal":
            pool_k = oa.normal(shape, seed=seed)
            pool_v = oa.normal(shape, seed=seed + 1)
        else:   # the reference test's value range, test_paged_attn.c:201
            pool_k = oa.uniform(shape, 0.0, 100.0, seed=seed)
            pool_v = oa.uniform(shape, 0.0, 100.0, seed=seed + 1)
        self.pool_k, self.pool_v = pool_k, pool_v
        lib = self.eng.lib
        pa.check(lib.pa_memcpy_h2d(self.eng.pool_k(layer), pool_k.ctypes.data, pool_k.nbytes, None), "h2d")
        pa.check(lib.pa_memcpy_h2d(self.eng.pool_v(layer), pool_v.ctypes.data, pool_v.nbytes, None), "h2d")
        # block tables: allocator order, or a seeded permutation dealt round-robin (fragmented)
        if shuffle:
            perm = rng.permutation(self.max_blocks)
            cur = 0
            for s in range(self.B):
                if ctx[s] == 0:
                    continue
                blocks = perm[cur:cur + pages[s]]
                cur += pages[s]
                assert self.eng.seq_adopt(s, blocks, ctx[s]) == 0, pa.last_error()
        else:
            for s in range(self.B):
                if ctx[s] > 0:
                    assert self.eng.step_begin([s], [ctx[s]]) == 0, pa.last_error()
        # oracle twin: same logical content
        for s in range(self.B):
            tbl = self.eng.table(s)
            for j, idx in enumerate(tbl):
                oidx = self.orc.request_block(s)
                k, v = self.orc.page_arrays(oidx)
                n = min(bs, ctx[s] - j * bs)
                k[:n] = pool_k[idx * bs: idx * bs + n]
                v[:n] = pool_v[idx * bs: idx * bs + n]
                self.orc.set_filled(oidx, n)
        self.seq_ids = list(range(self.B))

    def close(self):
        self.eng.close()
        self.orc.close()

    def decode(self, q, kv_start=None, path=0, hpg=0, stages=0, grid=0, static_pct=0, dyn_units=0):
        """Run pa_decode over all sequences (read-only step) and return (B, C)."""
        eng = self.eng
        eng.tune(pa.PA_TUNE_DECODE_PATH, path)
        eng.tune(pa.PA_TUNE_HEADS_PER_TILE, hpg)
        eng.tune(pa.PA_TUNE_STAGES, stages)
        eng.tune(pa.PA_TUNE_GRID, grid)
        eng.tune(pa.PA_TUNE_STATIC_PCT, static_pct)
        eng.tune(pa.PA_TUNE_DYN_UNITS, dyn_units)
        assert eng.step_begin_readonly(self.seq_ids) == 0, pa.last_error()
        if kv_start is not None:
            assert eng.step_set_kv_start(kv_start) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        d_q = pa.DevBuf.from_numpy(q)
        d_out = pa.DevBuf(self.B * self.C * 4)
        pa.check(eng.lib.pa_memset(d_out.ptr, 0xff, self.B * self.C * 4, None), "memset")   # NaN canary
        pa.check(eng.decode(self.layer, d_q.ptr, q.shape[1], d_out.ptr, self.C), "decode")
        eng.sync()
        out = d_out.download((self.B, self.C))
        d_q.free(); d_out.free()
        return out

    def oracle_decode_f64(self, q, kv_start=None):
        return self.orc.decode_batch_f64(self.seq_ids, self.NH, q[:, :self.C], kv_start=kv_start)

    def oracle_decode(self, q, kv_start=None):
        return self.orc.decode_batch(self.seq_ids, self.NH, q[:, :self.C], kv_start=kv_start)
