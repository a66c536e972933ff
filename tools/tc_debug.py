"""Debug helper for the tcgen05 prefill kernel: dumps raw S / un-normalised O and compares with numpy."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pa = ge.build(quiet=True)
lib = pa.load()
NH, hs, bs, T = int(os.environ.get("NH", 1)), int(os.environ.get("HS", 64)), 16, int(os.environ.get("T", 128))
nwg = int(os.environ.get("NWG", 1))
C_ = NH * hs
eng = pa.PagedAttn(bs, T // bs + 8, 1, NH, hs, n_layers=1, device=0, max_batch_tokens=T + 8)
rng = np.random.default_rng(3)
qkv = rng.standard_normal((T, 3 * C_), dtype=np.float32)
d = pa.DevBuf.from_numpy(qkv)
o = pa.DevBuf(T * C_ * 4)
for dbg in (2, 0):
    eng.tune(pa.PA_TUNE_PREFILL_PATH, 3)
    eng.tune(pa.PA_TUNE_TC_WARPGROUPS, nwg)
    eng.tune(pa.PA_TUNE_TC_DEBUG, dbg)
    assert eng.step_begin([0], [T]) == 0
    pa.check(eng.upload(), "upload")
    pa.check(lib.pa_memset(o.ptr, 0, T * C_ * 4, None), "memset")
    pa.check(eng.append(0, d.ptr + C_ * 4, d.ptr + 2 * C_ * 4, 3 * C_), "append")
    pa.check(eng.prefill(0, d.ptr, 3 * C_, o.ptr, C_), "prefill")
    eng.sync()
    got = o.download((T, C_))
    pa.check(eng.step_rollback(), "rollback")
    q = qkv[:, :C_].reshape(T, NH, hs); k = qkv[:, C_:2*C_].reshape(T, NH, hs); v = qkv[:, 2*C_:].reshape(T, NH, hs)
    S = np.einsum("tnd,snd->nts", q, k)            # (NH, T, T)
    if dbg == 1:
        g = got.reshape(T, NH, hs)
        want = S[:, :128, :hs].transpose(1, 0, 2)
        n = min(T, 128)
        print("S dump: max|got|", np.abs(g[:n]).max(), "max|want|", np.abs(want).max(), "max diff", np.abs(g[:n] - want[:n]).max())
        print(" got[0,0,:8]", g[0, 0, :8]); print("want[0,0,:8]", want[0, 0, :8])
        print(" got[5,0,:8]", g[5, 0, :8]); print("want[5,0,:8]", want[5, 0, :8])
        # is got a permutation of want?  look for want[0,0,0] in got
    elif dbg == 3:
        sc = S / np.sqrt(hs)
        mask = np.tril(np.ones((T, T), dtype=bool))
        sc = np.where(mask[None], sc, -np.inf)
        m = np.maximum(sc.max(-1, keepdims=True), -10000.0)
        P = np.exp(sc - m)
        g = got.reshape(T, NH, hs)
        want = P[:, :128, :hs].transpose(1, 0, 2)
        print("P dump: max diff (cols < hs-1)", np.abs(g[:128, :, :hs-1] - want[:128, :, :hs-1]).max(), "psum got", g[[0, 5, 100], 0, hs-1], "want", P[0, [0, 5, 100], :128].sum(-1))
        print(" got[5,0,:8]", g[5, 0, :8]); print("want[5,0,:8]", want[5, 0, :8])
    else:
        sc = S / np.sqrt(hs)
        mask = np.tril(np.ones((T, T), dtype=bool))
        sc = np.where(mask[None], sc, -np.inf)
        m = np.maximum(sc.max(-1, keepdims=True), -10000.0)
        P = np.exp(sc - m)
        O = np.einsum("nts,snd->tnd", P, v)
        if dbg == 0:
            O = O / P.sum(-1).transpose(1, 0)[:, :, None]
        g = got.reshape(T, NH, hs)
        print("dbg", dbg, "max|got|", np.abs(g).max(), "max|want|", np.abs(O).max(), "max diff", np.abs(g - O).max())
        print(" got[5,0,:6]", g[5, 0, :6]); print("want[5,0,:6]", O[5, 0, :6])
        print(" got[100,0,:6]", g[100, 0, :6]); print("want[100,0,:6]", O[100, 0, :6])
eng.close()
