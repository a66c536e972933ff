"""Error structure of the 3xTF32 tensor-core GEMM against fp64 (debug helper)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pa = ge.build(quiet=True)
lib = pa.load()
rng = np.random.default_rng(0)
for K in (768, 1600, 4096):
    M, N = 128, 256
    x = rng.standard_normal((M, K), dtype=np.float32)
    w = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    dx, dw, do = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf(M * N * 4)
    pa.check(lib.pa_matmul_bias(dx.ptr, K, dw.ptr, None, do.ptr, N, M, N, K, None), "gemm")
    pa.check(lib.pa_device_sync(), "sync")
    got = do.download((M, N)).astype(np.float64)
    ref = x.astype(np.float64) @ w.astype(np.float64).T
    ref32 = (x @ w.T).astype(np.float64)
    e = got - ref
    print(f"K={K}: max|err|/max|ref| = {np.abs(e).max()/np.abs(ref).max():.2e}  numpy-fp32: {np.abs(ref32-ref).max()/np.abs(ref).max():.2e}  "
          f"mean(err*sign(ref))/max|ref| = {(e*np.sign(ref)).mean()/np.abs(ref).max():.2e}  rms = {np.sqrt((e**2).mean())/np.abs(ref).max():.2e}  "
          f"corr(err, ref) = {np.corrcoef(e.ravel(), ref.ravel())[0,1]:.3f}")
