"""first-slab latency of the decode-size projection against the weight row stride (PA_GEMM_DEBUG=1 prints the CTA-0 timeline)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pa = ge.build(quiet=True)
lib = pa.load()
rng = np.random.default_rng(0)
flush = pa.DevBuf(512 << 20)
for (K, N) in ((3072, 768), (3104, 768), (3080, 768), (2048, 768), (4096, 768), (768, 3072), (768, 2304), (1600, 1600), (6400, 1600), (6432, 1600)):
    M = 64
    x = rng.standard_normal((M, K), dtype=np.float32)
    w = (rng.standard_normal((N, K), dtype=np.float32) / np.sqrt(K)).astype(np.float32)
    dx, dw, do = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf(M * N * 4)
    for rep in range(3):
        pa.check(lib.pa_memset(flush.ptr, rep, 512 << 20, None), "flush")
        pa.check(lib.pa_device_sync(), "sync")
        sys.stderr.write(f"K={K} N={N} rep={rep}: "); sys.stderr.flush()
        pa.check(lib.pa_matmul_bias(dx.ptr, K, dw.ptr, None, do.ptr, N, M, N, K, None), "gemm")
        pa.check(lib.pa_device_sync(), "sync")
    dx.free(); dw.free(); do.free()
