#!/usr/bin/env python
"""SURVEY 8f.1 measurement: QKV projection with the KV append fused into its epilogue
(pa_qkv_append), alone and followed by the paged decode kernel, next to the reference's
matmul_cached + add_to_cache on the host cores (oracle port, bounded sample).

  python tools/qkv_bench.py [--shape 124m|xl] [--B n] [--ctx n] [--iters n]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402

SHAPES = {"124m": (12, 64), "xl": (25, 64)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="124m", choices=sorted(SHAPES))
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=576)
    ap.add_argument("--bs", type=int, default=16)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--split", type=int, default=0, help="tensor-core GEMM: K splits per cluster")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 fp32 SIMT, 2 tcgen05 3xTF32, 3 tcgen05 TF32")
    args = ap.parse_args()
    pa = ge.build(quiet=True)
    lib = pa.load()
    if lib.pa_device_count() < 1:
        raise SystemExit("qkv_bench: no CUDA device; libpaged_attn has no CPU fallback")
    NH, hs = SHAPES[args.shape]
    C_ = NH * hs
    B, bs, ctx = args.B, args.bs, args.ctx
    pages = (ctx + bs - 1) // bs + 1
    eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=1, device=0, max_batch_tokens=B)
    eng.tune(pa.PA_TUNE_GEMM_PATH, args.path)
    eng.tune(pa.PA_TUNE_GEMM_SPLIT_K, args.split)
    rng = np.random.default_rng(11)
    perm = rng.permutation(B * pages + 8)
    for s in range(B):
        assert eng.seq_adopt(s, perm[s * pages: s * pages + (ctx - 1 + bs - 1) // bs], ctx - 1) == 0, pa.last_error()
    x = rng.standard_normal((B, C_), dtype=np.float32)
    w = (rng.standard_normal((3 * C_, C_), dtype=np.float32) / np.sqrt(C_)).astype(np.float32)
    bias = rng.standard_normal((3 * C_,), dtype=np.float32)
    dx, dw, db = pa.DevBuf.from_numpy(x), pa.DevBuf.from_numpy(w), pa.DevBuf.from_numpy(bias)
    dq, do = pa.DevBuf(B * C_ * 4), pa.DevBuf(B * C_ * 4)
    stream = lib.pa_stream_of(eng.h)
    ev = [lib.pa_event_create() for _ in range(3)]
    scratch = pa.DevBuf(256 << 20)      # > L2: evicts the weights between steps and lets the host run ahead of the GPU

    def step():
        assert eng.step_begin(list(range(B)), [1] * B) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        lib.pa_flush_l2(scratch.ptr, 256 << 20, stream)
        lib.pa_event_record(ev[0], stream)
        pa.check(eng.qkv_append(0, dx.ptr, C_, dw.ptr, db.ptr, dq.ptr, C_), "qkv_append")
        lib.pa_event_record(ev[1], stream)
        pa.check(eng.decode(0, dq.ptr, C_, do.ptr, C_), "decode")
        lib.pa_event_record(ev[2], stream)
        eng.sync()
        t = (lib.pa_event_elapsed_ms(ev[0], ev[1]), lib.pa_event_elapsed_ms(ev[1], ev[2]))
        pa.check(eng.step_rollback(), "rollback")
        return t

    for _ in range(5):
        step()
    ts = np.array([step() for _ in range(args.iters)])
    t_qkv, t_dec = np.median(ts[:, 0]) / 1e3, np.median(ts[:, 1]) / 1e3
    flops = 2.0 * B * 3 * C_ * C_
    line = {"tool": "qkv_bench", "shape": args.shape, "B": B, "ctx": ctx, "C": C_, "path": args.path, "split": args.split,
            "qkv_append_us": t_qkv * 1e6, "qkv_tflops": flops / t_qkv / 1e12,
            "qkv_weight_gbs": 3 * C_ * C_ * 4 / t_qkv / 1e9,
            "decode_us": t_dec * 1e6, "qkv_share_of_layer": t_qkv / (t_qkv + t_dec)}
    if not args.no_cpu:
        import oracle_api as oa
        ol = oa.load_oracle("fast")
        out = np.zeros((B, 3 * C_), dtype=np.float32)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            ol.orc_matmul_cached(oa.fptr(out), oa.fptr(x), oa.fptr(w), oa.fptr(bias), B, 1, C_, 3 * C_)
            best = min(best, time.perf_counter() - t0)
        line["cpu_matmul_cached_us"] = best * 1e6
        line["cpu_cores"] = ol.orc_omp_threads()
        line["cpu_kind"] = "port (oracle orc_matmul_cached, -O3 -Ofast -fopenmp)"
    print(json.dumps(line))
    eng.close()


if __name__ == "__main__":
    main()
