#!/usr/bin/env python
"""Whole-model decode step on one B200: GPT-2 shaped model with random-init weights,
`pa_model_decode_step` (host token ids in, host token ids out).  Used for the ncu launch list of a step.

  python tools/model_bench.py [--shape 124m|xl] [--B n] [--ctx n] [--layers n] [--steps n]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

SHAPES = {"124m": (12, 64, 12), "xl": (25, 64, 48)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="124m", choices=sorted(SHAPES))
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=576)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--bs", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--gemm-path", type=int, default=0)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--prefill", type=int, default=0, help="instead of decode steps: prefill prompts of this many tokens for the B sequences in ONE step")
    ap.add_argument("--replicas", type=int, default=1, help="EXPERIMENT: split the B sequences over this many engine + model pairs on the same GPU, "
                    "each on its own stream, stepped concurrently (pa_model_forward_async on each, then pa_model_wait on each)")
    ap.add_argument("--attn-grid", type=int, default=0, help="with --replicas: CTAs of the decode attention launch (PA_TUNE_GRID)")
    ap.add_argument("--model-path", type=int, default=0, help="0 auto, 1 chain of per-op kernels, 2 persistent step kernel")
    args = ap.parse_args()
    pa = ge.build(quiet=True)
    lib = pa.load()
    if lib.pa_device_count() < 1:
        raise SystemExit("model_bench: no CUDA device; libpaged_attn has no CPU fallback")
    NH, hs, L = SHAPES[args.shape]
    if args.layers:
        L = args.layers
    B, bs, ctx = args.B, args.bs, args.ctx
    pages = (ctx + bs - 1) // bs + 1
    eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=L, device=0, max_batch_tokens=B)
    eng.tune(pa.PA_TUNE_GEMM_PATH, args.gemm_path)
    eng.tune(pa.PA_TUNE_NO_PDL, 1 if args.no_pdl else 0)
    eng.tune(pa.PA_TUNE_MODEL_PATH, args.model_path)
    rng = np.random.default_rng(3)
    perm = rng.permutation(B * pages + 8)
    if args.prefill:
        # prompt prefill through pa_model_forward: B prompts of `prefill` tokens in one step (causal prefill kernel,
        # projections at B * prefill rows, KV append fused into the QKV projection), sequences freed after every step
        P = args.prefill
        eng.close()
        pages = (P + bs - 1) // bs + 1
        eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=L, device=0, max_batch_tokens=B * P)
        eng.tune(pa.PA_TUNE_GEMM_PATH, args.gemm_path)
        model = pa.Model(eng, max(1024, P + 8), 50257, params=None, seed=1, max_batch=B * P)
        seq = np.arange(B, dtype=np.int32)
        toks = rng.integers(0, 50257, size=B * P).astype(np.int32)

        def pstep():
            nxt = model.forward(seq, [P] * B, toks, None)
            for s in range(B):
                eng.seq_free(s)
            return nxt
        for _ in range(2):
            pstep()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pstep()
        dt = (time.perf_counter() - t0) / args.steps
        print(json.dumps({"tool": "model_bench", "mode": "prefill", "shape": args.shape, "B": B, "prompt": P, "layers": L,
                          "ms_per_step": dt * 1e3, "prompt_tokens_per_s": B * P / dt, "gemm_path": args.gemm_path}))
        model.close(); eng.close()
        return
    if args.replicas > 1:
        import ctypes as C
        eng.close()
        R = args.replicas
        Bp = B // R
        V = 50257
        engs, models = [], []
        for r in range(R):
            e = pa.PagedAttn(bs, Bp * pages + 8, Bp, NH, hs, n_layers=L, device=0, max_batch_tokens=Bp)
            e.tune(pa.PA_TUNE_GEMM_PATH, args.gemm_path)
            e.tune(pa.PA_TUNE_NO_PDL, 1 if args.no_pdl else 0)
            e.tune(pa.PA_TUNE_MODEL_PATH, args.model_path)
            e.tune(pa.PA_TUNE_GRID, args.attn_grid)
            pr = rng.permutation(Bp * pages + 8)
            for s in range(Bp):
                assert e.seq_adopt(s, pr[s * pages: s * pages + (ctx - 1 + bs - 1) // bs], ctx - 1) == 0, pa.last_error()
            engs.append(e)
            models.append(pa.Model(e, max(1024, ctx + 8), V, params=None, seed=1, max_batch=Bp))
        seq = np.arange(Bp, dtype=np.int32)
        ones = np.ones(Bp, dtype=np.int32)
        tok = rng.integers(0, V, size=Bp).astype(np.int32)
        coins = rng.random(Bp).astype(np.float32)
        nxt = np.zeros(Bp, dtype=np.int32)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))

        def rstep():
            for m in models:
                pa.check(lib.pa_model_forward_async(m.m, ip(seq), ip(ones), ip(tok), coins.ctypes.data, Bp), "forward_async")
            for m in models:
                pa.check(lib.pa_model_wait(m.m, ip(nxt)), "wait")
            for e in engs:
                pa.check(e.step_rollback(), "rollback")
        for _ in range(3):
            rstep()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rstep()
        dt = (time.perf_counter() - t0) / args.steps
        print(json.dumps({"tool": "model_bench", "mode": "replicas", "replicas": R, "shape": args.shape, "B": B, "ctx": ctx, "layers": L,
                          "ms_per_step": dt * 1e3, "tokens_per_s": Bp * R / dt}))
        for m in models:
            m.close()
        for e in engs:
            e.close()
        return
    for s in range(B):
        assert eng.seq_adopt(s, perm[s * pages: s * pages + (ctx - 1 + bs - 1) // bs], ctx - 1) == 0, pa.last_error()
    V = 50257
    model = pa.Model(eng, max(1024, ctx + 8), V, params=None, seed=1, max_batch=B)
    seq = np.arange(B, dtype=np.int32)
    tok = rng.integers(0, V, size=B).astype(np.int32)
    coins = rng.random(B).astype(np.float32)

    def step():
        nxt = model.decode_step(seq, tok, coins)
        pa.check(eng.step_rollback(), "rollback")
        return nxt
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"tool": "model_bench", "shape": args.shape, "B": B, "ctx": ctx, "layers": L, "ms_per_step": dt * 1e3,
                      "tokens_per_s": B / dt, "gemm_path": args.gemm_path, "pdl": not args.no_pdl,
                      "model_path": args.model_path}))
    model.close(); eng.close()


if __name__ == "__main__":
    main()
