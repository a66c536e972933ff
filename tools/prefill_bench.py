#!/usr/bin/env python
"""Prefill (causal multi-row paged attention) throughput on one B200, through the C ABI.

One 'pass' = pa_append (K/V of the prompt rows) + pa_prefill for one layer over a batch of
prompts.  Reported: TFLOP/s counted as 4*hs flop per (query row, visible key, head) -- QK^T and
PV, causal (only the keys a row actually sees) -- against the fp32 FFMA peak of the chip
(148 SMs x 128 lanes x 2 flop x SM clock) for the SIMT path and the measured dense tensor peak
for the tcgen05 path.  Results go to stdout as one JSON line per configuration.

  python tools/prefill_bench.py [--shape 124m|xl|long] [--B n] [--T n] [--before n] [--path 0|1|2|3]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

SHAPES = {"124m": (12, 64), "xl": (25, 64), "long": (32, 128)}
PATH_NAMES = {0: "auto", 1: "tiled fp32 SIMT", 2: "generic rows", 3: "tcgen05 tf32", 4: "tcgen05 3xtf32 (fp32-accurate)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="124m", choices=sorted(SHAPES))
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=1024, help="new (query) rows per sequence")
    ap.add_argument("--before", type=int, default=0, help="tokens already cached per sequence (chunked prefill)")
    ap.add_argument("--bs", type=int, default=16)
    ap.add_argument("--path", type=int, default=0)
    ap.add_argument("--nwg", type=int, default=0, help="tcgen05 path: softmax warpgroups per CTA")
    ap.add_argument("--bn", type=int, default=0, help="tcgen05 path, head_dim 64: keys per tile (64/128)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clocks", type=float, default=0.0, metavar="SECONDS",
                    help="afterwards repeat the attention launch alone for this long and report SM clock / power under load (NVML)")
    ap.add_argument("--check", action="store_true", help="compare against the rows kernel (path 2) on the same inputs")
    args = ap.parse_args()

    pa = ge.build(quiet=True)
    lib = pa.load()
    if lib.pa_device_count() < 1:
        raise SystemExit("prefill_bench: no CUDA device; libpaged_attn has no CPU fallback")
    NH, hs = SHAPES[args.shape]
    C_ = NH * hs
    B, T, bs = args.B, args.T, args.bs
    pages = (args.before + T + bs - 1) // bs + 1
    eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=1, device=0, max_batch_tokens=B * max(T, args.before) + 8)
    rng = np.random.default_rng(5)
    # shuffled block tables holding `before` cached tokens
    perm = rng.permutation(B * pages + 8)
    if args.before > 0:
        nb = (args.before + bs - 1) // bs
        for s in range(B):
            assert eng.seq_adopt(s, perm[s * pages: s * pages + nb], args.before) == 0, pa.last_error()
        pool = rng.standard_normal(((B * pages + 8) * bs, C_), dtype=np.float32)
        pa.check(lib.pa_memcpy_h2d(eng.pool_k(0), pool.ctypes.data, pool.nbytes, None), "h2d")
        pa.check(lib.pa_memcpy_h2d(eng.pool_v(0), pool.ctypes.data, pool.nbytes, None), "h2d")
    ntok = B * T
    qkv = rng.standard_normal((ntok, 3 * C_), dtype=np.float32)
    d = pa.DevBuf.from_numpy(qkv)
    o = pa.DevBuf(ntok * C_ * 4)
    stream = lib.pa_stream_of(eng.h)
    e0, e1 = lib.pa_event_create(), lib.pa_event_create()
    eng.tune(pa.PA_TUNE_PREFILL_PATH, args.path)
    eng.tune(pa.PA_TUNE_TC_WARPGROUPS, args.nwg)
    eng.tune(pa.PA_TUNE_TC_KEY_TILE, args.bn)

    def one_pass(timed):
        assert eng.step_begin(list(range(B)), [T] * B) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        if timed:
            lib.pa_event_record(e0, stream)
        pa.check(eng.append(0, d.ptr + C_ * 4, d.ptr + 2 * C_ * 4, 3 * C_), "append")
        pa.check(eng.prefill(0, d.ptr, 3 * C_, o.ptr, C_), "prefill")
        if timed:
            lib.pa_event_record(e1, stream)
        eng.sync()
        ms = lib.pa_event_elapsed_ms(e0, e1) if timed else 0.0
        pa.check(eng.step_rollback(), "rollback")       # same cache state for every pass
        return ms

    for _ in range(args.warmup):
        one_pass(False)
    ms = [one_pass(True) for _ in range(args.iters)]
    t = float(np.median(ms)) / 1e3
    keys_seen = B * sum(args.before + j + 1 for j in range(T))
    flops = 4.0 * hs * NH * keys_seen
    line = {"tool": "prefill_bench", "shape": args.shape, "NH": NH, "hs": hs, "B": B, "T": T, "before": args.before,
            "bs": bs, "path": PATH_NAMES[args.path], "nwg": args.nwg, "bn": args.bn, "ms": t * 1e3, "tflops": flops / t / 1e12,
            "tokens_per_s": ntok / t, "ms_all": [round(x, 4) for x in ms]}
    if args.clocks > 0:
        # the step's attention launch back to back (no append: the cache is already in its post-step state for the
        # last pass) while a thread samples NVML: what clock does the kernel actually run at?
        import threading, time
        import pynvml
        pynvml.nvmlInit()
        dev = pynvml.nvmlDeviceGetHandleByIndex(0)
        samples, stop = [], threading.Event()

        def sampler():
            while not stop.is_set():
                samples.append((pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM),
                                pynvml.nvmlDeviceGetPowerUsage(dev) / 1000.0))
                time.sleep(0.02)
        assert eng.step_begin(list(range(B)), [T] * B) == 0, pa.last_error()
        pa.check(eng.upload(), "upload")
        pa.check(eng.append(0, d.ptr + C_ * 4, d.ptr + 2 * C_ * 4, 3 * C_), "append")
        th = threading.Thread(target=sampler)
        th.start()
        t_end = time.time() + args.clocks
        n_launch = 0
        lib.pa_event_record(e0, stream)
        while time.time() < t_end:
            for _ in range(20):
                pa.check(eng.prefill(0, d.ptr, 3 * C_, o.ptr, C_), "prefill")
            n_launch += 20
            eng.sync()
        lib.pa_event_record(e1, stream)
        eng.sync()
        stop.set(); th.join()
        pa.check(eng.step_rollback(), "rollback")
        half = samples[len(samples) // 2:]               # (the first half: ramp)
        ms_k = lib.pa_event_elapsed_ms(e0, e1) / n_launch
        line["sustained"] = {"attention_ms": ms_k, "tflops": flops / (ms_k * 1e-3) / 1e12,
                             "sm_mhz": float(np.median([x[0] for x in half])), "power_w": float(np.median([x[1] for x in half])),
                             "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(dev, pynvml.NVML_CLOCK_SM), "seconds": args.clocks}
    if args.check:
        got = o.download((ntok, C_))
        eng.tune(pa.PA_TUNE_PREFILL_PATH, 2)
        one_pass(False)
        want = o.download((ntok, C_))
        line["max_rel_err_vs_rows_kernel"] = float(np.abs(got - want).max() / np.abs(want).max())
    print(json.dumps(line))
    eng.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
