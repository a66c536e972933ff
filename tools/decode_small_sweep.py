#!/usr/bin/env python
"""Stream decode kernel (PA_TUNE_DECODE_PATH 1) against the small-batch kernel (3) over a few sequences:
us per pa_decode launch (CUDA events around 50 back-to-back launches on the handle's stream), to place
the automatic switch (PA_DECODE_SMALL_MAX).

  python tools/decode_small_sweep.py [--NH 12 --hs 64]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--NH", type=int, default=12)
    ap.add_argument("--hs", type=int, default=64)
    ap.add_argument("--bs", type=int, default=16)
    args = ap.parse_args()
    pa = ge.build(quiet=True)
    lib = pa.load()
    if lib.pa_device_count() < 1:
        raise SystemExit("decode_small_sweep: no CUDA device; libpaged_attn has no CPU fallback")
    NH, hs, bs = args.NH, args.hs, args.bs
    Cc = NH * hs
    rng = np.random.default_rng(1)
    for ctx in (256, 1024):
        for B in (1, 2, 4, 8, 16, 32, 64):
            pages = (ctx + bs - 1) // bs
            eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=1, device=0, max_batch_tokens=B)
            perm = rng.permutation(B * pages + 8)
            for s in range(B):
                assert eng.seq_adopt(s, perm[s * pages:(s + 1) * pages], ctx) == 0, pa.last_error()
            q = pa.DevBuf.from_numpy(rng.standard_normal((B, Cc)).astype(np.float32))
            out = pa.DevBuf(B * Cc * 4)
            assert eng.step_begin_readonly(list(range(B))) == 0, pa.last_error()
            pa.check(eng.upload(), "upload")
            st = lib.pa_stream_of(eng.h)
            e0, e1 = lib.pa_event_create(), lib.pa_event_create()
            row = {"ctx": ctx, "B": B, "token_heads": B * ctx * NH}
            for path in (1, 3):
                eng.tune(pa.PA_TUNE_DECODE_PATH, path)
                for _ in range(5):
                    pa.check(eng.decode(0, q.ptr, Cc, out.ptr, Cc), "decode")
                lib.pa_event_record(e0, st)
                for _ in range(50):
                    pa.check(eng.decode(0, q.ptr, Cc, out.ptr, Cc), "decode")
                lib.pa_event_record(e1, st)
                row["stream_us" if path == 1 else "small_us"] = round(lib.pa_event_elapsed_ms(e0, e1) * 1e3 / 50, 2)
            print(json.dumps(row))
            eng.close()


if __name__ == "__main__":
    main()
