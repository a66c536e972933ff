"""Generate the committed golden vectors from the reference's OWN compiled code
(oracle/_ref/libref_*_strict.so, built by oracle/build_ref.sh from /root/reference).
Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Outputs (small, committed):
  golden_window_bs32.npz   paged_infer.c main's sliding-window pattern through the real
                           add_to_cache + collect_kv_blocks + attention_paged (T=32, C=48, NH=4)
  golden_attn_bs16.npz     one full-window attention_paged call, bs=16, C=128, NH=2, T=64, offset=3
  golden_attn_bs2.npz      test_paged_attn.c's shape (T=20, C=10, NH=2, block 2), U[0,100) inputs
  golden_trace.json        a randomized allocator trace with every return value and snapshots
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_api as oa                                   # noqa: E402
from trace_driver import make_trace, run_trace            # noqa: E402


def fill_pages(mgr, prompt, kv, bs):
    for t0 in range(0, kv.shape[0], bs):
        idx = mgr.request_block(prompt)
        k, v = mgr.page_arrays(idx)
        n = min(bs, kv.shape[0] - t0)
        k[:n], v[:n] = kv[t0:t0 + n, 0], kv[t0:t0 + n, 1]
        mgr.set_filled(idx, n)


def window_case():
    bs, mb, mp, T, C, NH = 32, 100, 100, 32, 48, 4
    ref = oa.RefManager(C, bs, mb, mp, "strict")
    stream = oa.normal((T + 18, 3 * C), seed=2024)
    outs, tables, epochs = [], [], []
    for step in range(19):
        window = np.ascontiguousarray(stream[step:step + T][None])
        ref.add_to_cache(window, 1, T, T if step == 0 else 1)
        _, out = ref.attend(0, window, 1, T, NH, step)
        outs.append(out[0]); tables.append(ref.table(0) + [-1] * (2 - len(ref.table(0)))); epochs.append(ref.epoch())
    filled = [ref.block_info(i)[0] for i in ref.table(0)]
    ref.close()
    np.savez_compressed(os.path.join(HERE, "golden_window_bs32.npz"), stream=stream, outs=np.stack(outs),
                        tables=np.array(tables, dtype=np.int32), epochs=np.array(epochs, dtype=np.int32),
                        filled=np.array(filled, dtype=np.int32), geom=np.array([bs, mb, mp, T, C, NH], dtype=np.int32))


def attn_case(name, bs, mb, mp, T, C, NH, offset, dist):
    ref = oa.RefManager(C, bs, mb, mp, "strict")
    ntok = T + offset
    if dist == "normal":
        kv = oa.normal((ntok, 2, C), seed=1234 + T)
        inp = oa.normal((1, T, 3 * C), seed=99 + C)
    else:
        kv = oa.uniform((ntok, 2, C), 0.0, 100.0, seed=1234 + T)
        inp = oa.uniform((1, T, 3 * C), 0.0, 100.0, seed=99 + C)
    fill_pages(ref, 0, kv, bs)
    _, out, pre, att = ref.attend(0, inp, 1, T, NH, offset, want_scratch=True)
    ref.close()
    np.savez_compressed(os.path.join(HERE, name), kv=kv, inp=inp, out=out, att=att,
                        geom=np.array([bs, mb, mp, T, C, NH, offset], dtype=np.int32))


def trace_case():
    bs, mb, mp = 16, 12, 8
    ops = make_trace(1234, 400, mp, bs)
    ref = oa.RefManager(4, bs, mb, mp, "strict")
    log = run_trace(ref, ops, mp, mb, snap_every=10)
    ref.close()
    with open(os.path.join(HERE, "golden_trace.json"), "w") as f:
        json.dump({"geom": [bs, mb, mp], "log": log}, f, separators=(",", ":"))


if __name__ == "__main__":
    assert oa.have_ref(32, 100, 100), "build oracle/_ref first: oracle/build_ref.sh"
    window_case()
    attn_case("golden_attn_bs16.npz", 16, 100, 100, 64, 128, 2, 3, "normal")
    attn_case("golden_attn_bs2.npz", 2, 64, 8, 20, 10, 2, 0, "uniform")
    trace_case()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
