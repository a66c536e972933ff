#!/usr/bin/env python
"""bench.py -- paged-attention decode benchmark (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg1] [--impl reference]

A "step" is one decode step of the hot path over the whole batch: block-manager scheduling on the
host, ONE table mirror copy, and for every layer the KV-append kernel + the paged decode-attention
kernel.  N=1 workload = BASELINE.json configs[1] (64 sequences x 1024 ctx, GPT-2 small shape,
shuffled block tables) with all 12 layers of GPT-2 124M, so a step streams 4.8 GB >> L2.

value   = algorithmic bytes of the step (BASELINE.md formula) / device time, inputs resident in HBM
e2e     = the same through the host-buffer entry (pa_decode_step_host: pinned H2D of q|k|v,
          append, decode, D2H of the outputs, per layer, inside the timed region)
roofline= the decode kernel alone, CUDA events around each launch inside the timed region
cpu_baseline / --impl reference = the reference's own CPU code (oracle/_ref, compiled from
          /root/reference) timed on this box's host cores on a bounded sample of the same workload

One process per GPU under torchrun for N>1 (sequences sharded; the path has no collective).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "paged_attn_decode_hbm_throughput"
UNIT = "GB/s"

WORKLOADS = {
    # BASELINE.json configs[1] + the 12 layers of GPT-2 124M
    "cfg2": dict(name="GPT-2 124M shape (12 layers, 12 heads, head_dim 64), 64 sequences x 1024 ctx, block 16, "
                      "shuffled block tables, KV append + paged decode attention per layer",
                 NH=12, hs=64, bs=16, L=12, B=64, ctx="fixed", ctx_len=1024),
    # BASELINE.json configs[2]
    "cfg3": dict(name="GPT-2 124M paged decode, batch 256, mixed ctx U{128..1024}, block 16, KV append + attention",
                 NH=12, hs=64, bs=16, L=12, B=256, ctx="uniform", ctx_lo=128, ctx_hi=1024),
    # BASELINE.json configs[3], one GPU's shard at 8 GPUs (512/8 sequences); 4 of the 48 layers keep set-up short
    "xl": dict(name="GPT-2 XL shape (25 heads, head_dim 64) paged decode, 64 sequences x 1024 ctx per GPU, block 16, 4 of 48 layers",
               NH=25, hs=64, bs=16, L=4, B=64, ctx="fixed", ctx_len=1024),
    # BASELINE.json configs[4], one GPU's share: 32k contexts, head_dim 128
    "long": dict(name="long-context paged decode, 8 sequences x 32768 ctx, 32 heads x head_dim 128, block 16, 1 layer",
                 NH=32, hs=128, bs=16, L=1, B=8, ctx="fixed", ctx_len=32768),
    # BASELINE.json configs[0] (L2-resident, latency-bound)
    "cfg1": dict(name="GPT-2 124M paged decode batch 1, block 16, ctx 256 (L2-resident)",
                 NH=12, hs=64, bs=16, L=12, B=1, ctx="fixed", ctx_len=256),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def context_lengths(w, rank):
    if w["ctx"] == "fixed":
        return [w["ctx_len"]] * w["B"]
    rng = np.random.default_rng(42 + rank)
    return rng.integers(w["ctx_lo"], w["ctx_hi"] + 1, size=w["B"]).tolist()


def decode_bytes(ctx, C_, bs):
    """K and V of valid tokens + q + out + block table + context lengths."""
    B = len(ctx)
    return sum(2 * c * C_ * 4 for c in ctx) + 2 * B * C_ * 4 + sum((c + bs - 1) // bs * 4 for c in ctx) + B * 4


def append_bytes(B, C_):
    return 4 * B * C_ * 4 + B * 4


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.path = f"/tmp/pa_clocks_{os.getpid()}.csv"
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.t_marks = []

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 7:
                try:
                    rows.append((float(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                except ValueError:
                    pass
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not rows:
            return None
        # samples under load = upper half by power draw (the sampler also sees set-up time)
        pw = sorted(r[2] for r in rows)
        thr = pw[len(pw) // 2]
        load = [r for r in rows if r[2] >= thr] or rows
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in load for i in range(4) if r[3][i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(r[0] for r in load), "sm_max_mhz": max(r[1] for r in rows),
                "power_w_max": max(r[2] for r in rows), "reasons": reasons, "samples": len(rows),
                "samples_under_load": len(load)}


# ------------------------------------------------------------------------------------ reference arm
def reference_arm(args, w, quiet=False):
    """Times the reference's own CPU implementation (oracle/_ref, else the oracle port) of ONE
    decode step for ONE sequence and ONE layer of the workload per 'step': add_to_cache(n_tail=1)
    + collect_kv_blocks + attention_paged over the full T=ctx window, exactly what
    paged_infer.c:706-715 executes per generated token (the reference recomputes all T rows)."""
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host thread it can get
    os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    import oracle_api as oa
    NH, hs, bs = w["NH"], w["hs"], w["bs"]
    C_ = NH * hs
    T = w["ctx_len"] if w["ctx"] == "fixed" else (w["ctx_lo"] + w["ctx_hi"]) // 2
    T = min(T, 2048)          # bounded sample: the as-written reference is O(T^2)
    geom = (16, 4352, 256)
    have = bs == 16 and oa.have_ref(*geom, "fast")
    rng = np.random.default_rng(1234)
    inp = rng.standard_normal((1, T, 3 * C_), dtype=np.float32)
    bytes_per = decode_bytes([T], C_, bs) + append_bytes(1, C_)
    if have:
        m = oa.RefManager(C_, *geom, flavor="fast")
        cores = m.lib.ref_omp_threads()
        kind = "reference"
        # cache holds T-1 tokens; each step appends the T-th and attends (then rolls back by
        # rebuilding is too slow: instead keep appending into fresh managers every 16 steps)
        def fresh():
            mm = oa.RefManager(C_, *geom, flavor="fast")
            kv = rng.standard_normal((T - 1, 2, C_), dtype=np.float32)
            for t0 in range(0, T - 1, bs):
                idx = mm.request_block(0)
                k, v = mm.page_arrays(idx)
                n = min(bs, T - 1 - t0)
                k[:n], v[:n] = kv[t0:t0 + n, 0], kv[t0:t0 + n, 1]
                mm.set_filled(idx, n)
            return mm
        m.close()
        out = np.zeros((1, T, C_), dtype=np.float32)

        def one_step():
            mm = fresh()
            mm.lib.ref_silence(1)
            t0 = time.perf_counter()
            mm.lib.ref_add_to_cache(mm.m, oa.fptr(inp), 1, T, C_, 1)
            t1 = time.perf_counter()
            dt = mm.lib.ref_time_attend_prompt(mm.m, 0, oa.fptr(out), oa.fptr(inp), 1, T, C_, NH, 0, 1)
            mm.lib.ref_silence(0)
            mm.close()
            return (t1 - t0) + dt, dt
        sample = (f"reference attention_paged+add_to_cache as written (full T={T} window recomputed per decode step, "
                  f"paged_infer.c:706-715), 1 sequence x 1 layer per step, -O3 -Ofast -fopenmp")
    else:
        orc = oa.OrcManager(C_, bs, (T + bs - 1) // bs + 2, 1, flavor="fast")
        cores = orc.lib.orc_omp_threads()
        kind = "port"
        kv = rng.standard_normal((T, 2, C_), dtype=np.float32)
        for t0 in range(0, T, bs):
            idx = orc.request_block(0)
            k, v = orc.page_arrays(idx)
            n = min(bs, T - t0)
            k[:n], v[:n] = kv[t0:t0 + n, 0], kv[t0:t0 + n, 1]
            orc.set_filled(idx, n)

        def one_step():
            t0 = time.perf_counter()
            orc.attend(0, inp, 1, T, NH, 0)
            t1 = time.perf_counter()
            return t1 - t0, t1 - t0
        sample = f"oracle port of attention_paged (full T={T} window), 1 sequence x 1 layer per step"
    for _ in range(args.warmup):
        one_step()
    times = [one_step()[0] for _ in range(args.steps)]
    t = sum(times) / len(times)
    v = bytes_per / t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tokens_per_s": 1.0 / t}
    if not quiet:
        emit(line)
    return line


def cpu_baseline_sample(w, budget_s=15.0):
    """Bounded CPU sample for the default run (rank 0, N=1): reference as written + last-row port."""
    import oracle_api as oa

    class A:
        pass
    a = A()
    a.gpus, a.steps, a.warmup = 1, 1, 1
    first = reference_arm(a, w, quiet=True)
    per = first["ms_per_step"] / 1e3
    a.steps = int(max(2, min(40, budget_s * 0.6 / max(per, 1e-3))))
    a.warmup = 0
    line = reference_arm(a, w, quiet=True)
    cb = dict(line["cpu_baseline"])
    cb["steps"] = a.steps
    # last-row restatement (what a decode step actually needs) on a slice of the batch
    NH, hs, bs = w["NH"], w["hs"], w["bs"]
    C_ = NH * hs
    T = w["ctx_len"] if w["ctx"] == "fixed" else (w["ctx_lo"] + w["ctx_hi"]) // 2
    nseq = min(w["B"], 16)
    orc = oa.OrcManager(C_, bs, nseq * ((T + bs - 1) // bs) + 2, nseq, flavor="fast")
    rng = np.random.default_rng(7)
    for s in range(nseq):
        for t0 in range(0, T, bs):
            idx = orc.request_block(s)
            k, v = orc.page_arrays(idx)
            n = min(bs, T - t0)
            k[:n] = rng.standard_normal((n, C_), dtype=np.float32)
            v[:n] = rng.standard_normal((n, C_), dtype=np.float32)
            orc.set_filled(idx, n)
    q = rng.standard_normal((nseq, C_), dtype=np.float32)
    out = np.zeros((nseq, C_), dtype=np.float32)
    seq = np.arange(nseq, dtype=np.int32)
    best = orc.lib.orc_time_decode_batch(orc.m, oa.iptr(seq), nseq, NH, oa.fptr(q), C_, oa.fptr(out), C_, 5)
    cb["last_row_port"] = {"value": decode_bytes([T] * nseq, C_, bs) / best / 1e9, "unit": UNIT,
                           "sample": f"oracle last-row restatement (paged_infer.c:182-236 for t=T-1), {nseq} sequences x 1 layer, best of 5",
                           "cores": orc.lib.orc_omp_threads()}
    orc.close()
    return cb


def cpu_model_sample(model, pa_mod, B, L, NH, C_, V, maxT, bs, steps=4):
    """Whole-model decode on the host cores beside the device number: the oracle's restatement of gpt2_forward
    (oracle/paged_oracle.c orc_model_decode_step, all layers, -O3 -Ofast -fopenmp) on the SAME weights,
    `steps` decode steps of B sequences from an empty cache (the projections are >95 % of its work at
    these context lengths).  Bounded: only offered for small batches."""
    import oracle_api as oa
    lib = pa_mod.load()
    n = int(model.n_params)
    params = np.empty(n, dtype=np.float32)
    pa_mod.check(lib.pa_memcpy_d2h(params.ctypes.data, lib.pa_model_params(model.m), n * 4, None), "params d2h")
    pa_mod.check(lib.pa_device_sync(), "sync")
    ol = oa.load_oracle("fast")
    pages = (steps + 2 + bs - 1) // bs + 1
    mgrs = [oa.OrcManager(C_, bs, B * pages + 2, B, flavor="fast") for _ in range(L)]
    arr = (C.c_void_p * L)(*[m.m for m in mgrs])
    seq = np.arange(B, dtype=np.int32)
    tok = np.arange(B, dtype=np.int32) + 11
    logits = np.zeros((B, V), dtype=np.float32)
    times = []
    try:
        for step in range(steps + 1):
            pos = np.full(B, step, dtype=np.int32)
            t0 = time.perf_counter()
            rc = ol.orc_model_decode_step(arr, L, NH, C_, V, maxT, oa.fptr(params), oa.iptr(seq), oa.iptr(tok), oa.iptr(pos), B,
                                          oa.fptr(logits))
            if rc != 0:
                raise RuntimeError(f"orc_model_decode_step rc={rc}")
            times.append(time.perf_counter() - t0)
            tok = logits.argmax(axis=1).astype(np.int32)
    finally:
        for m in mgrs:
            m.close()
    per = min(times[1:])
    return {"tokens_per_s": B / per, "ms_per_step": per * 1e3, "cores": ol.orc_omp_threads(), "kind": "port",
            "sample": f"oracle gpt2_forward restatement (all {L} layers, LM head, V={V}), best of {steps} decode steps of {B} "
                      f"sequence(s) from an empty cache, -O3 -Ofast -fopenmp"}


# ------------------------------------------------------------------------------------ our arm
_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries write to stdout too (NCCL prints its version there on the first collective): everything
    # but the JSON line goes to stderr, so stdout carries exactly one line
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model", action="store_true", help="skip the whole-model decode step section (SURVEY 8f.2)")
    ap.add_argument("--no-clock-hold", action="store_true", help="skip the untimed >=1.5 s continuation (ncu runs)")
    ap.add_argument("--hpg", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--static-pct", type=int, default=0)
    ap.add_argument("--dyn-units", type=int, default=0)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-zerocopy", action="store_true", help="e2e: stage pinned host buffers with copies")
    ap.add_argument("--timeline", default="", help="write a per-CTA timeline of one decode launch to this file")
    ap.add_argument("--no-fuse", action="store_true", help="separate KV-append kernel instead of the fused decode+append")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = dict(WORKLOADS[args.workload])
    if args.layers > 0:
        w["L"] = args.layers

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import __graft_entry__ as ge
    if args.impl == "reference":
        if rank == 0:
            ge.build(quiet=True)
            reference_arm(args, w)
        return 0

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if rank == 0:
        pa = ge.build(quiet=True)
    if dist is not None:
        dist.barrier()
    pa = ge.load_binding()
    lib = pa.load()
    if lib.pa_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libpaged_attn has no CPU fallback")

    NH, hs, bs, L, B = w["NH"], w["hs"], w["bs"], w["L"], w["B"]
    C_ = NH * hs
    ctx = context_lengths(w, rank)
    pages = [(c + bs - 1) // bs for c in ctx]
    n_blocks = sum(pages) + 64
    eng = pa.PagedAttn(bs, n_blocks, B, NH, hs, n_layers=L, device=local_rank, max_batch_tokens=B)
    eng.tune(pa.PA_TUNE_HEADS_PER_TILE, args.hpg)
    eng.tune(pa.PA_TUNE_STAGES, args.stages)
    eng.tune(pa.PA_TUNE_GRID, args.grid)
    eng.tune(pa.PA_TUNE_STATIC_PCT, args.static_pct)
    eng.tune(pa.PA_TUNE_DYN_UNITS, args.dyn_units)
    eng.tune(pa.PA_TUNE_NO_PDL, 1 if args.no_pdl else 0)
    eng.tune(pa.PA_TUNE_NO_ZEROCOPY, 1 if args.no_zerocopy else 0)

    # ---- synthetic state: random-init K/V pools (seeded), shuffled block tables ------------
    rng = np.random.default_rng(1234 + rank)
    chunk = rng.standard_normal((min(n_blocks, 1024) * bs, C_), dtype=np.float32)
    for layer in range(L):
        for which, base in ((0, eng.pool_k(layer)), (1, eng.pool_v(layer))):
            off = 0
            total = n_blocks * bs * C_ * 4
            shift = (layer * 2 + which) * 4096 % chunk.nbytes     # decorrelate layers a little
            while off < total:
                n = min(chunk.nbytes - shift, total - off)
                pa.check(lib.pa_memcpy_h2d(base + off, chunk.ctypes.data + shift, n, None), "pool fill")
                off += n
                shift = 0
    perm = rng.permutation(sum(pages))        # fragmented: a seeded permutation dealt to the sequences
    cur = 0
    for s in range(B):
        # one token short: every timed step appends the last token, decodes at the named ctx,
        # and is then rolled back on the host so that all steps see the same context lengths
        n_tok = ctx[s] - 1
        n_pg = (n_tok + bs - 1) // bs
        blocks = perm[cur:cur + pages[s]]
        cur += pages[s]
        if n_tok > 0:
            assert eng.seq_adopt(s, blocks[:n_pg], n_tok) == 0, pa.last_error()
    seq_ids = np.arange(B, dtype=np.int32)
    ones = np.ones(B, dtype=np.int32)

    stream = lib.pa_stream_create()
    qkv_host = lib.pa_host_alloc(B * 3 * C_ * 4)
    out_host = lib.pa_host_alloc(B * C_ * 4)
    qkv_np = np.ctypeslib.as_array(C.cast(qkv_host, C.POINTER(C.c_float)), (B, 3 * C_))
    qkv_np[:] = rng.standard_normal((B, 3 * C_), dtype=np.float32)
    d_qkv = pa.DevBuf.from_numpy(qkv_np)
    d_out = pa.DevBuf(B * C_ * 4)

    def rollback():
        pa.check(eng.step_rollback(), "rollback")

    # kernel time: events around the L back-to-back decode launches of a step (no event between
    # launches: that would serialise them and defeat programmatic dependent launch)
    evs = [lib.pa_event_create() for _ in range(2)]
    dec_ms = []

    def step(timed_kernels=False):
        pa.check(eng.step_begin(seq_ids, ones), "step_begin")
        pa.check(eng.upload(stream), "upload")
        if timed_kernels:
            lib.pa_event_record(evs[0], stream)
        for layer in range(L):
            if args.no_fuse:
                pa.check(eng.append(layer, d_qkv.ptr + C_ * 4, d_qkv.ptr + 2 * C_ * 4, 3 * C_, stream), "append")
                pa.check(eng.decode(layer, d_qkv.ptr, 3 * C_, d_out.ptr, C_, stream), "decode")
            else:
                pa.check(eng.decode_append(layer, d_qkv.ptr, d_qkv.ptr + C_ * 4, d_qkv.ptr + 2 * C_ * 4, 3 * C_,
                                           d_out.ptr, C_, stream), "decode_append")
        if timed_kernels:
            lib.pa_event_record(evs[1], stream)
        rollback()

    def collect_kernel_times():
        dec_ms.append(lib.pa_event_elapsed_ms(evs[0], evs[1]) / L)

    def barrier():
        pa.check(lib.pa_device_sync(), "sync")
        if dist is not None:
            dist.barrier()
            pa.check(lib.pa_device_sync(), "sync")

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = eng.launches()
    e0, e1 = lib.pa_event_create(), lib.pa_event_create()
    # the kernel events are recorded in every timed step and read back every 8th (reading needs a
    # sync, which would serialise host and device if done each step)
    lib.pa_event_record(e0, stream)
    for i in range(args.steps):
        step(timed_kernels=True)
        if i % 8 == 7 or i == args.steps - 1:
            collect_kernel_times()
    lib.pa_event_record(e1, stream)
    barrier()
    ms_total = allmax(lib.pa_event_elapsed_ms(e0, e1))
    launches = eng.launches() - launches0
    # keep the same load on for >= 1.5 s in total so the clock sampler sees it (untimed)
    t_end = time.time() + (0.0 if args.no_clock_hold else max(0.0, 1.5 - ms_total / 1e3))
    while time.time() < t_end:
        step()
        lib.pa_stream_sync(stream)
    barrier()
    clocks = sampler.stop() if sampler else None

    if args.timeline and rank == 0:
        eng.tune(pa.PA_TUNE_DEBUG_TIMELINE, 1)
        step()
        tl = eng.debug_timeline().astype(np.int64)
        eng.tune(pa.PA_TUNE_DEBUG_TIMELINE, 0)
        t0 = tl[:, 0].min()
        with open(args.timeline, "w") as f:
            f.write("cta,entry_ns,first_issue_ns,first_tile_ns,last_tile_ns,prod_wait_cyc,cons_wait_cyc,seg_cyc,tiles\n")
            for i, r in enumerate(tl):
                f.write(f"{i},{r[0]-t0},{r[1]-t0},{r[2]-t0},{r[3]-t0},{r[4]},{r[5]},{r[6]},{r[7]}\n")

    step_bytes = L * (decode_bytes(ctx, C_, bs) + append_bytes(B, C_))
    ms_per_step = ms_total / args.steps
    value = world * step_bytes / (ms_per_step * 1e-3) / 1e9
    tokens_per_s = world * B / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through pa_decode_step_host ------------------------------------
    e2e = None
    if not args.no_e2e:
        out_np = np.ctypeslib.as_array(C.cast(out_host, C.POINTER(C.c_float)), (B, C_))

        hstream = lib.pa_stream_of(eng.h)

        def e2e_step():
            pa.check(eng.step_begin(seq_ids, ones), "step_begin")
            for layer in range(L):      # the layers of a step are queued behind each other, one sync per step
                pa.check(eng.decode_step_host_async(layer, qkv_host, out_host), "decode_step_host_async")
            pa.check(lib.pa_decode_step_host_sync(eng.h), "sync")
            rollback()
            return float(out_np[0, 0])      # the step's result is read on the host
        k_e2e = max(5, min(args.steps, 50))
        for _ in range(3):
            e2e_step()
        barrier()
        hs_ = lib.pa_stream_of(eng.h)
        lib.pa_event_record(e0, hs_)
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
        lib.pa_event_record(e1, hs_)
        barrier()
        wall = time.perf_counter() - t0
        ms_e2e = allmax(max(lib.pa_event_elapsed_ms(e0, e1), 0.0)) / k_e2e
        table_bytes = (4 * B + 2 + B + 4 + B * ((max(pages) + 3) & ~3)) * 4
        e2e = {"value": world * step_bytes / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": L * B * 3 * C_ * 4 + table_bytes, "d2h_bytes_per_step": L * B * C_ * 4,
               "ms_per_step": ms_e2e, "steps": k_e2e, "tokens_per_s": world * B / (ms_e2e * 1e-3),
               "entry": "pa_decode_step_host_async per layer + one stream sync per step: host q|k|v rows in pinned memory are " +
                        ("copied H2D, then fused append+decode, then D2H copy" if args.no_zerocopy else
                         "pulled over PCIe by the kernel's bulk copies, outputs stored to pinned host memory by the kernel"),
               "wall_ms_per_step": wall * 1e3 / k_e2e}

    # ---- whole-model decode step: embedding, L x {ln, QKV+append, paged attention,
    # attproj, ln, MLP}, LM head, sampler through pa_model_decode_step (host tokens in, host tokens
    # out); with N > 1 the sampled tokens of all ranks are gathered with one NCCL all_gather per step
    model_info = None
    if not args.no_model and hs == 64 and max(ctx) <= 4096:
        try:
            V, maxT = 50257, max(1024, max(ctx) + 8)
            model = pa.Model(eng, maxT, V, params=None, seed=1337 + rank, max_batch=B)
            toks = rng.integers(0, V, size=B).astype(np.int32)
            coins = rng.random(B).astype(np.float32)
            gathered = None
            if dist is not None:
                import torch
                t_local = torch.zeros(B, dtype=torch.int32, device=f"cuda:{local_rank}")
                gathered = torch.zeros(world * B, dtype=torch.int32, device=f"cuda:{local_rank}")

            def model_step():
                nxt = model.decode_step(seq_ids, toks, coins)
                rollback()
                if dist is not None:
                    t_local.copy_(torch.from_numpy(nxt))
                    dist.all_gather_into_tensor(gathered, t_local)
                return nxt
            for _ in range(3):
                model_step()
            barrier()
            k_model = max(5, min(args.steps, 30))
            hs_ = lib.pa_stream_of(eng.h)
            l0 = eng.launches()
            lib.pa_event_record(e0, hs_)
            t0 = time.perf_counter()
            for _ in range(k_model):
                model_step()
            lib.pa_event_record(e1, hs_)
            barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3 / k_model
            ms_model = allmax(max(lib.pa_event_elapsed_ms(e0, e1) / k_model, wall_ms))
            per_step = (eng.launches() - l0) / k_model
            persistent = per_step <= 1.0
            model_info = {"tokens_per_s": world * B / (ms_model * 1e-3), "ms_per_step": ms_model, "steps": k_model,
                          "vocab": V, "layers": L, "gpu_launches_per_step": per_step,
                          "path": ("ONE persistent cooperative kernel for the whole step (pa_model_mega.cu: weight-streaming "
                                   "fp32 GEMV phases, chunked paged attention, grid barriers)" if persistent else
                                   "chain of per-op kernels (layernorm, tcgen05 3xTF32 projections, paged decode attention, sampler)"),
                          "sampler": "softmax + multinomial (sample_mult) fused, host coins",
                          "projections": ("fp32 FMA weight-streaming GEMV with bias/GELU/residual epilogues" if persistent else
                                          "tcgen05 3xTF32 (fp32-accurate) GEMMs with bias/GELU/residual epilogues"),
                          "weights": "random init on the device (no checkpoint offline)",
                          "token_gather": ("NCCL all_gather of int32 next tokens, every step" if dist is not None else None),
                          "entry": "pa_model_decode_step (host token ids in, host token ids out, sync per step)"}
            if world == 1 and B <= 8 and not args.no_cpu_baseline:
                try:
                    model_info["cpu_baseline"] = cpu_model_sample(model, pa, B, L, NH, C_, V, maxT, bs)
                except Exception as ex:      # a report, never a reason to lose the line
                    model_info["cpu_baseline"] = {"tokens_per_s": None, "sample": f"failed: {ex!r}"}
            model.close()
        except Exception as ex:
            model_info = {"error": repr(ex)}

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak, peak_src = peaks()
    kbytes = decode_bytes(ctx, C_, bs) + append_bytes(B, C_)      # per layer: append + decode
    kms = sum(dec_ms) / len(dec_ms)
    achieved = kbytes / (kms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    last_ctas = lib.pa_tune_get(eng.h, 12)       # 0: the launch did not go through the stream kernel
    kname = (f"pa_decode_stream_kernel<{hs},{bs}>" if last_ctas else
             "pa_decode_small_kernel (one CTA per sequence and head; chosen below ~98 k token-heads, latency-bound)")
    roofline = {"bound": "hbm", "kernel": kname + ("" if args.no_fuse else " (KV append fused)"), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": kbytes, "avg_launch_ms": kms, "launches_timed": len(dec_ms) * L,
                "timing": f"CUDA events around the {L} back-to-back per-layer launches of a step, / {L}",
                "frac_of_nominal_8TBps": achieved / 8000.0,
                "kernel_share_of_step": kms * L / ms_per_step,
                "plan": {"heads_per_tile": lib.pa_tune_get(eng.h, 10), "ring_stages": lib.pa_tune_get(eng.h, 11),
                         "ctas": lib.pa_tune_get(eng.h, 12)}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = cpu_baseline_sample(w)
        except Exception as ex:     # the baseline is a report, never a reason to lose the line
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"failed: {ex!r}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "workload_id": args.workload, "layers": L, "batch_per_gpu": B,
                       "ctx_mean": sum(ctx) / len(ctx), "block_size": bs,
                       "l2": f"inputs larger than L2: {step_bytes / 1e9:.2f} GB streamed per step vs 126 MB L2",
                       "parallelism": f"sequences sharded over {world} GPU(s), no data-path collective"},
            "tokens_per_s": tokens_per_s, "frac_of_measured_peak": value / world / peak,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "model": model_info}
    if model_info and "ms_per_step" in model_info and model_info["gpu_launches_per_step"] > 1:
        model_info["attention_share_of_step"] = kms * L / model_info["ms_per_step"]
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
