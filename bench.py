#!/usr/bin/env python
"""bench.py -- paged-attention decode benchmark (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3|cfg4|cfg5] [--configs all|none|a,b]
                  [--impl reference]

A "step" is one decode step of the hot path over the whole batch: block-manager scheduling on the
host, ONE table mirror copy, and for every layer the fused KV-append + paged decode-attention kernel.
The headline workload (top level of the JSON line) is BASELINE.json configs[1] -- 64 sequences x 1024 ctx,
GPT-2 small shape, shuffled block tables -- with all 12 layers of GPT-2 124M, so a step streams
4.8 GB >> L2.  The line's "configs" section carries SHORT runs of the other four BASELINE configs at
their stated shapes (cfg1 batch 1; cfg3 batch 256 mixed contexts; cfg4 GPT-2 XL, all 48 layers, batch 512
sharded over the GPUs; cfg5 32k contexts at head_dim 128, prompt prefill + decode), each with its own
roofline / e2e / verified entries.

value    = algorithmic bytes of the step (BASELINE.md formula) / device time, inputs resident in HBM
e2e      = the same through the host-buffer entry (pa_decode_step_host_async: pinned host q|k|v rows in,
           host outputs back, inside the timed region)
roofline = the decode kernel alone, CUDA events around the back-to-back per-layer launches of a step
verified = after the timed loops, output rows of the TIMED configuration are compared with the CPU oracle's
           last-row restatement (paged_infer.c:182-236) on the same pages; the run FAILS above 1e-5
cpu_baseline / --impl reference = the reference's own CPU code (oracle/_ref, compiled from
           /root/reference) timed on this box's host cores on a bounded sample of the same workload

One process per GPU under torchrun for N>1 (sequences sharded; the attention path has no collective; the
whole-model section gathers the sampled tokens of all ranks with NCCL inside the library: pa_group_*).
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
VERIFY_TOL = 1e-5          # north_star: max relative error 1e-5 (max|a-b| / max|ref|)
TOTAL_XL_BATCH = 512       # BASELINE configs[3]: batch 512 sharded over the GPUs

WORKLOADS = {
    # BASELINE.json configs[1] + the 12 layers of GPT-2 124M
    "cfg2": dict(name="GPT-2 124M shape (12 layers, 12 heads, head_dim 64), 64 sequences x 1024 ctx, block 16, "
                      "shuffled block tables, KV append + paged decode attention per layer",
                 NH=12, hs=64, bs=16, L=12, B=64, ctx="fixed", ctx_len=1024),
    # BASELINE.json configs[2]
    "cfg3": dict(name="GPT-2 124M paged decode, batch 256, mixed ctx U{128..1024}, block 16, KV append + attention",
                 NH=12, hs=64, bs=16, L=12, B=256, ctx="uniform", ctx_lo=128, ctx_hi=1024),
    # BASELINE.json configs[3]: all 48 layers; batch 512 / N sequences per GPU (N = 2, 4, 8); one GPU cannot hold the
    # 322 GB of KV of 512 sequences and runs the 2-GPU shard (256 sequences, 161 GB)
    "cfg4": dict(name="GPT-2 XL shape (48 layers, 25 heads, head_dim 64) paged decode, batch 512 sharded over the GPUs, "
                      "1024 ctx, block 16",
                 NH=25, hs=64, bs=16, L=48, B=None, ctx="fixed", ctx_len=1024),
    # BASELINE.json configs[4], one GPU's share: 32k contexts, head_dim 128; the prompt is prefilled through the paged
    # layout in chunks, then decoded
    "cfg5": dict(name="long-context paged attention, 8 sequences x 32768 ctx per GPU, 32 heads x head_dim 128, block 16, "
                      "1 layer: chunked prompt prefill, then decode",
                 NH=32, hs=128, bs=16, L=1, B=8, ctx="fixed", ctx_len=32768, prefill_chunk=2048),
    # BASELINE.json configs[0] (L2-resident, latency-bound)
    "cfg1": dict(name="GPT-2 124M paged decode batch 1, block 16, ctx 256 (L2-resident)",
                 NH=12, hs=64, bs=16, L=12, B=1, ctx="fixed", ctx_len=256),
}
ALIASES = {"xl": "cfg4", "long": "cfg5"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops"]), "measured dense bf16 (MEASURED_PEAKS.json bf16_tflops)"
        except Exception:
            pass
    return 1650.0, "fallback (B200_PROFILING.md)"


def resolve_workload(wid, world):
    wid = ALIASES.get(wid, wid)
    w = dict(WORKLOADS[wid])
    if wid == "cfg4":
        w["B"] = TOTAL_XL_BATCH // max(world, 2)
        w["shard"] = (f"{TOTAL_XL_BATCH} sequences / {world} GPUs" if world >= 2 else
                      "one GPU cannot hold batch 512 (322 GB of KV): it runs the 2-GPU shard, 256 sequences = 161 GB")
    return wid, w


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


def config_of(wid, w, world):
    """The `config` object of the JSON line: identical for this arm and for --impl reference."""
    ctx = context_lengths(w, 0)
    step_bytes = w["L"] * (decode_bytes(ctx, w["NH"] * w["hs"], w["bs"]) + append_bytes(w["B"], w["NH"] * w["hs"]))
    return {"workload": w["name"], "workload_id": wid, "layers": w["L"], "batch_per_gpu": w["B"],
            "ctx_mean": sum(ctx) / len(ctx), "block_size": w["bs"],
            "l2": f"inputs larger than L2: {step_bytes / 1e9:.2f} GB streamed per step vs 126 MB L2" if step_bytes > 4e8 else
                  f"L2-resident by definition of the config ({step_bytes / 1e6:.1f} MB per step): latency-bound, reported as such",
            "parallelism": f"sequences sharded over {world} GPU(s), no data-path collective"}


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
def reference_arm(args, wid, w, quiet=False):
    """Times the reference's own CPU implementation (oracle/_ref, else the oracle port) of the path on the
    host cores.  Each 'step' is a BOUNDED SAMPLE of the workload's step -- one sequence x one layer of it:
    add_to_cache(n_tail=1) + collect_kv_blocks + attention_paged over the full T=ctx window, exactly what
    paged_infer.c:706-715 executes per generated token (the reference recomputes all T rows) -- and the
    value is the same normalised metric (algorithmic GB/s of the path).  Loads nothing of the product."""
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
        sample = (f"bounded sample of the step: 1 of the {w['B']} sequences x 1 of the {w['L']} layers per step; the reference's "
                  f"attention_paged + add_to_cache as written (full T={T} window recomputed per decode step, "
                  f"paged_infer.c:706-715), -O3 -Ofast -fopenmp, all host threads")
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
        sample = (f"bounded sample of the step: 1 of the {w['B']} sequences x 1 of the {w['L']} layers per step; oracle port of "
                  f"attention_paged (full T={T} window)")
    for _ in range(args.warmup):
        one_step()
    times = [one_step()[0] for _ in range(args.steps)]
    t = sum(times) / len(times)
    v = bytes_per / t / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wid, w, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "tokens_per_s": 1.0 / t}
    if not quiet:
        emit(line)
    return line


def cpu_baseline_sample(wid, w, budget_s=15.0):
    """Bounded CPU sample for the default run (rank 0, N=1): reference as written + last-row port."""
    import oracle_api as oa

    class A:
        pass
    a = A()
    a.gpus, a.steps, a.warmup = 1, 1, 1
    first = reference_arm(a, wid, w, quiet=True)
    per = first["ms_per_step"] / 1e3
    a.steps = int(max(2, min(40, budget_s * 0.6 / max(per, 1e-3))))
    a.warmup = 0
    line = reference_arm(a, wid, w, quiet=True)
    cb = dict(line["cpu_baseline"])
    cb["steps"] = a.steps
    # last-row restatement (what a decode step actually needs) on a slice of the batch
    NH, hs, bs = w["NH"], w["hs"], w["bs"]
    C_ = NH * hs
    T = w["ctx_len"] if w["ctx"] == "fixed" else (w["ctx_lo"] + w["ctx_hi"]) // 2
    T = min(T, 4096)
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
                           "sample": f"oracle last-row restatement (paged_infer.c:182-236 for t=T-1), {nseq} sequences x 1 layer at ctx {T}, best of 5",
                           "cores": orc.lib.orc_omp_threads()}
    orc.close()
    return cb


def cpu_model_sample(model, pa_mod, B, L, NH, C_, V, maxT, bs, steps=4):
    """Whole-model decode on the host cores beside the device number: the oracle's restatement of gpt2_forward
    (oracle/paged_oracle.c orc_model_decode_step, all layers, -O3 -Ofast -fopenmp) on the SAME weights,
    `steps` decode steps of B sequences from an empty cache (the projections are >95 % of its work at
    these context lengths).  Bounded: only offered for small batches."""
    import oracle_api as oa
    params = download_params(model, pa_mod)
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


def download_params(model, pa_mod):
    lib = pa_mod.load()
    n = int(model.n_params)
    params = np.empty(n, dtype=np.float32)
    pa_mod.check(lib.pa_memcpy_d2h(params.ctypes.data, lib.pa_model_params(model.m), n * 4, None), "params d2h")
    pa_mod.check(lib.pa_device_sync(), "sync")
    return params


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


class Ctx:
    """Per-process state shared by the workloads of one bench invocation."""

    def __init__(self, args, pa, lib, dist, rank, local_rank, world):
        self.args, self.pa, self.lib, self.dist = args, pa, lib, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.group_id = None

    def barrier(self):
        self.pa.check(self.lib.pa_device_sync(), "sync")
        if self.dist is not None:
            self.dist.barrier()
            self.pa.check(self.lib.pa_device_sync(), "sync")

    def allmax(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{self.local_rank}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def unique_id(self):
        """128-byte NCCL id made by rank 0 inside the library, handed to the other ranks by the launcher's own
        rendezvous (torch.distributed here; an MPI host would MPI_Bcast it)."""
        import torch
        buf = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            raw = (C.c_ubyte * 128)()
            self.pa.check(self.lib.pa_comm_unique_id(raw), "pa_comm_unique_id")
            buf = torch.tensor(list(raw), dtype=torch.uint8)
        buf = buf.to(f"cuda:{self.local_rank}")
        self.dist.broadcast(buf, src=0)
        return bytes(buf.cpu().tolist())


def fill_pools(cx, eng, L, n_floats, seed):
    """Random-init K/V pools written by the device (N(0,1) from a counter-based hash): no host copy, any size."""
    lib, pa = cx.lib, cx.pa
    for layer in range(L):
        pa.check(lib.pa_fill_normal(eng.pool_k(layer), n_floats, 1.0, 0.0, seed + 2 * layer, None), "fill K")
        pa.check(lib.pa_fill_normal(eng.pool_v(layer), n_floats, 1.0, 0.0, seed + 2 * layer + 1, None), "fill V")
    pa.check(lib.pa_device_sync(), "sync")


def oracle_row(cx, eng, layer, s, q_row, NH, hs, bs):
    """Last-row restatement of attention_paged (paged_infer.c:182-236) for sequence s on the CPU oracle, fed with the
    pages the device holds for it in `layer` (downloaded through the block table): the fp32 restatement (C,) and the
    same arithmetic carried out in fp64 (the exact answer the fp32 reference approximates)."""
    import oracle_api as oa
    lib, pa = cx.lib, cx.pa
    C_ = NH * hs
    table = list(eng.table(s))
    n_tok = eng.seq_len(s)
    page_bytes = bs * C_ * 4
    orc = oa.OrcManager(C_, bs, len(table) + 1, 1)
    try:
        pk, pv = eng.pool_k(layer), eng.pool_v(layer)
        # coalesce physically consecutive pages into one copy each
        kbuf = np.empty((len(table), bs, C_), dtype=np.float32)
        vbuf = np.empty((len(table), bs, C_), dtype=np.float32)
        for j, idx in enumerate(table):
            pa.check(lib.pa_memcpy_d2h(kbuf[j].ctypes.data, pk + idx * page_bytes, page_bytes, None), "d2h")
            pa.check(lib.pa_memcpy_d2h(vbuf[j].ctypes.data, pv + idx * page_bytes, page_bytes, None), "d2h")
        pa.check(lib.pa_device_sync(), "sync")
        for j in range(len(table)):
            oidx = orc.request_block(0)
            k, v = orc.page_arrays(oidx)
            n = min(bs, n_tok - j * bs)
            k[:n], v[:n] = kbuf[j, :n], vbuf[j, :n]
            orc.set_filled(oidx, n)
        q1 = np.ascontiguousarray(q_row[None, :C_])
        return orc.decode_batch([0], NH, q1)[0], orc.decode_batch_f64([0], NH, q1)[0]
    finally:
        orc.close()


def rel_err(got, want):
    return float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max() / max(np.abs(want).max(), 1e-30))


def row_err(got, want32_64):
    """(error against the fp32 oracle, error against its fp64 evaluation).  The bar ("tolerance
    evaluated against both the fp32 last-row restatement and the fp64 restatement"): within VERIFY_TOL of the fp32
    reference -- or, where the reference's own sequential fp32 sums (tens of thousands of terms at 32k context) sit
    further than that from the exact result, within VERIFY_TOL of the exact result."""
    return rel_err(got, want32_64[0]), rel_err(got, want32_64[1])


def run_workload(cx, wid, w, primary):
    """One workload on this rank's GPU.  Returns a dict of results (rank-local values already reduced with max over
    ranks where they are times)."""
    args, pa, lib, dist, rank, world = cx.args, cx.pa, cx.lib, cx.dist, cx.rank, cx.world
    NH, hs, bs, L, B = w["NH"], w["hs"], w["bs"], w["L"], w["B"]
    C_ = NH * hs
    ctx = context_lengths(w, rank)
    pages = [(c + bs - 1) // bs for c in ctx]
    n_blocks = sum(pages) + 64
    steps = args.steps if primary else max(3, min(args.steps, 50 if B * L <= 64 else (10 if L >= 12 else 20)))
    warmup = args.warmup if primary else 3
    res = {"workload": w["name"]}
    chunk = w.get("prefill_chunk", 0)
    eng = pa.PagedAttn(bs, n_blocks, B, NH, hs, n_layers=L, device=cx.local_rank, max_batch_tokens=max(B, B * chunk))
    try:
        if primary:
            eng.tune(pa.PA_TUNE_HEADS_PER_TILE, args.hpg)
            eng.tune(pa.PA_TUNE_STAGES, args.stages)
            eng.tune(pa.PA_TUNE_GRID, args.grid)
            eng.tune(pa.PA_TUNE_STATIC_PCT, args.static_pct)
            eng.tune(pa.PA_TUNE_DYN_UNITS, args.dyn_units)
        eng.tune(pa.PA_TUNE_NO_PDL, 1 if args.no_pdl else 0)
        eng.tune(pa.PA_TUNE_NO_ZEROCOPY, args.host_mode)
        rng = np.random.default_rng(1234 + rank)
        stream = lib.pa_stream_create()
        seq_ids = np.arange(B, dtype=np.int32)
        ones = np.ones(B, dtype=np.int32)
        perm = rng.permutation(sum(pages))        # fragmented: a seeded permutation dealt to the sequences

        # ---- synthetic state ----------------------------------------------------------------
        if chunk:
            # cfg5: the cache is PRODUCED by prefilling the prompts through the paged layout (chunks of `chunk` tokens
            # per sequence per step: KV append + causal multi-row attention), timed; decode then runs on that cache
            res["prefill"] = prefill_section(cx, eng, w, ctx, perm, pages, stream)
        else:
            fill_pools(cx, eng, L, n_blocks * bs * C_, seed=977 * (rank + 1))
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

        # host buffers of the e2e entry: every layer has its own q|k|v rows (L x B x 3C) and its own output rows
        # (L x B x C) in pinned memory, so h2d/d2h_bytes_per_step are bytes that really cross PCIe each step; the
        # rows are the same for every layer (and equal to the device-resident arm's), so one oracle row checks both
        qkv_host = lib.pa_host_alloc(L * B * 3 * C_ * 4)
        out_host = lib.pa_host_alloc(L * B * C_ * 4)
        qkv_all = np.ctypeslib.as_array(C.cast(qkv_host, C.POINTER(C.c_float)), (L, B, 3 * C_))
        qkv_all[:] = rng.standard_normal((B, 3 * C_), dtype=np.float32)[None]
        qkv_np = qkv_all[0]
        out_all = np.ctypeslib.as_array(C.cast(out_host, C.POINTER(C.c_float)), (L, B, C_))
        out_np = out_all[L - 1]
        d_qkv = pa.DevBuf.from_numpy(qkv_np)
        d_out = pa.DevBuf(B * C_ * 4)

        def rollback():
            pa.check(eng.step_rollback(), "rollback")

        # kernel time: events around the L back-to-back decode launches of a step (no event between
        # launches: that would serialise them and defeat programmatic dependent launch)
        evs = [lib.pa_event_create() for _ in range(2)]
        dec_ms = []

        def step(timed_kernels=False, keep=False):
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
            if not keep:
                rollback()

        sampler = ClockSampler(cx.local_rank) if (rank == 0 and primary) else None
        for _ in range(warmup):
            step()
        cx.barrier()
        launches0 = eng.launches()
        e0, e1 = lib.pa_event_create(), lib.pa_event_create()
        # the kernel events are recorded in every timed step and read back every 8th (reading needs a
        # sync, which would serialise host and device if done each step)
        lib.pa_event_record(e0, stream)
        for i in range(steps):
            step(timed_kernels=True)
            if i % 8 == 7 or i == steps - 1:
                dec_ms.append(lib.pa_event_elapsed_ms(evs[0], evs[1]) / L)
        lib.pa_event_record(e1, stream)
        cx.barrier()
        ms_total = cx.allmax(lib.pa_event_elapsed_ms(e0, e1))
        launches = eng.launches() - launches0
        if primary:
            # keep the same load on for >= 1.5 s in total so the clock sampler sees it (untimed)
            t_end = time.time() + (0.0 if args.no_clock_hold else max(0.0, 1.5 - ms_total / 1e3))
            while time.time() < t_end:
                step()
                lib.pa_stream_sync(stream)
            cx.barrier()
        res["clocks"] = sampler.stop() if sampler else None

        if primary and args.timeline and rank == 0:
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
        ms_per_step = ms_total / steps
        res.update(value=world * step_bytes / (ms_per_step * 1e-3) / 1e9, ms_per_step=ms_per_step, steps=steps, warmup=warmup,
                   tokens_per_s=world * B / (ms_per_step * 1e-3), gpu_launches=int(launches), step_bytes=step_bytes,
                   batch_per_gpu=B, layers=L, ctx_mean=sum(ctx) / len(ctx))

        # ---- e2e: host buffers through the host-buffer entry of the C ABI --------------------------
        # Every step: pa_step_begin (tables), ONE call queueing all layers (the kernels pull the pinned host q|k|v rows
        # over PCIe and store their outputs to pinned host memory), a completion ticket, and the host READS the step's
        # result.  The host queues step n+1 before it waits for step n's ticket (the attention path's inputs are the
        # host's rows, never its own previous outputs), so the device does not idle during the host's turnaround; the
        # one-sync-per-step variant is timed beside it.
        e2e_step = None
        if not args.no_e2e:
            out_host2 = lib.pa_host_alloc(L * B * C_ * 4)
            out_np2 = np.ctypeslib.as_array(C.cast(out_host2, C.POINTER(C.c_float)), (L, B, C_))[L - 1]
            outs = [(out_host, out_np), (out_host2, out_np2)]
            in_stride, out_stride = B * 3 * C_, B * C_

            def e2e_queue(buf):
                pa.check(eng.step_begin(seq_ids, ones), "step_begin")
                pa.check(lib.pa_decode_step_host_layers_async(eng.h, qkv_host, in_stride, outs[buf][0], out_stride), "decode_step_host_layers_async")
                t = lib.pa_decode_step_host_mark(eng.h)
                pa.check(min(t, 0), "mark")
                rollback()                      # host-side integer state only: the queued kernels carry their tables
                return t

            def e2e_step(keep=False):           # one step, waited for at once (verification, and the per-step-sync timing)
                pa.check(eng.step_begin(seq_ids, ones), "step_begin")
                pa.check(lib.pa_decode_step_host_layers_async(eng.h, qkv_host, in_stride, out_host, out_stride), "decode_step_host_layers_async")
                pa.check(lib.pa_decode_step_host_sync(eng.h), "sync")
                if not keep:
                    rollback()
                return float(out_np[0, 0])      # the step's result is read on the host

            def e2e_run(k, pipelined):
                hs_ = lib.pa_stream_of(eng.h)
                cx.barrier()
                lib.pa_event_record(e0, hs_)
                t0 = time.perf_counter()
                acc = 0.0
                if pipelined:
                    prev = None
                    for i in range(k):
                        t = e2e_queue(i & 1)
                        if prev is not None:
                            pa.check(lib.pa_decode_step_host_wait(eng.h, prev[0]), "wait")
                            acc += float(outs[prev[1]][1][0, 0])         # step i-1's result, read while step i runs
                        prev = (t, i & 1)
                    pa.check(lib.pa_decode_step_host_wait(eng.h, prev[0]), "wait")
                    acc += float(outs[prev[1]][1][0, 0])
                else:
                    for _ in range(k):
                        acc += e2e_step()
                lib.pa_event_record(e1, hs_)
                cx.barrier()
                wall = time.perf_counter() - t0
                return cx.allmax(max(lib.pa_event_elapsed_ms(e0, e1), 0.0)) / k, wall * 1e3 / k
            # its own step count (reported as e2e.steps): enough steps that filling and draining the one-step-ahead pipeline
            # does not dominate, about 0.1 s of device time at most
            k_e2e = int(max(10, min(100, 0.1 / max(ms_per_step * 1e-3, 1e-6))))
            for _ in range(3):
                e2e_step()
            e2e_run(4, True)
            ms_sync, wall_sync = e2e_run(k_e2e, False)
            ms_e2e, wall = e2e_run(k_e2e, True)
            table_bytes = (4 * B + 2 + B + 4 + B * ((max(pages) + 3) & ~3)) * 4
            res["e2e"] = {"value": world * step_bytes / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
                          "h2d_bytes_per_step": L * B * 3 * C_ * 4 + table_bytes, "d2h_bytes_per_step": L * B * C_ * 4,
                          "ms_per_step": ms_e2e, "steps": k_e2e, "tokens_per_s": world * B / (ms_e2e * 1e-3),
                          "entry": "per step: pa_step_begin + pa_decode_step_host_layers_async (one call queues every layer) + "
                                   "pa_decode_step_host_mark; the host queues step n+1, then waits for step n's ticket "
                                   "(pa_decode_step_host_wait) and reads its output rows; " +
                                   {0: ("the step's pinned host q|k|v rows (every layer's) go H2D in one copy and its output rows D2H in one "
                                        "copy, on copy streams beside the kernels (double buffered)" if L * B * 3 * C_ * 4 / 1e6 * 4.0 < L * 7.0 else
                                        "host q|k|v rows in pinned memory are pulled over PCIe by the kernel's bulk copies, outputs stored to "
                                        "pinned host memory by the kernel (zero-copy: chosen by the library for steps whose staged copies "
                                        "would cost more HBM write interference than the per-layer PCIe latency they save)"),
                                    1: "per-layer staged copies on copy streams",
                                    2: "host q|k|v rows in pinned memory are pulled over PCIe by the kernel's bulk copies, outputs stored "
                                       "to pinned host memory by the kernel (zero-copy)",
                                    3: "whole-step staged copies"}[args.host_mode],
                          "wall_ms_per_step": wall,
                          "sync_per_step": {"value": world * step_bytes / (ms_sync * 1e-3) / 1e9, "ms_per_step": ms_sync,
                                            "wall_ms_per_step": wall_sync,
                                            "entry": "the same with one full stream synchronisation per step before the next is queued"}}
            res["e2e"]["frac_of_measured_peak"] = res["e2e"]["value"] / world / peaks()[0]
        else:
            res["e2e"] = None

        # ---- verified: output rows of THIS configuration against the CPU oracle ------------------
        if not args.no_verify:
            res["verified"] = verify_attention(cx, eng, w, ctx, step, e2e_step, rollback, qkv_np, d_out, out_np)

        # ---- roofline of the dominant kernel ---------------------------------------------------
        peak, peak_src = peaks()
        kbytes = decode_bytes(ctx, C_, bs) + append_bytes(B, C_)      # per layer: append + decode
        kms = sum(dec_ms) / len(dec_ms)
        achieved = kbytes / (kms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):       # the ncu --set full figure of this kernel AT THIS SHAPE (entries carry the shape's algorithmic bytes)
            try:
                for key, ent in json.load(open(tpath)).items():
                    if isinstance(ent, dict) and abs(ent.get("algorithmic_bytes_per_launch", 0) - kbytes) <= 0.01 * kbytes:
                        traffic = ent.get("dram_bytes_per_launch")
                        break
            except Exception:
                traffic = None
        last_ctas = lib.pa_tune_get(eng.h, 12)       # 0: the launch did not go through the stream kernel
        kname = (f"pa_decode_stream_kernel<{hs},{bs}>" if last_ctas else
                 "pa_decode_small_kernel (one CTA per sequence and head; chosen below ~98 k token-heads, latency-bound)")
        res["roofline"] = {"bound": "hbm", "kernel": kname + ("" if args.no_fuse else " (KV append fused)"), "achieved": achieved,
                           "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": kbytes, "avg_launch_ms": kms, "launches_timed": len(dec_ms) * L,
                           "timing": f"CUDA events around the {L} back-to-back per-layer launches of a step, / {L}",
                           "frac_of_nominal_8TBps": achieved / 8000.0,
                           "kernel_share_of_step": kms * L / ms_per_step,
                           "plan": {"heads_per_tile": lib.pa_tune_get(eng.h, 10), "ring_stages": lib.pa_tune_get(eng.h, 11),
                                    "ctas": lib.pa_tune_get(eng.h, 12)}}
        res["frac_of_measured_peak"] = res["value"] / world / peak

        # ---- whole-model decode step -------------------------------------------------
        if not args.no_model and hs == 64 and max(ctx) <= 4096 and L <= 12:
            try:
                res["model"] = model_section(cx, eng, w, ctx, kms)
            except Exception as ex:
                res["model"] = {"error": repr(ex)}
        lib.pa_host_free(qkv_host)
        lib.pa_host_free(out_host)
        if not args.no_e2e:
            lib.pa_host_free(out_host2)
        lib.pa_stream_destroy(stream)
    finally:
        eng.close()
    return res


def verify_attention(cx, eng, w, ctx, step, e2e_step, rollback, qkv_np, d_out, out_np):
    """One more step of the timed configuration WITHOUT the roll-back; the output rows of up to three sequences
    (longest, shortest, first) of the LAST layer are compared with the CPU oracle fed with the very pages the device
    holds (so the fused KV append is checked too).  Same check on the host-buffer (e2e) entry's output rows."""
    NH, hs, bs, L = w["NH"], w["hs"], w["bs"], w["L"]
    C_ = NH * hs
    pick = sorted({int(np.argmax(ctx)), int(np.argmin(ctx)), 0})
    out = {"tolerance": VERIFY_TOL, "checker": "CPU oracle, last-row restatement of attention_paged (paged_infer.c:182-236), "
                                               "on the pages downloaded from the device pool",
           "sequences": pick, "layer": L - 1}
    step(keep=True)
    cx.pa.check(cx.lib.pa_device_sync(), "sync")
    got = d_out.download((len(ctx), C_))
    want = {s: oracle_row(cx, eng, L - 1, s, qkv_np[s], NH, hs, bs) for s in pick}
    rollback()
    errs = [row_err(got[s], want[s]) for s in pick]
    out["max_rel_err"] = cx.allmax(max(e[0] for e in errs))
    out["max_rel_err_vs_fp64"] = cx.allmax(max(e[1] for e in errs))
    out["reference_fp32_vs_fp64"] = max(rel_err(want[s][0], want[s][1]) for s in pick)
    worst = cx.allmax(max(min(e) for e in errs))
    if e2e_step is not None:
        e2e_step(keep=True)
        got_h = out_np.copy()
        rollback()
        errs_h = [row_err(got_h[s], want[s]) for s in pick]
        out["e2e_max_rel_err"] = cx.allmax(max(e[0] for e in errs_h))
        out["e2e_max_rel_err_vs_fp64"] = cx.allmax(max(e[1] for e in errs_h))
        worst = max(worst, cx.allmax(max(min(e) for e in errs_h)))
    if not (worst <= VERIFY_TOL):
        raise SystemExit(f"bench.py: VERIFICATION FAILED for {w['name']}: max rel err {worst:.3e} > {VERIFY_TOL}")
    return out


def prefill_section(cx, eng, w, ctx, perm, pages, stream):
    """cfg5: prefill the prompts through the paged layout.  Every sequence gets ctx-1 tokens in chunks (one step
    = `chunk` new tokens for each of the B sequences: KV-append kernel + causal multi-row paged attention over
    everything cached so far), block tables fragmented by pre-adopting nothing: pages come from the allocator
    as the prompt grows.  Timed with CUDA events; flops = 4*hs per (query row, visible key, head)."""
    args, pa, lib = cx.args, cx.pa, cx.lib
    NH, hs, bs, B, chunk = w["NH"], w["hs"], w["bs"], w["B"], w["prefill_chunk"]
    C_ = NH * hs
    n_tok = ctx[0] - 1                       # one short: the decode steps append the last one
    d_in = pa.DevBuf(B * chunk * 3 * C_ * 4)
    d_o = pa.DevBuf(B * chunk * C_ * 4)
    seq_ids = np.arange(B, dtype=np.int32)
    out = {}
    paths = [(0, "default: tcgen05 3xTF32, fp32-accurate (tolerance 1e-5)")]
    if not args.no_tc_prefill:
        paths.append((3, "tcgen05 TF32 (opt-in, tolerance 5e-3)"))
    e0, e1 = lib.pa_event_create(), lib.pa_event_create()
    hstream = lib.pa_stream_of(eng.h)
    for path, label in paths:
        for s in range(B):
            eng.seq_free(s)
        eng.tune(pa.PA_TUNE_PREFILL_PATH, path)
        done, total_ms, keys_seen, chunk_i = 0, 0.0, 0, 0
        check = None
        while done < n_tok:
            n = min(chunk, n_tok - done)
            # fresh N(0,1) q|k|v rows for this chunk, written by the device
            pa.check(lib.pa_fill_normal(d_in.ptr, B * n * 3 * C_, 1.0, 0.0, 31337 + 7919 * chunk_i + cx.rank, hstream), "fill")
            pa.check(eng.step_begin(seq_ids, np.full(B, n, dtype=np.int32)), "step_begin")
            pa.check(eng.upload(hstream), "upload")
            lib.pa_event_record(e0, hstream)
            pa.check(eng.append(0, d_in.ptr + C_ * 4, d_in.ptr + 2 * C_ * 4, 3 * C_, hstream), "append")
            pa.check(eng.prefill(0, d_in.ptr, 3 * C_, d_o.ptr, C_, hstream), "prefill")
            lib.pa_event_record(e1, hstream)
            pa.check(lib.pa_stream_sync(hstream), "sync")
            total_ms += lib.pa_event_elapsed_ms(e0, e1)
            keys_seen += B * sum(done + j + 1 for j in range(n))
            done += n
            chunk_i += 1
            if done >= n_tok and not args.no_verify:
                # the LAST prompt row of sequence 0 against the oracle's last-row restatement on the cache just built
                rows = d_in.download((B * n, 3 * C_))
                outs = d_o.download((B * n, C_))
                want = oracle_row(cx, eng, 0, 0, rows[n - 1], NH, hs, bs)
                check = min(row_err(outs[n - 1], want))
        tol = VERIFY_TOL if path != 3 else 5e-3
        if check is not None and not (check <= tol):
            raise SystemExit(f"bench.py: VERIFICATION FAILED for the cfg5 prefill ({label}): {check:.3e} > {tol}")
        ms = cx.allmax(total_ms)
        flops = 4.0 * hs * NH * keys_seen
        tp, tp_src = tensor_peak()
        ent = {"path": label, "ms": ms, "prompt_tokens_per_gpu": B * n_tok, "chunk_tokens": chunk,
               "tflops_per_gpu": flops / (ms * 1e-3) / 1e12, "prompt_tokens_per_s": cx.world * B * n_tok / (ms * 1e-3),
               "verified_last_row_max_rel_err": check, "tolerance": tol,
               "verified_against": "the closer of the oracle's fp32 last-row restatement and its fp64 evaluation (32k-term fp32 sums)"}
        ent["frac_of_dense_tf32_peak"] = ent["tflops_per_gpu"] / (tp / 2.0)
        ent["peak_note"] = (f"dense TF32 peak taken as half of the {tp_src} = {tp / 2:.0f} TFLOP/s" +
                            ("; the 3xTF32 split issues three MMAs per product, so its ceiling is a third of that" if path == 0 else ""))
        out["fp32" if path == 0 else "tf32"] = ent
    eng.tune(pa.PA_TUNE_PREFILL_PATH, 0)
    d_in.free()
    d_o.free()
    return out


def prefill_headline(cx):
    """Prompt prefill at the GPT-2 124M head shape (12 heads x 64, block 16): 16 prompts x 2048 tokens, one layer,
    KV append + causal multi-row paged attention per pass; the default path (tcgen05 3xTF32, fp32-accurate), the fp32
    SIMT kernel it replaced as default, and the opt-in plain-TF32 kernel.  The last prompt row of sequence 0 is checked
    against the oracle for each."""
    args, pa, lib = cx.args, cx.pa, cx.lib
    NH, hs, bs, B, T = 12, 64, 16, 16, 2048
    C_ = NH * hs
    pages = (T + bs - 1) // bs + 1
    eng = pa.PagedAttn(bs, B * pages + 8, B, NH, hs, n_layers=1, device=cx.local_rank, max_batch_tokens=B * T)
    out = {"shape": f"{B} prompts x {T} tokens, {NH} heads x head_dim {hs}, block {bs}, 1 layer (append + prefill per pass)"}
    try:
        d_in, d_o = pa.DevBuf(B * T * 3 * C_ * 4), pa.DevBuf(B * T * C_ * 4)
        hstream = lib.pa_stream_of(eng.h)
        pa.check(lib.pa_fill_normal(d_in.ptr, B * T * 3 * C_, 1.0, 0.0, 4242 + cx.rank, hstream), "fill")
        seq_ids = np.arange(B, dtype=np.int32)
        e0, e1 = lib.pa_event_create(), lib.pa_event_create()
        keys_seen = B * T * (T + 1) // 2
        flops = 4.0 * hs * NH * keys_seen
        tp, tp_src = tensor_peak()
        for path, key, label, tol in ((0, "default", "tcgen05 3xTF32, fp32-accurate (automatic choice)", VERIFY_TOL),
                                      (1, "simt", "tiled fp32 SIMT (PA_TUNE_PREFILL_PATH=1)", VERIFY_TOL),
                                      (3, "tf32", "tcgen05 plain TF32 (opt-in, PA_TUNE_PREFILL_PATH=3)", 5e-3)):
            eng.tune(pa.PA_TUNE_PREFILL_PATH, path)
            ms = []
            check = None
            for it in range(3 + 5):
                pa.check(eng.step_begin(seq_ids, np.full(B, T, dtype=np.int32)), "step_begin")
                pa.check(eng.upload(hstream), "upload")
                lib.pa_event_record(e0, hstream)
                pa.check(eng.append(0, d_in.ptr + C_ * 4, d_in.ptr + 2 * C_ * 4, 3 * C_, hstream), "append")
                pa.check(eng.prefill(0, d_in.ptr, 3 * C_, d_o.ptr, C_, hstream), "prefill")
                lib.pa_event_record(e1, hstream)
                pa.check(lib.pa_stream_sync(hstream), "sync")
                if it >= 3:
                    ms.append(lib.pa_event_elapsed_ms(e0, e1))
                if it == 7 and not args.no_verify:
                    q_last = d_in.download((1, 3 * C_), offset_bytes=(T - 1) * 3 * C_ * 4)[0]
                    o_last = d_o.download((1, C_), offset_bytes=(T - 1) * C_ * 4)[0]
                    check = min(row_err(o_last, oracle_row(cx, eng, 0, 0, q_last, NH, hs, bs)))
                pa.check(eng.step_rollback(), "rollback")
            if check is not None and not (check <= tol):
                raise SystemExit(f"bench.py: VERIFICATION FAILED for the prefill ({label}): {check:.3e} > {tol}")
            t = cx.allmax(float(np.median(ms)))
            out[key] = {"path": label, "ms": t, "tflops_per_gpu": flops / (t * 1e-3) / 1e12,
                        "prompt_tokens_per_s": cx.world * B * T / (t * 1e-3), "verified_last_row_max_rel_err": check, "tolerance": tol,
                        "frac_of_dense_tf32_peak": flops / (t * 1e-3) / 1e12 / (tp / 2.0)}
        out["peak_note"] = (f"dense TF32 peak taken as half of the {tp_src} = {tp / 2:.0f} TFLOP/s; the 3xTF32 split issues three "
                            f"MMAs per product (ceiling: a third of it); fp32 FFMA peak for the SIMT kernel: 74.5 TFLOP/s")
        out["flops"] = "4 * head_dim per (query row, visible key, head): QK^T and PV, causal"
        d_in.free()
        d_o.free()
    finally:
        eng.close()
    return out


def model_section(cx, eng, w, ctx, kms):
    """embedding, L x {ln, QKV+append, paged attention, attproj, ln, MLP}, LM head, sampler through
    pa_model_decode_step (host tokens in, host tokens out).  With N > 1 the step goes through pa_group_model_step:
    the sampled tokens of all ranks are all-gathered by NCCL INSIDE the library, enqueued behind the sampler on the
    handle's stream (no host synchronisation between the step and the collective)."""
    args, pa, lib, dist, rank, world = cx.args, cx.pa, cx.lib, cx.dist, cx.rank, cx.world
    NH, hs, bs, L, B = w["NH"], w["hs"], w["bs"], w["L"], w["B"]
    C_ = NH * hs
    V, maxT = 50257, max(1024, max(ctx) + 8)
    rng = np.random.default_rng(99 + rank)
    model = pa.Model(eng, maxT, V, params=None, seed=1337, max_batch=B)       # the same weights on every rank (replicated)
    group = None
    try:
        seq_ids = np.arange(B, dtype=np.int32)
        toks = rng.integers(0, V, size=B).astype(np.int32)
        coins = rng.random(B).astype(np.float32)
        all_next = np.zeros(world * B, dtype=np.int32)
        if dist is not None:
            group = C.c_void_p()
            pa.check(lib.pa_group_join(eng.h, cx.unique_id(), rank, world, C.byref(group)), "pa_group_join")
            models = (C.c_void_p * 1)(model.m)
            seqs = (pa.c_int_p * 1)(pa.iptr(seq_ids))
            tokp = (pa.c_int_p * 1)(pa.iptr(toks))
            coinp = (C.c_void_p * 1)(coins.ctypes.data)
            nxt_local = np.zeros(B, dtype=np.int32)
            nxtp = (pa.c_int_p * 1)(pa.iptr(nxt_local))

        def model_step(keep=False):
            if group is not None:
                # own tokens back after one wait; the all-gather of this step's tokens runs on a side stream beside the
                # next step and is handed out by the next call (all_next: the previous step's tokens of every rank)
                pa.check(min(lib.pa_group_model_step_overlapped(group, models, seqs, tokp, coinp, B, nxtp, pa.iptr(all_next)), 0),
                         "pa_group_model_step_overlapped")
                nxt = nxt_local
            else:
                nxt = model.decode_step(seq_ids, toks, coins)
            if not keep:
                pa.check(eng.step_rollback(), "rollback")
            return nxt
        for _ in range(3):
            model_step()
        cx.barrier()
        k_model = max(5, min(args.steps, 30))
        hs_ = lib.pa_stream_of(eng.h)
        e0, e1 = lib.pa_event_create(), lib.pa_event_create()
        l0 = eng.launches()
        lib.pa_event_record(e0, hs_)
        t0 = time.perf_counter()
        for _ in range(k_model):
            model_step()
        if group is not None:
            pa.check(min(lib.pa_group_gather_flush(group, pa.iptr(all_next)), 0), "pa_group_gather_flush")     # the last step's gather, inside the timed region
        lib.pa_event_record(e1, hs_)
        cx.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / k_model
        ms_model = cx.allmax(max(lib.pa_event_elapsed_ms(e0, e1) / k_model, wall_ms))
        per_step = (eng.launches() - l0) / k_model
        persistent = per_step <= (2.0 if group is not None else 1.0)
        if group is not None:
            # the blocking variant beside it: the gather on the handle's stream, waited for inside the step
            cx.barrier()
            t0 = time.perf_counter()
            for _ in range(k_model):
                pa.check(lib.pa_group_model_step(group, models, seqs, tokp, coinp, B, pa.iptr(all_next)), "pa_group_model_step")
                pa.check(eng.step_rollback(), "rollback")
            cx.barrier()
            blocking_ms = cx.allmax((time.perf_counter() - t0) * 1e3 / k_model)
        info = {"tokens_per_s": world * B / (ms_model * 1e-3), "ms_per_step": ms_model, "steps": k_model,
                "vocab": V, "layers": L, "gpu_launches_per_step": per_step,
                "path": ("ONE persistent cooperative kernel for the whole step (pa_model_mega.cu: weight-streaming "
                         "fp32 GEMV phases, chunked paged attention, grid barriers)" if persistent else
                         "chain of per-op kernels (layernorm, tcgen05 3xTF32 projections, paged decode attention, sampler)"),
                "sampler": "softmax + multinomial (sample_mult) fused, host coins",
                "projections": ("fp32 FMA weight-streaming GEMV with bias/GELU/residual epilogues" if persistent else
                                "tcgen05 3xTF32 (fp32-accurate) GEMMs with bias/GELU/residual epilogues"),
                "weights": "random init on the device (no checkpoint offline)",
                "token_gather": (f"ncclAllGather of int32 next tokens inside the library (pa_group_model_step_overlapped, NCCL "
                                 f"{lib.pa_nccl_version()}), every step, on a side stream behind the sampler: a rank's next step consumes "
                                 f"only its own tokens, the gathered tokens of step n are handed out while step n+1 runs"
                                 if group is not None else None),
                "entry": ("pa_group_model_step_overlapped" if group is not None else "pa_model_decode_step") +
                         " (host token ids in, host token ids out, one sync per step)"}
        if group is not None:
            info["blocking_gather_ms_per_step"] = blocking_ms
            info["overlapped_gather_ms_per_step"] = ms_model
            if blocking_ms < ms_model:       # a handful of sequences (one persistent kernel per step, which owns every SM): the
                # side-stream collective only gets in its way -- the blocking pa_group_model_step is the faster entry there
                info.update(ms_per_step=blocking_ms, tokens_per_s=world * B / (blocking_ms * 1e-3),
                            entry="pa_group_model_step (host token ids in, host token ids out, gather on the handle's stream, one sync per step)")
        if not persistent:
            info["attention_share_of_step"] = kms * L / ms_model
        # verified: the logits row of sequence 0 of a step of THIS configuration against the oracle's gpt2_forward
        # restatement fed with the device's own weights and that sequence's cached pages of every layer
        if not args.no_verify and rank == 0:
            # (a LOCAL step: the other ranks do not take part, so it must not contain the collective)
            info["verified"] = verify_model(cx, eng, model, w, ctx, lambda keep=False: model.decode_step(seq_ids, toks, coins),
                                            seq_ids, toks)
        if world == 1 and B <= 8 and not args.no_cpu_baseline:
            try:
                info["cpu_baseline"] = cpu_model_sample(model, pa, B, L, NH, C_, V, maxT, bs)
            except Exception as ex:      # a report, never a reason to lose the line
                info["cpu_baseline"] = {"tokens_per_s": None, "sample": f"failed: {ex!r}"}
        if dist is not None:
            dist.barrier()
        return info
    finally:
        if group is not None:
            lib.pa_group_destroy(group)
        model.close()


def verify_model(cx, eng, model, w, ctx, model_step, seq_ids, toks):
    import oracle_api as oa
    pa, lib = cx.pa, cx.lib
    NH, hs, bs, L = w["NH"], w["hs"], w["bs"], w["L"]
    C_ = NH * hs
    V, maxT = model.V, model.cfg.max_seq_len
    s = 0
    model_step(keep=True)
    got = model.logits(len(seq_ids))[s].copy()
    params = download_params(model, pa)
    table = list(eng.table(s))
    n_tok = eng.seq_len(s) - 1                 # the cache BEFORE this step (the oracle appends the new token itself)
    page_bytes = bs * C_ * 4
    ol = oa.load_oracle("fast")
    mgrs = [oa.OrcManager(C_, bs, len(table) + 1, 1, flavor="fast") for _ in range(L)]
    try:
        buf = np.empty((bs, C_), dtype=np.float32)
        for layer in range(L):
            pk, pv = eng.pool_k(layer), eng.pool_v(layer)
            for j, idx in enumerate(table):
                n = min(bs, n_tok - j * bs)
                if n <= 0:
                    break
                oidx = mgrs[layer].request_block(0)
                k, v = mgrs[layer].page_arrays(oidx)
                pa.check(lib.pa_memcpy_d2h(buf.ctypes.data, pk + idx * page_bytes, page_bytes, None), "d2h")
                pa.check(lib.pa_device_sync(), "sync")
                k[:n] = buf[:n]
                pa.check(lib.pa_memcpy_d2h(buf.ctypes.data, pv + idx * page_bytes, page_bytes, None), "d2h")
                pa.check(lib.pa_device_sync(), "sync")
                v[:n] = buf[:n]
                mgrs[layer].set_filled(oidx, n)
        arr = (C.c_void_p * L)(*[m.m for m in mgrs])
        want = np.zeros((1, V), dtype=np.float32)
        seq0 = np.zeros(1, dtype=np.int32)
        tok0 = np.array([toks[s]], dtype=np.int32)
        pos0 = np.array([n_tok], dtype=np.int32)
        rc = ol.orc_model_decode_step(arr, L, NH, C_, V, maxT, oa.fptr(params), oa.iptr(seq0), oa.iptr(tok0), oa.iptr(pos0), 1,
                                      oa.fptr(want))
        if rc != 0:
            raise RuntimeError(f"orc_model_decode_step rc={rc}")
    finally:
        for m in mgrs:
            m.close()
        pa.check(eng.step_rollback(), "rollback")
    err = rel_err(got, want[0])
    tol = VERIFY_TOL * L
    if not (err <= tol):
        raise SystemExit(f"bench.py: VERIFICATION FAILED for the whole-model step of {w['name']}: logits err {err:.3e} > {tol}")
    return {"logits_max_rel_err": err, "tolerance": tol, "sequence": s,
            "checker": "oracle gpt2_forward restatement (all layers, LM head) on the device's weights and this sequence's cached pages"}


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
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + sorted(ALIASES))
    ap.add_argument("--configs", default="all", help="the other BASELINE configs run short beside the headline workload: "
                                                     "all (default, only with the default workload), none, or a comma list")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the timed configuration's output rows")
    ap.add_argument("--no-model", action="store_true", help="skip the whole-model decode step section (SURVEY 8f.2)")
    ap.add_argument("--no-tc-prefill", action="store_true", help="cfg5: skip the opt-in tcgen05 TF32 prefill pass")
    ap.add_argument("--no-prefill", action="store_true", help="skip the prompt-prefill section of the headline line")
    ap.add_argument("--no-clock-hold", action="store_true", help="skip the untimed >=1.5 s continuation (ncu runs)")
    ap.add_argument("--hpg", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--static-pct", type=int, default=0)
    ap.add_argument("--dyn-units", type=int, default=0)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--host-mode", type=int, default=0, help="e2e (PA_TUNE_NO_ZEROCOPY): 0 auto = whole-step staged copies, 1 per-layer staged, 2 zero-copy, 3 whole-step staged")
    ap.add_argument("--timeline", default="", help="write a per-CTA timeline of one decode launch to this file")
    ap.add_argument("--no-fuse", action="store_true", help="separate KV-append kernel instead of the fused decode+append")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wid, w = resolve_workload(args.workload, world if args.impl == "ours" else args.gpus)
    if args.layers > 0:
        w["L"] = args.layers

    import __graft_entry__ as ge
    if args.impl == "reference":
        if rank == 0:
            ge.build_oracle(quiet=True)       # the checkers only: this arm never loads libpaged_attn.so
            reference_arm(args, wid, w)
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
    cx = Ctx(args, pa, lib, dist, rank, local_rank, world)

    head = run_workload(cx, wid, w, primary=True)

    prefill = None
    if wid == "cfg2" and not args.no_prefill:
        try:
            prefill = prefill_headline(cx)
        except SystemExit:
            raise
        except Exception as ex:
            prefill = {"error": repr(ex)}

    others = {}
    if args.configs != "none" and (args.configs != "all" or wid == "cfg2"):
        want = ["cfg1", "cfg3", "cfg4", "cfg5"] if args.configs == "all" else [ALIASES.get(c, c) for c in args.configs.split(",") if c]
        for oid in want:
            if oid == wid or oid not in WORKLOADS:
                continue
            _, ow = resolve_workload(oid, world)
            t0 = time.time()
            try:
                r = run_workload(cx, oid, ow, primary=False)
                r["seconds"] = time.time() - t0
            except SystemExit:
                raise
            except Exception as ex:       # a short side run never costs the headline line
                r = {"workload": ow["name"], "error": repr(ex)}
            others[oid] = r

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = cpu_baseline_sample(wid, w)
        except Exception as ex:     # the baseline is a report, never a reason to lose the line
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                            "sample": f"failed: {ex!r}"}

    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wid, w, world),
            "tokens_per_s": head["tokens_per_s"], "frac_of_measured_peak": head["frac_of_measured_peak"],
            "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "roofline": head["roofline"], "verified": head.get("verified"), "cpu_baseline": cpu_baseline,
            "model": head.get("model"), "prefill": prefill if prefill is not None else head.get("prefill")}
    if others:
        for r in others.values():
            r.pop("clocks", None)
        line["configs"] = others
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
