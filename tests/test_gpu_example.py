"""The plain-C host example (examples/paged_decode_host.c: gcc only, reference call site
paged_infer.c:710-715 unchanged + the batched API) runs on the GPU and matches the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

import __graft_entry__ as ge
import oracle_api as oa

pytestmark = pytest.mark.gpu


def checksum(x):
    x = np.asarray(x, dtype=np.float64).ravel()
    return float((x * ((np.arange(x.size) % 7) + 1)).sum())


def test_plain_c_host_example():
    exe = os.path.join(ge.ROOT, "examples", "paged_decode_host")
    assert os.path.exists(exe), "build() should have produced the example"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    m1 = re.search(r"part1 blocks=(\d+) filled0=(\d+) filled1=(\d+) lru_epoch=(\d+) checksum=([-\d.]+)", r.stdout)
    m2 = re.search(r"part2 ctx=(\d+) pages=(\d+) table0=\[(\d+) (\d+) (\d+)\] checksum=([-\d.]+)", r.stdout)
    assert m1 and m2, r.stdout

    # ---- part 1: the reference call site with main's sliding window
    T, C_, NH, steps, bs = 32, 768, 12, 19, 32
    stream = oa.uniform(((T + steps) * 3 * C_,), -1.0, 1.0, seed=1337).reshape(T + steps, 3 * C_)
    orc = oa.OrcManager(C_, bs, 100, 100)
    want = 0.0
    for step in range(steps):
        window = np.ascontiguousarray(stream[step:step + T][None])
        orc.add_to_cache(window, 1, T, T if step == 0 else 1)
        _, out = orc.attend(0, window, 1, T, NH, step)
        want += checksum(out)
    assert [int(m1.group(i)) for i in (1, 2, 3, 4)] == [2, 32, 18, 19]          # bit-exact integer state
    assert len(orc.table(0)) == 2 and orc.epoch() == 19
    assert abs(float(m1.group(5)) - want) <= 1e-5 * max(1.0, abs(want)) + 1e-3
    orc.close()

    # ---- part 2: batched decode, 8 sequences x 2 layers x 40 steps
    B, L, bs = 8, 2, 16
    data = oa.uniform((40 * L * B * 3 * C_,), -1.0, 1.0, seed=42).reshape(40, L, B, 3 * C_)
    orcs = [oa.OrcManager(C_, bs, 256, B) for _ in range(L)]    # same allocator trace per layer
    want = 0.0
    for step in range(40):
        for layer in range(L):
            qkv = data[step, layer]
            for s in range(B):
                orcs[layer].add_to_cache(qkv[s][None, None, :], 1, 1, 1, prompt=s)
            want += checksum(orcs[layer].decode_batch(list(range(B)), NH, qkv[:, :C_]))
    assert int(m2.group(1)) == 40 and int(m2.group(2)) == 3
    assert [int(m2.group(i)) for i in (3, 4, 5)] == orcs[0].table(0)           # interleaved first-fit: 0, 8, 16
    assert abs(float(m2.group(6)) - want) <= 1e-5 * max(1.0, abs(want)) + 1e-3
    for o in orcs:
        o.close()


def test_plain_c_generate_example_matches_oracle_generation():
    """examples/generate.c (gcc only): synthetic checkpoint + token file in the reference's formats,
    whole prompts prefilled in one step, then sampling with the reference's RNG and sample_mult.
    The CPU oracle fed token by token with the same coins must generate the same token ids."""
    import ctypes as C
    exe = os.path.join(ge.ROOT, "examples", "generate")
    assert os.path.exists(exe), "build() should have produced the example"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr + r.stdout
    got = {}
    for line in r.stdout.splitlines():
        m = re.match(r"sequence (\d+):((?: \d+)+)", line)
        if m:
            got[int(m.group(1))] = [int(t) for t in m.group(2).split()]
    assert sorted(got) == [0, 1, 2, 3] and all(len(v) == 18 for v in got.values()), r.stdout

    pa = ge.load_binding()
    lib = pa.load()
    ol = oa.load_oracle()
    cfg = pa.PaModelConfig()
    pa.check(lib.pa_checkpoint_read_config(b"/tmp/pa_synth_gpt2.bin", C.byref(cfg)), "config")
    n = lib.pa_model_param_count(C.byref(cfg))
    params = np.zeros(n, dtype=np.float32)
    pa.check(lib.pa_checkpoint_read_params(b"/tmp/pa_synth_gpt2.bin", params.ctypes.data, n), "params")
    ids = np.fromfile("/tmp/pa_synth_tokens.bin", dtype=np.int32)
    B, P, total = 4, 32, 50
    L, NH, Cc, V, maxT = cfg.n_layers, cfg.n_heads, cfg.channels, cfg.vocab_size, cfg.max_seq_len
    mgrs = [oa.OrcManager(Cc, 16, 64, B) for _ in range(L)]
    arr = (C.c_void_p * L)(*[m.m for m in mgrs])

    def step(seqs, toks, pos):
        seq = np.ascontiguousarray(seqs, dtype=np.int32); tok = np.ascontiguousarray(toks, dtype=np.int32)
        ps = np.ascontiguousarray(pos, dtype=np.int32)
        logits = np.zeros((len(seq), V), dtype=np.float32)
        assert ol.orc_model_decode_step(arr, L, NH, Cc, V, maxT, oa.fptr(params), oa.iptr(seq), oa.iptr(tok), oa.iptr(ps),
                                        len(seq), oa.fptr(logits)) == 0
        return logits

    def sample(logits, coins):
        probs = np.zeros_like(logits)
        ol.orc_softmax_forward(oa.fptr(probs), oa.fptr(logits), len(logits), 1, V)
        return [ol.orc_sample_mult(oa.fptr(np.ascontiguousarray(probs[i])), V, float(coins[i])) for i in range(len(logits))]

    state = C.c_ulonglong(1337)
    draw = lambda: [ol.orc_random_f32(C.byref(state)) for _ in range(B)]      # noqa: E731
    try:
        coins = draw()
        for t in range(P):
            logits = step(list(range(B)), [ids[b * P + t] for b in range(B)], [t] * B)
        nxt = sample(logits, coins)
        want = {b: [] for b in range(B)}
        for t in range(P, total):
            for b in range(B):
                want[b].append(nxt[b])
            coins = draw()
            if t + 1 < total:
                nxt = sample(step(list(range(B)), nxt, [t] * B), coins)
        assert got == want
    finally:
        for m in mgrs:
            m.close()


def _multi_tokens(stdout):
    got = {}
    for line in stdout.splitlines():
        m = re.match(r"rank (\d+) sequence (\d+):((?: \d+)+)", line)
        if m:
            got[(int(m.group(1)), int(m.group(2)))] = [int(t) for t in m.group(3).split()]
    return got


def test_plain_c_multi_gpu_example_rank0_equals_single_gpu_generation():
    """examples/generate_multi.c (gcc only): ONE process drives a pa_group; the sampled tokens travel through
    pa_group_model_step (forward enqueued, all-gather on the handle's stream, one wait).  With one GPU the
    group's rank 0 must generate exactly what examples/generate.c generates (already pinned to the oracle)."""
    single = subprocess.run([os.path.join(ge.ROOT, "examples", "generate")], capture_output=True, text=True, timeout=180)
    assert single.returncode == 0, single.stderr
    want = {}
    for line in single.stdout.splitlines():
        m = re.match(r"sequence (\d+):((?: \d+)+)", line)
        if m:
            want[(0, int(m.group(1)))] = [int(t) for t in m.group(2).split()]
    exe = os.path.join(ge.ROOT, "examples", "generate_multi")
    assert os.path.exists(exe), "build() should have produced the example"
    r = subprocess.run([exe, "1"], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr + r.stdout
    assert _multi_tokens(r.stdout) == want and len(want) == 4


def test_plain_c_multi_gpu_example_sharded_over_all_gpus():
    """With n GPUs: rank r's lines (printed from the NCCL-gathered buffer) equal what one GPU playing rank r
    alone generates.  Needs >= 2 GPUs (`gpurun --gpus 2`); the driver's 1-GPU test box skips it."""
    pa = ge.load_binding()
    n = min(pa.load().pa_device_count(), 8)
    if n < 2:
        pytest.skip("one GPU visible")
    exe = os.path.join(ge.ROOT, "examples", "generate_multi")
    r = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    got = _multi_tokens(r.stdout)
    assert len(got) == 4 * n
    assert f"group: {n} GPU(s)" in r.stdout and "NCCL 2" in r.stdout
    for rank in range(n):
        alone = subprocess.run([exe, "1", str(rank)], capture_output=True, text=True, timeout=180)
        assert alone.returncode == 0, alone.stderr
        want = _multi_tokens(alone.stdout)
        assert {k: v for k, v in got.items() if k[0] == rank} == want, rank
