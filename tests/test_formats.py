"""the reference's on-disk formats (checkpoint, token stream, tokenizer), host-only.
Files written by this library are read by the REFERENCE's own readers (compiled from
/root/reference into oracle/_ref) and must come back bit-identical; and the other way round a file
in the documented layout is read by this library.  CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import __graft_entry__ as ge
import oracle_api as oa

pa = ge.load_binding()
GEOM = (32, 100, 100)


def need_ref():
    if not oa.have_ref(*GEOM):
        pytest.skip("compiled reference not available")
    return oa.load_ref(*GEOM)


def test_checkpoint_roundtrip_through_the_reference_reader(tmp_path):
    lib = pa.load()
    cfg = pa.PaModelConfig(48, 97, 2, 2, 16)                  # maxT, V, L, NH, C
    n = lib.pa_model_param_count(C.byref(cfg))
    params = oa.normal((n,), seed=21)
    path = str(tmp_path / "gpt2_tiny.bin").encode()
    pa.check(lib.pa_checkpoint_write(path, C.byref(cfg), params.ctypes.data), "write")
    assert os.path.getsize(path) == 256 * 4 + n * 4
    # our reader
    cfg2 = pa.PaModelConfig()
    pa.check(lib.pa_checkpoint_read_config(path, C.byref(cfg2)), "read config")
    assert [cfg2.max_seq_len, cfg2.vocab_size, cfg2.n_layers, cfg2.n_heads, cfg2.channels] == [48, 97, 2, 2, 16]
    back = np.zeros(n, dtype=np.float32)
    pa.check(lib.pa_checkpoint_read_params(path, back.ctypes.data, n), "read params")
    assert np.array_equal(back.view(np.uint32), params.view(np.uint32))
    # the reference's gpt2_build_from_checkpoint (paged_infer.c:436-502)
    rl = need_ref()
    c5 = np.zeros(5, dtype=np.int32)
    ref_params = np.zeros(n, dtype=np.float32)
    rl.ref_silence(1)
    got_n = rl.ref_checkpoint_load(path, oa.iptr(c5), oa.fptr(ref_params), n)
    rl.ref_silence(0)
    assert got_n == n and list(c5) == [48, 97, 2, 2, 16]
    assert np.array_equal(ref_params.view(np.uint32), params.view(np.uint32))


def test_checkpoint_rejects_bad_files(tmp_path):
    lib = pa.load()
    cfg = pa.PaModelConfig()
    assert lib.pa_checkpoint_read_config(str(tmp_path / "missing.bin").encode(), C.byref(cfg)) == pa.PA_ERR_INVALID
    assert "Error opening model file" in pa.last_error()
    bad = tmp_path / "bad.bin"
    hdr = np.zeros(256, dtype=np.int32); hdr[0] = 123
    bad.write_bytes(hdr.tobytes())
    assert lib.pa_checkpoint_read_config(str(bad).encode(), C.byref(cfg)) == pa.PA_ERR_INVALID
    assert "Bad magic" in pa.last_error()
    hdr[0] = 20240326; hdr[1] = 3
    bad.write_bytes(hdr.tobytes())
    assert lib.pa_checkpoint_read_config(str(bad).encode(), C.byref(cfg)) == pa.PA_ERR_INVALID
    assert "Bad version" in pa.last_error()


def _bf16_rne(x):
    """torch's .to(torch.bfloat16) on fp32: round to nearest even on the upper 16 bits."""
    b = x.view(np.uint32).astype(np.uint64)
    return ((b + 0x7fff + ((b >> 16) & 1)) >> 16).astype(np.uint16)


def test_checkpoint_version_2_bf16_layout_of_the_reference_writer(tmp_path):
    """Version 2 of the checkpoint (train_gpt2.py:298-320 write_model(dtype="bfloat16")): header version 2, the ten
    weight/bias tensors as bf16 in the order of write_tensors_bf16 (train_gpt2.py:266-297), then the six layernorm
    tensors in fp32.  The writer's bytes equal a numpy restatement of that function; the reader widens bf16 exactly
    and returns the parameters in MODEL order."""
    lib = pa.load()
    maxT, V, L, NH, Cc = 24, 53, 3, 2, 16
    cfg = pa.PaModelConfig(maxT, V, L, NH, Cc)
    n = lib.pa_model_param_count(C.byref(cfg))
    params = oa.normal((n,), seed=33)
    sizes = [V * Cc, maxT * Cc, L * Cc, L * Cc, L * 3 * Cc * Cc, L * 3 * Cc, L * Cc * Cc, L * Cc, L * Cc, L * Cc,
             L * 4 * Cc * Cc, L * 4 * Cc, L * 4 * Cc * Cc, L * Cc, Cc, Cc]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    tens = [params[offs[i]:offs[i + 1]] for i in range(16)]
    path = str(tmp_path / "gpt2_tiny_bf16.bin").encode()
    pa.check(lib.pa_checkpoint_write_bf16(path, C.byref(cfg), params.ctypes.data), "write bf16")
    hdr = np.zeros(256, dtype=np.int32)
    hdr[:7] = [20240326, 2, maxT, V, L, NH, Cc]
    want = hdr.tobytes()
    for t in (0, 1, 4, 5, 6, 7, 10, 11, 12, 13):                # wte wpe qkvw qkvb attprojw attprojb fcw fcb fcprojw fcprojb
        want += _bf16_rne(tens[t]).tobytes()
    for t in (2, 3, 8, 9, 14, 15):                              # ln1w ln1b ln2w ln2b lnfw lnfb stay fp32
        want += tens[t].tobytes()
    assert open(path, "rb").read() == want
    cfg2 = pa.PaModelConfig()
    pa.check(lib.pa_checkpoint_read_config(path, C.byref(cfg2)), "read config")
    assert [cfg2.max_seq_len, cfg2.vocab_size, cfg2.n_layers, cfg2.n_heads, cfg2.channels] == [maxT, V, L, NH, Cc]
    back = np.full(n, np.nan, dtype=np.float32)
    pa.check(lib.pa_checkpoint_read_params(path, back.ctypes.data, n), "read params")
    for t in range(16):
        got = back[offs[t]:offs[t + 1]]
        if t in (2, 3, 8, 9, 14, 15):
            assert np.array_equal(got.view(np.uint32), tens[t].view(np.uint32))
        else:
            assert np.array_equal(got.view(np.uint32), _bf16_rne(tens[t]).astype(np.uint32) << 16)
    # a truncated file is refused
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(want[:len(want) // 2])
    assert lib.pa_checkpoint_read_params(str(trunc).encode(), back.ctypes.data, n) == pa.PA_ERR_INVALID


def test_dataloader_matches_reference_including_wraparound(tmp_path):
    lib = pa.load()
    rl = need_ref()
    B, T = 3, 7
    ids = (np.arange(100, dtype=np.int32) * 37 % 50257).astype(np.int32)       # 100 ids: wraps after 4 batches
    path = str(tmp_path / "tokens.bin").encode()
    pa.check(lib.pa_tokens_write(path, oa.iptr(ids), ids.size), "tokens write")
    d = C.c_void_p()
    pa.check(lib.pa_dataloader_open(path, B, T, C.byref(d)), "open")
    rd = rl.ref_dataloader_open(path, B, T)
    try:
        assert lib.pa_dataloader_num_batches(d) == rl.ref_dataloader_num_batches(rd) == 100 * 4 // (B * T * 4)
        for it in range(11):
            inp, tgt = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
            pa.check(lib.pa_dataloader_next_batch(d, C.byref(inp), C.byref(tgt)), "next")
            ours = np.ctypeslib.as_array(inp, (B * T + 1,)).copy()
            ref = np.zeros(B * T + 1, dtype=np.int32)
            rl.ref_dataloader_next(rd, oa.iptr(ref))
            assert np.array_equal(ours, ref), it
            assert np.array_equal(np.ctypeslib.as_array(tgt, (B * T,)), ref[1:])
            if it == 5:
                lib.pa_dataloader_reset(d); rl.ref_dataloader_reset(rd)
    finally:
        lib.pa_dataloader_close(d); rl.ref_dataloader_free(rd)
    # too small a file: the reference exit(1)s (paged_infer.c:784-787); the library returns an error
    small = str(tmp_path / "small.bin").encode()
    pa.check(lib.pa_tokens_write(small, oa.iptr(ids), 5), "tokens write")
    assert lib.pa_dataloader_open(small, B, T, C.byref(d)) == pa.PA_ERR_INVALID
    assert "too small" in pa.last_error()


def test_tokenizer_matches_reference(tmp_path):
    lib = pa.load()
    rl = need_ref()
    pieces = [b"a", b" the", b"\xe2\x82\xac", b"\n", b"token with spaces", bytes(range(1, 100))]
    arr = (C.c_char_p * len(pieces))(*pieces)
    lens = (C.c_ubyte * len(pieces))(*[len(p) for p in pieces])
    path = str(tmp_path / "tok.bin").encode()
    pa.check(lib.pa_tokenizer_write(path, arr, lens, len(pieces)), "tokenizer write")
    t = C.c_void_p()
    pa.check(lib.pa_tokenizer_open(path, C.byref(t)), "open")
    rt = rl.ref_tokenizer_open(path)
    try:
        assert lib.pa_tokenizer_vocab_size(t) == rl.ref_tokenizer_vocab(rt) == len(pieces)
        for i, p in enumerate(pieces):
            assert lib.pa_tokenizer_decode(t, i) == rl.ref_tokenizer_decode(rt, i) == p
        assert lib.pa_tokenizer_decode(t, len(pieces)) is None
    finally:
        lib.pa_tokenizer_close(t)
    assert lib.pa_tokenizer_open(str(tmp_path / "none.bin").encode(), C.byref(t)) == pa.PA_ERR_INVALID
