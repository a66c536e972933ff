"""ctypes adapters for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``RefManager``  -- the reference's own compiled code, oracle/_ref/libref_*.so, built by
  oracle/build_ref.sh from /root/reference (block_manager.c, paged_infer.c).
* ``OrcManager``  -- the CPU restatement oracle/paged_oracle.c (run-time geometry).

Both expose the same small interface so one trace driver can run on either (and on the
product's block manager, see tests/pa_api.py).
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)
c_float_pp = C.POINTER(c_float_p)


def fptr(a):
    return a.ctypes.data_as(c_float_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


# --------------------------------------------------------------------------- reference
def ref_path(bs, mb, mp, flavor="strict"):
    return os.path.join(REF_DIR, f"libref_bs{bs}_mb{mb}_mp{mp}_{flavor}.so")


def have_ref(bs, mb, mp, flavor="strict"):
    return os.path.exists(ref_path(bs, mb, mp, flavor))


_ref_cache = {}


def load_ref(bs, mb, mp, flavor="strict"):
    key = (bs, mb, mp, flavor)
    if key in _ref_cache:
        return _ref_cache[key]
    lib = C.CDLL(ref_path(bs, mb, mp, flavor))
    vp = C.c_void_p
    lib.ref_create.restype = vp
    lib.ref_create.argtypes = [C.c_int]
    lib.ref_destroy.argtypes = [vp]
    for name in ("ref_request_block", "ref_get_current_block", "ref_block_count"):
        getattr(lib, name).argtypes = [vp, C.c_int]
        getattr(lib, name).restype = C.c_int
    lib.ref_free_blocks_for_prompt.argtypes = [vp, C.c_int]
    lib.ref_find_lru.argtypes = [vp]
    lib.ref_find_lru.restype = C.c_int
    lib.ref_page_out_lru.argtypes = [vp]
    lib.ref_get_next_block_id.argtypes = [vp, C.c_int, C.c_int]
    lib.ref_get_next_block_id.restype = C.c_int
    lib.ref_lru_epoch.argtypes = [vp]
    lib.ref_lru_epoch.restype = C.c_int
    lib.ref_block_table.argtypes = [vp, C.c_int, c_int_p, C.c_int]
    lib.ref_block_table.restype = C.c_int
    lib.ref_block_info.argtypes = [vp, C.c_int, c_int_p, c_int_p, c_int_p]
    lib.ref_block_ptrs.argtypes = [vp, C.c_int, c_float_pp, c_float_pp]
    lib.ref_touch.argtypes = [vp, C.c_int]
    lib.ref_set_filled.argtypes = [vp, C.c_int, C.c_int]
    lib.ref_add_to_cache.argtypes = [vp, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_attention_paged.argtypes = [c_float_p, c_float_p, c_float_p, c_float_p, c_float_pp, c_float_pp,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_attend_prompt.argtypes = [vp, C.c_int, c_float_p, c_float_p, c_float_p, c_float_p,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_attend_prompt.restype = C.c_int
    lib.ref_time_attend_prompt.argtypes = [vp, C.c_int, c_float_p, c_float_p,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_time_attend_prompt.restype = C.c_double
    lib.ref_checkpoint_load.argtypes = [C.c_char_p, c_int_p, c_float_p, C.c_long]
    lib.ref_checkpoint_load.restype = C.c_long
    lib.ref_dataloader_open.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.ref_dataloader_open.restype = vp
    lib.ref_dataloader_num_batches.argtypes = [vp]
    lib.ref_dataloader_next.argtypes = [vp, c_int_p]
    lib.ref_dataloader_reset.argtypes = [vp]
    lib.ref_dataloader_free.argtypes = [vp]
    lib.ref_tokenizer_open.argtypes = [C.c_char_p]
    lib.ref_tokenizer_open.restype = vp
    lib.ref_tokenizer_vocab.argtypes = [vp]
    lib.ref_tokenizer_decode.argtypes = [vp, C.c_uint]
    lib.ref_tokenizer_decode.restype = C.c_char_p
    lib.ref_encoder_forward.argtypes = [c_float_p, c_int_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int]
    lib.ref_layernorm_forward.argtypes = [c_float_p] * 6 + [C.c_int] * 3
    lib.ref_gelu_forward.argtypes = [c_float_p, c_float_p, C.c_int]
    lib.ref_residual_forward.argtypes = [c_float_p, c_float_p, c_float_p, C.c_int]
    lib.ref_softmax_forward.argtypes = [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int]
    lib.ref_sample_mult.argtypes = [c_float_p, C.c_int, C.c_float]
    lib.ref_sample_mult.restype = C.c_int
    for name in ("ref_matmul_forward", "ref_matmul_cached"):
        getattr(lib, name).argtypes = [c_float_p] * 4 + [C.c_int] * 4
    lib.ref_geometry.argtypes = [c_int_p, c_int_p, c_int_p]
    lib.ref_omp_threads.restype = C.c_int
    lib.ref_silence.argtypes = [C.c_int]
    lib.ref_random_u32.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.ref_random_u32.restype = C.c_uint
    lib.ref_random_f32.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.ref_random_f32.restype = C.c_float
    _ref_cache[key] = lib
    return lib


class _Silenced:
    def __init__(self, lib):
        self.lib = lib

    def __enter__(self):
        self.lib.ref_silence(1)

    def __exit__(self, *a):
        self.lib.ref_silence(0)


class RefManager:
    """The reference's BlockManager (block_manager.c:17-23) driven through ref_wrap.c."""
    kind = "reference"

    def __init__(self, channels, bs, mb, mp, flavor="strict"):
        self.lib = load_ref(bs, mb, mp, flavor)
        self.C, self.bs, self.max_blocks, self.max_prompts = channels, bs, mb, mp
        self.m = self.lib.ref_create(channels)

    def close(self):
        if self.m:
            with _Silenced(self.lib):
                self.lib.ref_destroy(self.m)
            self.m = None

    def silenced(self):
        return _Silenced(self.lib)

    def request_block(self, p):
        with _Silenced(self.lib):
            return self.lib.ref_request_block(self.m, p)

    def get_current_block(self, p):
        with _Silenced(self.lib):
            return self.lib.ref_get_current_block(self.m, p)

    def free_blocks_for_prompt(self, p):
        with _Silenced(self.lib):
            self.lib.ref_free_blocks_for_prompt(self.m, p)

    def find_lru(self):
        return self.lib.ref_find_lru(self.m)

    def page_out_lru(self):
        with _Silenced(self.lib):
            self.lib.ref_page_out_lru(self.m)

    def get_next_block_id(self, p, bid):
        return self.lib.ref_get_next_block_id(self.m, p, bid)

    def touch(self, idx):
        self.lib.ref_touch(self.m, idx)

    def set_filled(self, idx, f):
        self.lib.ref_set_filled(self.m, idx, f)

    def choose_page(self, p):
        """Page choice of add_to_cache (paged_infer.c:518-529) for prompt p, via the
        reference's own get_current_block / request_block."""
        cur = self.get_current_block(p)
        if cur >= 0:
            if self.block_info(cur)[0] >= self.bs:
                cur = self.request_block(p)
            else:
                self.touch(cur)
        else:
            cur = self.request_block(p)
        return cur

    def add_to_cache(self, qkv, B, T, n_tail):
        """The real add_to_cache (prompt 0 only, paged_infer.c:515)."""
        qkv = np.ascontiguousarray(qkv, dtype=np.float32)
        with _Silenced(self.lib):
            self.lib.ref_add_to_cache(self.m, fptr(qkv), B, T, self.C, n_tail)

    def epoch(self):
        return self.lib.ref_lru_epoch(self.m)

    def table(self, p):
        n = self.lib.ref_block_count(self.m, p)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.ref_block_table(self.m, p, iptr(out), n)
        return out[:n].tolist()

    def block_info(self, idx):
        f, p, l = C.c_int(), C.c_int(), C.c_int()
        self.lib.ref_block_info(self.m, idx, C.byref(f), C.byref(p), C.byref(l))
        return f.value, p.value, l.value

    def page_arrays(self, idx):
        k, v = c_float_p(), c_float_p()
        self.lib.ref_block_ptrs(self.m, idx, C.byref(k), C.byref(v))
        shape = (self.bs, self.C)
        return np.ctypeslib.as_array(k, shape), np.ctypeslib.as_array(v, shape)

    def attend(self, prompt, inp, B, T, NH, offset, want_scratch=False):
        inp = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.zeros((B, T, self.C), dtype=np.float32)
        if want_scratch:
            pre = np.zeros((B, NH, T, T), dtype=np.float32)
            att = np.zeros((B, NH, T, T), dtype=np.float32)
            rc = self.lib.ref_attend_prompt(self.m, prompt, fptr(out), fptr(pre), fptr(att), fptr(inp),
                                            B, T, self.C, NH, offset)
            return rc, out, pre, att
        rc = self.lib.ref_attend_prompt(self.m, prompt, fptr(out), None, None, fptr(inp), B, T, self.C, NH, offset)
        return rc, out


# --------------------------------------------------------------------------- restatement
_orc_cache = {}


def oracle_path(flavor="strict"):
    return os.path.join(ORACLE_DIR, f"liboracle_{flavor}.so")


def load_oracle(flavor="strict"):
    if flavor in _orc_cache:
        return _orc_cache[flavor]
    lib = C.CDLL(oracle_path(flavor))
    vp = C.c_void_p
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [C.c_int] * 5
    lib.orc_destroy.argtypes = [vp]
    for name in ("orc_request_block", "orc_get_current_block", "orc_block_count", "orc_choose_page",
                 "orc_context_len"):
        getattr(lib, name).argtypes = [vp, C.c_int]
        getattr(lib, name).restype = C.c_int
    lib.orc_free_blocks_for_prompt.argtypes = [vp, C.c_int]
    lib.orc_find_lru.argtypes = [vp]
    lib.orc_find_lru.restype = C.c_int
    lib.orc_page_out_lru.argtypes = [vp]
    lib.orc_get_next_block_id.argtypes = [vp, C.c_int, C.c_int]
    lib.orc_get_next_block_id.restype = C.c_int
    lib.orc_lru_epoch.argtypes = [vp]
    lib.orc_lru_epoch.restype = C.c_int
    lib.orc_block_table.argtypes = [vp, C.c_int, c_int_p, C.c_int]
    lib.orc_block_table.restype = C.c_int
    lib.orc_block_info.argtypes = [vp, C.c_int, c_int_p, c_int_p, c_int_p]
    lib.orc_block_ptrs.argtypes = [vp, C.c_int, c_float_pp, c_float_pp]
    lib.orc_touch.argtypes = [vp, C.c_int]
    lib.orc_set_filled.argtypes = [vp, C.c_int, C.c_int]
    lib.orc_slot.argtypes = [vp, C.c_int, C.c_int]
    lib.orc_slot.restype = C.c_int
    lib.orc_add_to_cache.argtypes = [vp, C.c_int, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_add_to_cache.restype = C.c_int
    lib.orc_attention_paged.argtypes = [c_float_p, c_float_p, c_float_p, c_float_p, c_float_pp, c_float_pp,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_decode_batch.argtypes = [vp, c_int_p, c_int_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, C.c_int]
    lib.orc_decode_batch.restype = C.c_int
    lib.orc_decode_batch_f64.argtypes = [vp, c_int_p, c_int_p, C.c_int, C.c_int, c_float_p, C.c_int,
                                         C.POINTER(C.c_double), C.c_int]
    lib.orc_attend_rows.argtypes = [vp, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, C.c_int, C.c_int,
                                    c_float_p, C.c_int, c_float_p, C.c_int]
    lib.orc_time_decode_batch.argtypes = [vp, c_int_p, C.c_int, C.c_int, c_float_p, C.c_int, c_float_p, C.c_int, C.c_int]
    lib.orc_time_decode_batch.restype = C.c_double
    lib.orc_encoder_forward.argtypes = [c_float_p, c_int_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int]
    lib.orc_layernorm_forward.argtypes = [c_float_p] * 6 + [C.c_int] * 3
    lib.orc_gelu_forward.argtypes = [c_float_p, c_float_p, C.c_int]
    lib.orc_residual_forward.argtypes = [c_float_p, c_float_p, c_float_p, C.c_int]
    lib.orc_softmax_forward.argtypes = [c_float_p, c_float_p, C.c_int, C.c_int, C.c_int]
    lib.orc_sample_mult.argtypes = [c_float_p, C.c_int, C.c_float]
    lib.orc_sample_mult.restype = C.c_int
    lib.orc_model_decode_step.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p,
                                          c_int_p, c_int_p, c_int_p, C.c_int, c_float_p]
    lib.orc_model_decode_step.restype = C.c_int
    for name in ("orc_matmul_forward", "orc_matmul_cached"):
        getattr(lib, name).argtypes = [c_float_p] * 4 + [C.c_int] * 4
    lib.orc_fill_normal.argtypes = [c_float_p, C.c_size_t, C.c_ulonglong]
    lib.orc_fill_uniform.argtypes = [c_float_p, C.c_size_t, C.c_float, C.c_float, C.c_ulonglong]
    lib.orc_random_u32.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.orc_random_u32.restype = C.c_uint
    lib.orc_random_f32.argtypes = [C.POINTER(C.c_ulonglong)]
    lib.orc_random_f32.restype = C.c_float
    lib.orc_omp_threads.restype = C.c_int
    _orc_cache[flavor] = lib
    return lib


def normal(shape, seed, flavor="strict"):
    """N(0,1) fp32 from the reference RNG (xorshift64*, paged_infer.c:826-835) + Box-Muller."""
    a = np.empty(shape, dtype=np.float32)
    load_oracle(flavor).orc_fill_normal(fptr(a), a.size, seed)
    return a


def uniform(shape, lo, hi, seed, flavor="strict"):
    a = np.empty(shape, dtype=np.float32)
    load_oracle(flavor).orc_fill_uniform(fptr(a), a.size, lo, hi, seed)
    return a


class OrcManager:
    """oracle/paged_oracle.c manager (run-time geometry)."""
    kind = "oracle"

    def __init__(self, channels, bs, mb, mp, flavor="strict", alloc_data=True):
        self.lib = load_oracle(flavor)
        self.C, self.bs, self.max_blocks, self.max_prompts = channels, bs, mb, mp
        self.m = self.lib.orc_create(channels, bs, mb, mp, 1 if alloc_data else 0)

    def close(self):
        if self.m:
            self.lib.orc_destroy(self.m)
            self.m = None

    def request_block(self, p):
        return self.lib.orc_request_block(self.m, p)

    def get_current_block(self, p):
        return self.lib.orc_get_current_block(self.m, p)

    def free_blocks_for_prompt(self, p):
        self.lib.orc_free_blocks_for_prompt(self.m, p)

    def find_lru(self):
        return self.lib.orc_find_lru(self.m)

    def page_out_lru(self):
        self.lib.orc_page_out_lru(self.m)

    def get_next_block_id(self, p, bid):
        return self.lib.orc_get_next_block_id(self.m, p, bid)

    def touch(self, idx):
        self.lib.orc_touch(self.m, idx)

    def set_filled(self, idx, f):
        self.lib.orc_set_filled(self.m, idx, f)

    def choose_page(self, p):
        return self.lib.orc_choose_page(self.m, p)

    def add_to_cache(self, qkv, B, T, n_tail, prompt=0):
        qkv = np.ascontiguousarray(qkv, dtype=np.float32)
        return self.lib.orc_add_to_cache(self.m, prompt, fptr(qkv), B, T, self.C, n_tail)

    def epoch(self):
        return self.lib.orc_lru_epoch(self.m)

    def table(self, p):
        n = self.lib.orc_block_count(self.m, p)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.orc_block_table(self.m, p, iptr(out), n)
        return out[:n].tolist()

    def block_info(self, idx):
        f, p, l = C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_block_info(self.m, idx, C.byref(f), C.byref(p), C.byref(l))
        return f.value, p.value, l.value

    def context_len(self, p):
        return self.lib.orc_context_len(self.m, p)

    def slot(self, p, pos):
        return self.lib.orc_slot(self.m, p, pos)

    def page_arrays(self, idx):
        k, v = c_float_p(), c_float_p()
        self.lib.orc_block_ptrs(self.m, idx, C.byref(k), C.byref(v))
        shape = (self.bs, self.C)
        return np.ctypeslib.as_array(k, shape), np.ctypeslib.as_array(v, shape)

    def _page_ptr_arrays(self, p):
        tbl = self.table(p)
        ks = (c_float_p * len(tbl))()
        vs = (c_float_p * len(tbl))()
        for i, idx in enumerate(tbl):
            k, v = c_float_p(), c_float_p()
            self.lib.orc_block_ptrs(self.m, idx, C.byref(k), C.byref(v))
            ks[i], vs[i] = k, v
        return ks, vs

    def attend(self, prompt, inp, B, T, NH, offset, want_scratch=False):
        inp = np.ascontiguousarray(inp, dtype=np.float32)
        out = np.zeros((B, T, self.C), dtype=np.float32)
        pre = np.zeros((B, NH, T, T), dtype=np.float32)
        att = np.zeros((B, NH, T, T), dtype=np.float32)
        ks, vs = self._page_ptr_arrays(prompt)
        if len(ks) == 0:
            return (-1, out, pre, att) if want_scratch else (-1, out)
        self.lib.orc_attention_paged(fptr(out), fptr(pre), fptr(att), fptr(inp), ks, vs,
                                     B, T, self.C, NH, offset, self.bs)
        return (len(ks), out, pre, att) if want_scratch else (len(ks), out)

    def decode_batch(self, seq_ids, NH, q, kv_start=None):
        """q: (nseq, C) -> out (nseq, C); row i = last-row attention of prompt seq_ids[i]."""
        seq = np.ascontiguousarray(seq_ids, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((len(seq), self.C), dtype=np.float32)
        ks = None if kv_start is None else np.ascontiguousarray(kv_start, dtype=np.int32)
        self.lib.orc_decode_batch(self.m, iptr(seq), None if ks is None else iptr(ks), len(seq), NH,
                                  fptr(q), q.shape[1], fptr(out), self.C)
        return out

    def decode_batch_f64(self, seq_ids, NH, q, kv_start=None):
        seq = np.ascontiguousarray(seq_ids, dtype=np.int32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((len(seq), self.C), dtype=np.float64)
        ks = None if kv_start is None else np.ascontiguousarray(kv_start, dtype=np.int32)
        self.lib.orc_decode_batch_f64(self.m, iptr(seq), None if ks is None else iptr(ks), len(seq), NH,
                                      fptr(q), q.shape[1], out.ctypes.data_as(C.POINTER(C.c_double)), self.C)
        return out

    def attend_rows(self, seq_ids, kv_start, base_len, n_q, NH, q):
        seq = np.ascontiguousarray(seq_ids, dtype=np.int32)
        ks = np.ascontiguousarray(kv_start, dtype=np.int32)
        bl = np.ascontiguousarray(base_len, dtype=np.int32)
        nq = np.ascontiguousarray(n_q, dtype=np.int32)
        row0 = np.concatenate([[0], np.cumsum(nq)[:-1]]).astype(np.int32)
        q = np.ascontiguousarray(q, dtype=np.float32)
        out = np.zeros((int(nq.sum()), self.C), dtype=np.float32)
        self.lib.orc_attend_rows(self.m, iptr(seq), iptr(ks), iptr(bl), iptr(nq), iptr(row0), len(seq), NH,
                                 fptr(q), q.shape[1], fptr(out), self.C)
        return out
