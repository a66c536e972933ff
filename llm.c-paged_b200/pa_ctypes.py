"""ctypes binding of include/paged_attn.h -- the harness side of the C ABI.

Used by tests/, bench.py and __graft_entry__.py only; the product is libpaged_attn.so and its
plain-C host interface.  Nothing here computes: every call lands in the shared library, and the
loader raises if the library is missing (there is no Python or CPU fallback).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# (PA_LIB_PATH: a developer hook -- A/B an alternative build of the same ABI on one box)
LIB_PATH = os.environ.get("PA_LIB_PATH") or os.path.join(HERE, "libpaged_attn.so")

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)
c_float_pp = C.POINTER(c_float_p)
vp = C.c_void_p

PA_HOST_ONLY = -1
PA_OK = 0
PA_ERR_INVALID, PA_ERR_NOMEM, PA_ERR_CUDA, PA_ERR_NO_DEVICE, PA_ERR_NO_BLOCKS, PA_ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
PA_TUNE_DECODE_PATH, PA_TUNE_HEADS_PER_TILE, PA_TUNE_STAGES, PA_TUNE_GRID, PA_TUNE_COUNT_LAUNCHES = 0, 1, 2, 3, 4
PA_TUNE_STATIC_PCT, PA_TUNE_DYN_UNITS, PA_TUNE_DEBUG_TIMELINE, PA_TUNE_NO_PDL, PA_TUNE_NO_ZEROCOPY = 5, 6, 7, 8, 9
PA_TUNE_LAST_HPG, PA_TUNE_LAST_STAGES, PA_TUNE_LAST_GRID, PA_TUNE_PREFILL_PATH = 10, 11, 12, 13
PA_TUNE_TC_WARPGROUPS = 14
PA_TUNE_TC_KEY_TILE = 15
PA_TUNE_GEMM_PATH = 16
PA_TUNE_GEMM_SPLIT_K = 17
PA_TUNE_MODEL_PATH = 18


class KVBlock(C.Structure):
    _fields_ = [("keys", vp), ("values", vp), ("filled", C.c_int), ("prompt_id", C.c_int), ("lru_counter", C.c_int)]


class BlockManager(C.Structure):
    _fields_ = [("C", C.c_int), ("blocks", C.POINTER(KVBlock)), ("prompt_block_list", C.POINTER(c_int_p)),
                ("prompt_block_count", c_int_p), ("lru_epoch", C.c_int),
                ("block_size", C.c_int), ("max_blocks", C.c_int), ("max_prompts", C.c_int),
                ("table_stride", C.c_int), ("block_table", c_int_p), ("pa", vp),
                ("refcount", c_int_p), ("prefix_cache", vp), ("pinned", C.POINTER(C.c_ubyte))]


class PaConfig(C.Structure):
    _fields_ = [("block_size", C.c_int), ("max_blocks", C.c_int), ("max_seqs", C.c_int),
                ("max_blocks_per_seq", C.c_int), ("n_layers", C.c_int), ("n_heads", C.c_int),
                ("head_dim", C.c_int), ("device", C.c_int), ("max_batch_tokens", C.c_int)]


BM_p = C.POINTER(BlockManager)
KV_p = C.POINTER(KVBlock)

_lib = None


class PaModelConfig(C.Structure):
    _fields_ = [("max_seq_len", C.c_int), ("vocab_size", C.c_int), ("n_layers", C.c_int), ("n_heads", C.c_int),
                ("channels", C.c_int)]


class PagedAttnError(RuntimeError):
    pass


def load():
    """Load libpaged_attn.so (raises if it has not been built -- no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PagedAttnError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(LIB_PATH)
    sig = {
        # compat
        "create_block_manager": (BM_p, [C.c_int]),
        "destroy_block_manager": (None, [BM_p]),
        "print_state": (None, [BM_p, C.c_int]),
        "get_next_block_id": (C.c_int, [BM_p, C.c_int, C.c_int]),
        "get_current_block": (KV_p, [BM_p, C.c_int]),
        "free_blocks_for_prompt": (None, [BM_p, C.c_int]),
        "find_least_recently_used_block": (C.c_int, [BM_p]),
        "page_out_lru_block": (None, [BM_p]),
        "request_block": (KV_p, [BM_p, C.c_int]),
        "collect_kv_blocks": (C.POINTER(c_float_pp), [BM_p, C.c_int, c_int_p]),
        "add_to_cache": (None, [BM_p, vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "attention_paged": (None, [vp, vp, vp, vp, c_float_pp, c_float_pp] + [C.c_int] * 5),
        "pa_set_default_geometry": (None, [C.c_int] * 3),
        "pa_default_block_size": (C.c_int, []),
        # extended
        "pa_create": (C.c_int, [C.POINTER(PaConfig), C.POINTER(vp)]),
        "pa_destroy": (None, [vp]),
        "pa_manager": (BM_p, [vp]),
        "pa_last_error": (C.c_char_p, []),
        "pa_version": (C.c_char_p, []),
        "pa_step_begin": (C.c_int, [vp, c_int_p, c_int_p, C.c_int]),
        "pa_step_set_kv_start": (C.c_int, [vp, c_int_p]),
        "pa_step_begin_readonly": (C.c_int, [vp, c_int_p, C.c_int]),
        "pa_step_slot_mapping": (c_int_p, [vp, c_int_p]),
        "pa_step_context_lens": (c_int_p, [vp, c_int_p]),
        "pa_step_block_table": (c_int_p, [vp, c_int_p, c_int_p]),
        "pa_step_upload": (C.c_int, [vp, vp]),
        "pa_step_validate": (C.c_int, [vp]),
        "pa_append": (C.c_int, [vp, C.c_int, vp, vp, C.c_int, vp]),
        "pa_decode": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]),
        "pa_prefill": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]),
        "pa_decode_append": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, vp, C.c_int, vp]),
        "pa_decode_step_host": (C.c_int, [vp, C.c_int, vp, vp]),
        "pa_decode_step_host_async": (C.c_int, [vp, C.c_int, vp, vp]),
        "pa_decode_step_host_sync": (C.c_int, [vp]),
        "pa_decode_step_host_layers_async": (C.c_int, [vp, vp, C.c_size_t, vp, C.c_size_t]),
        "pa_decode_step_host_mark": (C.c_int, [vp]),
        "pa_decode_step_host_wait": (C.c_int, [vp, C.c_int]),
        "pa_model_param_count": (C.c_size_t, [C.POINTER(PaModelConfig)]),
        "pa_model_create": (C.c_int, [vp, C.POINTER(PaModelConfig), vp, C.c_ulonglong, C.c_int, C.POINTER(vp)]),
        "pa_model_destroy": (None, [vp]),
        "pa_model_decode_step": (C.c_int, [vp, c_int_p, c_int_p, vp, C.c_int, c_int_p]),
        "pa_model_forward": (C.c_int, [vp, c_int_p, c_int_p, c_int_p, vp, C.c_int, c_int_p]),
        "pa_prefill_schedule": (C.c_int, [vp, C.c_int, C.c_int, c_int_p, C.c_size_t, c_int_p, c_int_p]),
        "pa_model_forward_async": (C.c_int, [vp, c_int_p, c_int_p, c_int_p, vp, C.c_int]),
        "pa_model_wait": (C.c_int, [vp, c_int_p]),
        "pa_model_next_tokens_dev": (vp, [vp]),
        "pa_model_want_device_tokens": (None, [vp, C.c_int]),
        "pa_model_handle": (vp, [vp]),
        "pa_group_create": (C.c_int, [C.POINTER(PaConfig), C.c_int, c_int_p, C.POINTER(vp)]),
        "pa_comm_unique_id": (C.c_int, [vp]),
        "pa_group_join": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(vp)]),
        "pa_group_destroy": (None, [vp]),
        "pa_group_size": (C.c_int, [vp]),
        "pa_group_local_count": (C.c_int, [vp]),
        "pa_group_rank": (C.c_int, [vp, C.c_int]),
        "pa_group_handle": (vp, [vp, C.c_int]),
        "pa_group_gather_tokens": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_int]),
        "pa_group_gather_logits": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.c_size_t]),
        "pa_group_model_step": (C.c_int, [vp, C.POINTER(vp), C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(vp), C.c_int, c_int_p]),
        "pa_group_model_step_overlapped": (C.c_int, [vp, C.POINTER(vp), C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(vp), C.c_int,
                                                     C.POINTER(c_int_p), c_int_p]),
        "pa_group_gather_flush": (C.c_int, [vp, c_int_p]),
        "pa_nccl_version": (C.c_int, []),
        "pa_fill_normal": (C.c_int, [vp, C.c_size_t, C.c_float, C.c_float, C.c_ulonglong, vp]),
        "pa_model_params": (vp, [vp]),
        "pa_model_logits": (vp, [vp, c_int_p]),
        "pa_checkpoint_read_config": (C.c_int, [C.c_char_p, C.POINTER(PaModelConfig)]),
        "pa_checkpoint_read_params": (C.c_int, [C.c_char_p, vp, C.c_size_t]),
        "pa_checkpoint_write": (C.c_int, [C.c_char_p, C.POINTER(PaModelConfig), vp]),
        "pa_checkpoint_write_bf16": (C.c_int, [C.c_char_p, C.POINTER(PaModelConfig), vp]),
        "pa_model_create_from_checkpoint": (C.c_int, [vp, C.c_char_p, C.c_int, C.POINTER(vp)]),
        "pa_dataloader_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]),
        "pa_dataloader_reset": (None, [vp]),
        "pa_dataloader_num_batches": (C.c_int, [vp]),
        "pa_dataloader_next_batch": (C.c_int, [vp, C.POINTER(c_int_p), C.POINTER(c_int_p)]),
        "pa_dataloader_close": (None, [vp]),
        "pa_tokens_write": (C.c_int, [C.c_char_p, c_int_p, C.c_size_t]),
        "pa_tokenizer_open": (C.c_int, [C.c_char_p, C.POINTER(vp)]),
        "pa_tokenizer_vocab_size": (C.c_uint, [vp]),
        "pa_tokenizer_decode": (C.c_char_p, [vp, C.c_uint]),
        "pa_tokenizer_close": (None, [vp]),
        "pa_tokenizer_write": (C.c_int, [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_ubyte), C.c_uint]),
        "pa_seq_fork": (C.c_int, [vp, C.c_int, C.c_int]),
        "pa_set_evict_swap": (C.c_int, [vp, C.c_int]),
        "pa_seq_swap_out": (C.c_int, [vp, C.c_int]),
        "pa_seq_swap_in": (C.c_int, [vp, C.c_int]),
        "pa_seq_swapped_tokens": (C.c_int, [vp, C.c_int]),
        "pa_swap_failures": (C.c_int, [vp]),
        "pa_prefix_insert": (C.c_int, [vp, C.c_int, c_int_p, C.c_int]),
        "pa_prefix_match": (C.c_int, [vp, C.c_int, c_int_p, C.c_int]),
        "pa_prefix_cached_pages": (C.c_int, [vp]),
        "pa_page_refcount": (C.c_int, [vp, C.c_int]),
        "pa_qkv_append": (C.c_int, [vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, vp]),
        "pa_matmul_bias": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        "matmul_forward": (None, [vp, vp, vp, vp] + [C.c_int] * 4),
        "matmul_cached": (None, [vp, vp, vp, vp] + [C.c_int] * 4),
        "pa_seq_len": (C.c_int, [vp, C.c_int]),
        "pa_seq_truncate": (C.c_int, [vp, C.c_int, C.c_int]),
        "pa_seq_free": (C.c_int, [vp, C.c_int]),
        "pa_step_rollback": (C.c_int, [vp]),
        "pa_seq_adopt": (C.c_int, [vp, C.c_int, c_int_p, C.c_int, C.c_int]),
        "pa_pool_k": (vp, [vp, C.c_int]),
        "pa_pool_v": (vp, [vp, C.c_int]),
        "pa_pool_bytes": (C.c_size_t, [vp]),
        "pa_device": (C.c_int, [vp]),
        "pa_stream_of": (vp, [vp]),
        "pa_sm_count": (C.c_int, [vp]),
        "pa_tune_set": (C.c_int, [vp, C.c_int, C.c_int]),
        "pa_tune_get": (C.c_int, [vp, C.c_int]),
        "pa_debug_timeline": (C.c_int, [vp, C.POINTER(C.c_ulonglong), C.c_int]),
        "pa_device_count": (C.c_int, []),
        "pa_set_device": (C.c_int, [C.c_int]),
        "pa_dev_alloc": (vp, [C.c_size_t]),
        "pa_dev_free": (None, [vp]),
        "pa_host_alloc": (vp, [C.c_size_t]),
        "pa_host_free": (None, [vp]),
        "pa_memcpy_h2d": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "pa_memcpy_d2h": (C.c_int, [vp, vp, C.c_size_t, vp]),
        "pa_memset": (C.c_int, [vp, C.c_int, C.c_size_t, vp]),
        "pa_stream_create": (vp, []),
        "pa_stream_destroy": (None, [vp]),
        "pa_stream_sync": (C.c_int, [vp]),
        "pa_device_sync": (C.c_int, []),
        "pa_event_create": (vp, []),
        "pa_event_destroy": (None, [vp]),
        "pa_event_record": (C.c_int, [vp, vp]),
        "pa_event_elapsed_ms": (C.c_float, [vp, vp]),
        "pa_flush_l2": (C.c_int, [vp, C.c_size_t, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    lib._pa_signatures = sig
    _lib = lib
    return lib


def last_error():
    return load().pa_last_error().decode()


def check(rc, what=""):
    if rc != PA_OK:
        raise PagedAttnError(f"{what}: rc={rc}: {last_error()}")


def iptr(a):
    return a.ctypes.data_as(c_int_p)


class DevBuf:
    """A device allocation made through the library (pa_dev_alloc)."""

    def __init__(self, nbytes):
        self.lib = load()
        self.nbytes = int(nbytes)
        self.ptr = self.lib.pa_dev_alloc(max(self.nbytes, 16))
        if not self.ptr:
            raise PagedAttnError(f"pa_dev_alloc({nbytes}): {last_error()}")

    @classmethod
    def from_numpy(cls, a, stream=None):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        b.upload(a, stream)
        return b

    def upload(self, a, stream=None):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(self.lib.pa_memcpy_h2d(self.ptr, a.ctypes.data, a.nbytes, stream), "h2d")
        if stream:
            check(self.lib.pa_stream_sync(stream), "sync")

    def download(self, shape, dtype=np.float32, stream=None, offset_bytes=0):
        out = np.empty(shape, dtype=dtype)
        check(self.lib.pa_memcpy_d2h(out.ctypes.data, self.ptr + offset_bytes, out.nbytes, stream), "d2h")
        if stream:
            check(self.lib.pa_stream_sync(stream), "sync")
        return out

    def free(self):
        if self.ptr:
            self.lib.pa_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PagedAttn:
    """One handle = one GPU's block manager + KV page pool (all layers)."""

    def __init__(self, block_size, max_blocks, max_seqs, n_heads, head_dim, n_layers=1, device=0,
                 max_blocks_per_seq=0, max_batch_tokens=0):
        self.lib = load()
        self.cfg = PaConfig(block_size, max_blocks, max_seqs, max_blocks_per_seq, n_layers, n_heads, head_dim,
                            device, max_batch_tokens)
        h = vp()
        check(self.lib.pa_create(C.byref(self.cfg), C.byref(h)), "pa_create")
        self.h = h
        self.C = n_heads * head_dim
        self.bs, self.max_blocks, self.max_seqs = block_size, max_blocks, max_seqs
        self.n_layers, self.NH, self.hs = n_layers, n_heads, head_dim
        self.mgr = self.lib.pa_manager(self.h)

    def close(self):
        if self.h:
            self.lib.pa_destroy(self.h)
            self.h = None

    # ---- integer side -------------------------------------------------------------------
    def step_begin(self, seq_ids, n_new):
        s = np.ascontiguousarray(seq_ids, dtype=np.int32)
        n = np.ascontiguousarray(n_new, dtype=np.int32)
        return self.lib.pa_step_begin(self.h, iptr(s), iptr(n), len(s))

    def step_begin_readonly(self, seq_ids):
        s = np.ascontiguousarray(seq_ids, dtype=np.int32)
        return self.lib.pa_step_begin_readonly(self.h, iptr(s), len(s))

    def step_set_kv_start(self, kv_start):
        k = np.ascontiguousarray(kv_start, dtype=np.int32)
        return self.lib.pa_step_set_kv_start(self.h, iptr(k))

    def slot_mapping(self):
        n = C.c_int()
        p = self.lib.pa_step_slot_mapping(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy() if n.value else np.zeros(0, np.int32)

    def context_lens(self):
        n = C.c_int()
        p = self.lib.pa_step_context_lens(self.h, C.byref(n))
        return np.ctypeslib.as_array(p, (n.value,)).copy()

    def step_block_table(self):
        n, st = C.c_int(), C.c_int()
        p = self.lib.pa_step_block_table(self.h, C.byref(n), C.byref(st))
        return np.ctypeslib.as_array(p, (n.value, st.value)).copy()

    def seq_len(self, s):
        return self.lib.pa_seq_len(self.h, s)

    def seq_truncate(self, s, n):
        return self.lib.pa_seq_truncate(self.h, s, n)

    def step_rollback(self):
        return self.lib.pa_step_rollback(self.h)

    def seq_free(self, s):
        return self.lib.pa_seq_free(self.h, s)

    def seq_fork(self, src, dst):
        return self.lib.pa_seq_fork(self.h, src, dst)

    def prefix_insert(self, s, tokens):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        return self.lib.pa_prefix_insert(self.h, s, iptr(t), len(t))

    def prefix_match(self, s, tokens):
        t = np.ascontiguousarray(tokens, dtype=np.int32)
        return self.lib.pa_prefix_match(self.h, s, iptr(t), len(t))

    def seq_adopt(self, s, blocks, n_tokens):
        b = np.ascontiguousarray(blocks, dtype=np.int32)
        return self.lib.pa_seq_adopt(self.h, s, iptr(b), len(b), n_tokens)

    def table(self, p):
        n = self.mgr.contents.prompt_block_count[p]
        return [self.mgr.contents.prompt_block_list[p][i] for i in range(n)]

    # ---- device side ---------------------------------------------------------------------
    def upload(self, stream=None):
        return self.lib.pa_step_upload(self.h, stream)

    def append(self, layer, k_ptr, v_ptr, row_stride, stream=None):
        return self.lib.pa_append(self.h, layer, k_ptr, v_ptr, row_stride, stream)

    def decode(self, layer, q_ptr, q_stride, out_ptr, out_stride, stream=None):
        return self.lib.pa_decode(self.h, layer, q_ptr, q_stride, out_ptr, out_stride, stream)

    def decode_append(self, layer, q_ptr, k_ptr, v_ptr, row_stride, out_ptr, out_stride, stream=None):
        return self.lib.pa_decode_append(self.h, layer, q_ptr, k_ptr, v_ptr, row_stride, out_ptr, out_stride, stream)

    def prefill(self, layer, q_ptr, q_stride, out_ptr, out_stride, stream=None):
        return self.lib.pa_prefill(self.h, layer, q_ptr, q_stride, out_ptr, out_stride, stream)

    def qkv_append(self, layer, x_ptr, x_stride, w_ptr, bias_ptr, q_ptr, q_stride, stream=None):
        return self.lib.pa_qkv_append(self.h, layer, x_ptr, x_stride, w_ptr, bias_ptr, q_ptr, q_stride, stream)

    def decode_step_host_async(self, layer, qkv_ptr, out_ptr):
        return self.lib.pa_decode_step_host_async(self.h, layer, qkv_ptr, out_ptr)

    def decode_step_host(self, layer, qkv_ptr, out_ptr):
        return self.lib.pa_decode_step_host(self.h, layer, qkv_ptr, out_ptr)

    def pool_k(self, layer=0):
        return self.lib.pa_pool_k(self.h, layer)

    def pool_v(self, layer=0):
        return self.lib.pa_pool_v(self.h, layer)

    def tune(self, key, value):
        return self.lib.pa_tune_set(self.h, key, value)

    def debug_timeline(self):
        buf = np.zeros((2048, 8), dtype=np.uint64)
        n = self.lib.pa_debug_timeline(self.h, buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 2048)
        if n < 0:
            raise PagedAttnError(last_error())
        return buf[:n]

    def launches(self):
        return self.lib.pa_tune_get(self.h, PA_TUNE_COUNT_LAUNCHES)

    def sync(self):
        check(self.lib.pa_device_sync(), "sync")

    def write_pool_rows(self, layer, slots, k_rows, v_rows):
        """Test helper: put K/V rows at pool slots with plain copies (NOT the append kernel)."""
        k_rows = np.ascontiguousarray(k_rows, dtype=np.float32)
        v_rows = np.ascontiguousarray(v_rows, dtype=np.float32)
        row_bytes = self.C * 4
        pk, pv = self.pool_k(layer), self.pool_v(layer)
        slots = np.asarray(slots)
        # coalesce runs of consecutive slots into single copies
        i = 0
        while i < len(slots):
            j = i
            while j + 1 < len(slots) and slots[j + 1] == slots[j] + 1:
                j += 1
            n = j - i + 1
            check(self.lib.pa_memcpy_h2d(pk + int(slots[i]) * row_bytes, k_rows[i:j + 1].ctypes.data, n * row_bytes, None), "h2d")
            check(self.lib.pa_memcpy_h2d(pv + int(slots[i]) * row_bytes, v_rows[i:j + 1].ctypes.data, n * row_bytes, None), "h2d")
            i = j + 1

    def read_pool_rows(self, layer, slots):
        row_bytes = self.C * 4
        pk, pv = self.pool_k(layer), self.pool_v(layer)
        k = np.empty((len(slots), self.C), dtype=np.float32)
        v = np.empty((len(slots), self.C), dtype=np.float32)
        for i, s in enumerate(slots):
            check(self.lib.pa_memcpy_d2h(k[i].ctypes.data, pk + int(s) * row_bytes, row_bytes, None), "d2h")
            check(self.lib.pa_memcpy_d2h(v[i].ctypes.data, pv + int(s) * row_bytes, row_bytes, None), "d2h")
        return k, v


class ManagerAdapter:
    """Drives a BlockManager* of the library through the interface tests/trace_driver.py uses."""
    kind = "product"

    def __init__(self, mgr):
        self.lib = load()
        self.m = mgr
        c = mgr.contents
        self.bs, self.max_blocks, self.max_prompts = c.block_size, c.max_blocks, c.max_prompts

    def _idx(self, kvp):
        if not kvp:
            return -1
        return (C.addressof(kvp.contents) - C.addressof(self.m.contents.blocks.contents)) // C.sizeof(KVBlock)

    def request_block(self, p):
        return self._idx(self.lib.request_block(self.m, p))

    def get_current_block(self, p):
        return self._idx(self.lib.get_current_block(self.m, p))

    def free_blocks_for_prompt(self, p):
        self.lib.free_blocks_for_prompt(self.m, p)

    def find_lru(self):
        return self.lib.find_least_recently_used_block(self.m)

    def page_out_lru(self):
        self.lib.page_out_lru_block(self.m)

    def get_next_block_id(self, p, bid):
        return self.lib.get_next_block_id(self.m, p, bid)

    def touch(self, idx):
        c = self.m.contents
        c.lru_epoch += 1
        c.blocks[idx].lru_counter = c.lru_epoch

    def set_filled(self, idx, f):
        self.m.contents.blocks[idx].filled = f

    def choose_page(self, p):
        """paged_infer.c:518-529 written against the compat API, as a host caller would."""
        cur = self.get_current_block(p)
        if cur >= 0:
            if self.m.contents.blocks[cur].filled >= self.bs:
                cur = self.request_block(p)
            else:
                self.touch(cur)
        else:
            cur = self.request_block(p)
        return cur

    def epoch(self):
        return self.m.contents.lru_epoch

    def table(self, p):
        c = self.m.contents
        return [c.prompt_block_list[p][i] for i in range(c.prompt_block_count[p])]

    def block_info(self, idx):
        b = self.m.contents.blocks[idx]
        return b.filled, b.prompt_id, b.lru_counter


class Model:
    """pa_model_*: the whole decode step (embedding, L layers, LM head, sampler) on a PagedAttn handle."""

    def __init__(self, eng, max_seq_len, vocab_size, params=None, seed=1337, max_batch=None):
        self.eng = eng
        self.lib = eng.lib
        cfg = eng.cfg
        self.cfg = PaModelConfig(max_seq_len, vocab_size, cfg.n_layers, cfg.n_heads, cfg.n_heads * cfg.head_dim)
        self.V = vocab_size
        self.n_params = self.lib.pa_model_param_count(C.byref(self.cfg))
        ptr = None
        if params is not None:
            params = np.ascontiguousarray(params, dtype=np.float32)
            assert params.size == self.n_params, (params.size, self.n_params)
            ptr = params.ctypes.data
        self.m = C.c_void_p()
        check(self.lib.pa_model_create(eng.h, C.byref(self.cfg), ptr, seed, max_batch or cfg.max_seqs, C.byref(self.m)),
              "pa_model_create")

    def decode_step(self, seq_ids, tokens, coins=None):
        seq = np.ascontiguousarray(seq_ids, dtype=np.int32)
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        nxt = np.zeros(len(seq), dtype=np.int32)
        cptr = None
        if coins is not None:
            coins = np.ascontiguousarray(coins, dtype=np.float32)
            cptr = coins.ctypes.data
        check(self.lib.pa_model_decode_step(self.m, iptr(seq), iptr(tok), cptr, len(seq), iptr(nxt)), "pa_model_decode_step")
        return nxt

    def forward(self, seq_ids, n_new, tokens, coins=None):
        seq = np.ascontiguousarray(seq_ids, dtype=np.int32)
        nn = np.ascontiguousarray(n_new, dtype=np.int32)
        tok = np.ascontiguousarray(tokens, dtype=np.int32)
        nxt = np.zeros(len(seq), dtype=np.int32)
        cptr = None
        if coins is not None:
            coins = np.ascontiguousarray(coins, dtype=np.float32)
            cptr = coins.ctypes.data
        check(self.lib.pa_model_forward(self.m, iptr(seq), iptr(nn), iptr(tok), cptr, len(seq), iptr(nxt)), "pa_model_forward")
        return nxt

    def logits(self, nseq):
        stride = C.c_int(0)
        p = self.lib.pa_model_logits(self.m, C.byref(stride))
        buf = np.zeros((nseq, stride.value), dtype=np.float32)
        check(self.lib.pa_memcpy_d2h(buf.ctypes.data, p, buf.nbytes, None), "d2h")
        return buf[:, :self.V]

    def close(self):
        if self.m:
            self.lib.pa_model_destroy(self.m)
            self.m = None
