"""The C-ABI library loads without a GPU and exports every symbol include/paged_attn.h declares."""
import ctypes as C
import os
import re
import subprocess

import pytest

import __graft_entry__ as ge

ROOT = ge.ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "paged_attn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))
    return sorted(set(re.findall(r"PA_API\s+[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)))


def test_header_declares_reference_names():
    names = header_symbols()
    # the reference's own interface for this path (block_manager.c:25-201, paged_infer.c:163,505)
    for ref_name in ("create_block_manager", "request_block", "get_current_block", "free_blocks_for_prompt",
                     "find_least_recently_used_block", "page_out_lru_block", "get_next_block_id",
                     "collect_kv_blocks", "print_state", "add_to_cache", "attention_paged"):
        assert ref_name in names
    assert len(names) > 40


def test_library_exports_every_declared_symbol():
    pa = ge.load_binding()
    lib = pa.load()
    missing = [n for n in header_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    # and the binding covers every declared symbol with a signature
    unbound = [n for n in header_symbols() if n not in lib._pa_signatures]
    assert not unbound, unbound


def test_no_internal_symbols_leak():
    pa = ge.load_binding()
    out = subprocess.run(["nm", "-D", "--defined-only", pa.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(header_symbols()), exported ^ set(header_symbols())


def test_no_oracle_in_product():
    """The product path must not link or name the oracle."""
    pa = ge.load_binding()
    ldd = subprocess.run(["ldd", pa.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libref" not in ldd
    for dirpath, _, files in os.walk(ge.PKG):
        for f in files:
            if f.endswith((".c", ".cu", ".h", ".py")):
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "oracle_api" not in src and "libref_" not in src, f


def test_compute_fails_loudly_without_device():
    pa = ge.load_binding()
    lib = pa.load()
    if lib.pa_device_count() > 0:
        pytest.skip("a GPU is present")
    cfg = pa.PaConfig(16, 8, 2, 0, 1, 2, 64, 0, 0)
    h = C.c_void_p()
    rc = lib.pa_create(C.byref(cfg), C.byref(h))
    assert rc == pa.PA_ERR_NO_DEVICE and "no CPU fallback" in pa.last_error()
    # host-only handles schedule but refuse to compute
    eng = pa.PagedAttn(16, 8, 2, 2, 64, device=pa.PA_HOST_ONLY)
    try:
        assert eng.step_begin([0, 1], [1, 1]) == 0
        assert eng.upload() == pa.PA_ERR_NO_DEVICE
        assert eng.decode(0, None, 128, None, 128) == pa.PA_ERR_NO_DEVICE
        assert eng.append(0, None, None, 128) == pa.PA_ERR_NO_DEVICE
        assert eng.decode_step_host(0, None, None) == pa.PA_ERR_NO_DEVICE
        assert eng.prefill(0, None, 128, None, 128) == pa.PA_ERR_NO_DEVICE
        assert eng.qkv_append(0, None, 128, None, None, None, 128) == pa.PA_ERR_NO_DEVICE
        mcfg = pa.PaModelConfig(32, 100, 1, 2, 128)
        m = C.c_void_p()
        assert lib.pa_model_create(eng.h, C.byref(mcfg), None, 1, 2, C.byref(m)) == pa.PA_ERR_NO_DEVICE
        assert "no CPU fallback" in pa.last_error()
    finally:
        eng.close()
