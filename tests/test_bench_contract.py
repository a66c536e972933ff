"""bench.py's contract, checked without a GPU: the reference arm (the reference's own CPU code from oracle/_ref, else
the oracle port) runs here, prints ONE JSON line with the keys the driver reads, loads nothing of the product, and
carries the same `config` object as the GPU arm would for the same workload and GPU count."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

import __graft_entry__ as ge

BENCH = os.path.join(ge.ROOT, "bench.py")


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", BENCH)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("gpus,workload", [(1, "cfg2"), (4, "cfg4"), (8, "cfg5")])
def test_reference_arm_line(gpus, workload):
    env = dict(os.environ)
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", str(gpus), "--steps", "1", "--warmup", "0",
                        "--workload", workload], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == gpus and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] and "bounded sample" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same config object as the GPU arm builds for this workload at this GPU count
    bench = load_bench()
    wid, w = bench.resolve_workload(workload, gpus)
    assert line["config"] == bench.config_of(wid, w, gpus)
    assert line["config"]["workload_id"] == wid
    # the reference arm never touches the product library
    assert "libpaged_attn" not in r.stderr


def test_workload_table_matches_baseline_configs():
    bench = load_bench()
    base = json.load(open(os.path.join(ge.ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    w = bench.WORKLOADS
    assert (w["cfg1"]["B"], w["cfg1"]["ctx_len"], w["cfg1"]["bs"], w["cfg1"]["L"]) == (1, 256, 16, 12)          # configs[0]
    assert (w["cfg2"]["B"], w["cfg2"]["ctx_len"], w["cfg2"]["NH"], w["cfg2"]["hs"]) == (64, 1024, 12, 64)       # configs[1]
    assert (w["cfg3"]["B"], w["cfg3"]["ctx_lo"], w["cfg3"]["ctx_hi"]) == (256, 128, 1024)                       # configs[2]
    assert (w["cfg4"]["NH"], w["cfg4"]["hs"], w["cfg4"]["L"]) == (25, 64, 48) and bench.TOTAL_XL_BATCH == 512   # configs[3]
    for n, per_gpu in ((2, 256), (4, 128), (8, 64), (1, 256)):
        assert bench.resolve_workload("cfg4", n)[1]["B"] == per_gpu
    assert (w["cfg5"]["ctx_len"], w["cfg5"]["hs"]) == (32768, 128) and w["cfg5"]["prefill_chunk"] > 0           # configs[4]
    # BASELINE.md bytes formula: K and V of valid tokens + q + out + block table + context lengths
    assert bench.decode_bytes([1024] * 64, 768, 16) == 64 * 2 * 1024 * 768 * 4 + 2 * 64 * 768 * 4 + 64 * 64 * 4 + 64 * 4
    assert bench.decode_bytes([1024] * 64, 768, 16) + bench.append_bytes(64, 768) == 403849728
