"""Seeded allocator traces run identically on the reference manager, the oracle
restatement and the product's block manager (integer parity)."""
import numpy as np


def make_trace(seed, n_ops, max_prompts, bs, p_free=0.06, p_request=0.08, p_pageout=0.02, n_active=None):
    rng = np.random.default_rng(seed)
    n_active = n_active or max_prompts
    ops = []
    for _ in range(n_ops):
        r = rng.random()
        p = int(rng.integers(0, n_active))
        if r < p_free:
            ops.append(("free", p))
        elif r < p_free + p_request:
            ops.append(("request", p))
        elif r < p_free + p_request + p_pageout:
            ops.append(("page_out",))
        elif r < p_free + p_request + p_pageout + 0.02:
            ops.append(("request", int(rng.choice([-1, max_prompts, max_prompts + 7]))))
        elif r < p_free + p_request + p_pageout + 0.05:
            ops.append(("find_lru",))
        elif r < p_free + p_request + p_pageout + 0.08:
            ops.append(("next", p, int(rng.integers(0, 12))))
        else:
            ops.append(("append", p, int(rng.integers(1, 4))))   # 1..3 decode appends in a row
    return ops


def snapshot(m, prompts, blocks):
    return {
        "epoch": m.epoch(),
        "tables": [m.table(p) for p in prompts],
        "blocks": [m.block_info(i) for i in blocks],
    }


def run_trace(m, ops, max_prompts, max_blocks, snap_every=1):
    """append = the page choice of add_to_cache (paged_infer.c:518-529) followed by
    filled += 1 (:570), i.e. one decode append; the K/V bytes are not part of this test."""
    log = []
    prompts = list(range(max_prompts))
    blocks = list(range(max_blocks))
    for i, op in enumerate(ops):
        kind = op[0]
        if kind == "append":
            res = []
            for _ in range(op[2]):
                idx = m.choose_page(op[1])
                if idx >= 0:
                    f = m.block_info(idx)[0]
                    m.set_filled(idx, f + 1)
                res.append(idx)
        elif kind == "request":
            res = m.request_block(op[1])
        elif kind == "free":
            m.free_blocks_for_prompt(op[1])
            res = None
        elif kind == "page_out":
            m.page_out_lru()
            res = None
        elif kind == "find_lru":
            res = m.find_lru()
        elif kind == "next":
            res = m.get_next_block_id(op[1], op[2])
        else:
            raise ValueError(kind)
        entry = {"op": op, "res": res}
        if i % snap_every == 0 or i == len(ops) - 1:
            entry["snap"] = snapshot(m, prompts, blocks)
        log.append(entry)
    return log
