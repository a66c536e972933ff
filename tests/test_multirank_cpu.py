"""N>1 host logic on CPU: world_size 2 over gloo.  Each rank owns its own block manager (host-only
handle) for its shard of the sequences -- the path has no data-path collective; the only exchange
is the per-step gather of one int32 per sequence (the sampled-token gather of north_star), here
the context lengths.  Rank 0 checks the gathered state against a single-process run."""
import os
import subprocess
import sys
import textwrap

import __graft_entry__ as ge

WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
    import __graft_entry__ as ge
    pa = ge.load_binding()

    def run_shard(seqs_global, steps, seed):
        """Deterministic decode trace for a shard; returns (tables, context lens, slot log)."""
        n = len(seqs_global)
        eng = pa.PagedAttn(16, 64, n, 2, 64, device=pa.PA_HOST_ONLY, max_batch_tokens=256)
        rng = np.random.default_rng(seed)
        prompt = [int(rng.integers(1, 40)) for _ in seqs_global]
        assert eng.step_begin(list(range(n)), prompt) == 0
        slots = [eng.slot_mapping().tolist()]
        for _ in range(steps):
            assert eng.step_begin(list(range(n)), [1] * n) == 0
            slots.append(eng.slot_mapping().tolist())
        out = dict(tables=[eng.table(i) for i in range(n)], ctx=eng.context_lens().tolist(), slots=slots)
        eng.close()
        return out

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B_total = 10
    mine = [s for s in range(B_total) if s % world == rank]          # sequences sharded round-robin
    res = run_shard(mine, steps=20, seed=100 + rank)
    # the per-step exchange: one int32 per sequence, gathered on every rank
    local = torch.tensor(res["ctx"] + [-1] * (B_total - len(res["ctx"])), dtype=torch.int32)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)          # max-over-ranks as bench.py does
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    objs = [None] * world
    dist.all_gather_object(objs, res)
    if rank == 0:
        ok = float(t.item()) == world
        for r in range(world):
            shard = [s for s in range(B_total) if s % world == r]
            want = run_shard(shard, steps=20, seed=100 + r)          # single-process rerun of that shard
            ok = ok and want == objs[r]
            ok = ok and gathered[r][:len(shard)].tolist() == want["ctx"]
        print("MULTIRANK_OK" if ok else "MULTIRANK_MISMATCH")
    dist.barrier()
    dist.destroy_process_group()
''')


def test_two_ranks_gloo(tmp_path):
    ge.build(quiet=True)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ge.ROOT, tests=ge.TESTS))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       capture_output=True, text=True, timeout=280, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "MULTIRANK_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
