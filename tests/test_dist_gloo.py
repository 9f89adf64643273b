"""The N>1 path of bench.py on CPU: two ranks over gloo, one independent sequence each, counters reduced
as MAX (time) / SUM (work).  No data-path collective exists for this workload."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    lv = G.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    d = importlib.import_module("limovelo_b200.dist")
    seqs = d.sequence_for_rank(rank, world)
    # rank r "measures" (r+1) ms per step on 1000*(r+1) points
    out = d.reduce_counters(step_ms=10.0 * (rank + 1), points=1000 * (rank + 1), matched=900 * (rank + 1),
                            e2e_s=0.02 * (rank + 1), e2e_points=1000 * (rank + 1), launches=17)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, seqs, out))


def test_two_rank_counter_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0] and res[1][1] == [1]                       # one sequence per rank
    for _, _, out in res:
        assert out["step_ms"] == 20.0 and out["e2e_s"] == 0.04         # MAX over ranks
        assert out["points"] == 3000 and out["matched"] == 2700 and out["launches"] == 34   # SUM
    sys.path.insert(0, ROOT)
    import __graft_entry__ as G
    G.load_package()
    import importlib
    d = importlib.import_module("limovelo_b200.dist")
    assert d.throughput(3000, 20.0) == 150000.0
    assert d.sequence_for_rank(1, 4, 8) == [1, 5]
