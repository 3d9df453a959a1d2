"""N>1 host logic on CPU: two gloo ranks shard the synthetic session list with the CHWBL ring and
aggregate their results (tokens add, time is the max) — the multi-process part of bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kubeai_b200.server import harness_config, synth_threads
    from kubeai_b200.sharding import aggregate, assign_threads, first_user_prefix, gather_ranks
    threads = synth_threads(harness_config(synth_threads=400, seed=2))
    mine = assign_threads(threads, world)[rank]
    keys = sorted(first_user_prefix(t) for t in mine)
    # every rank computes the same global list and the same partition: check by exchanging counts and a checksum
    n = torch.tensor([len(mine), sum(len(k) for k in keys)], dtype=torch.int64)
    gathered = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(gathered, n)
    tok, sec = aggregate(dist, tokens=1000.0 * (rank + 1), seconds=2.0 + rank)
    rows = gather_ranks(dist, [1000.0 * (rank + 1), 2.0 + rank, float(rank)])
    if rank == 0:
        out.put(dict(total=len(threads), counts=[int(g[0]) for g in gathered], tok=tok, sec=sec, rows=rows,
                     all_keys=sorted(first_user_prefix(t) for t in threads),
                     parts=[sorted(first_user_prefix(t) for t in p) for p in assign_threads(threads, world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_shard_sessions_with_the_ring_and_aggregate():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert sum(res["counts"]) == res["total"] == 400            # disjoint and complete
    assert sorted(res["parts"][0] + res["parts"][1]) == res["all_keys"]
    assert res["counts"] == [len(res["parts"][0]), len(res["parts"][1])]
    assert min(res["counts"]) > 0.35 * res["total"]             # 256 vnodes per replica: roughly even
    assert res["tok"] == 3000.0 and res["sec"] == 3.0           # tokens add, time is the max over ranks
    assert res["rows"] == [[1000.0, 2.0, 0.0], [2000.0, 3.0, 1.0]]   # bench.py prints every rank's own figures too


def test_assignment_is_stable_when_a_replica_is_added():
    """Consistent hashing: going from 4 to 5 replicas moves only ~1/5 of the sessions."""
    from kubeai_b200.server import harness_config, synth_threads
    from kubeai_b200.sharding import assign_threads, first_user_prefix
    threads = synth_threads(harness_config(synth_threads=600, seed=5))
    where = lambda parts: {first_user_prefix(t): i for i, p in enumerate(parts) for t in p}
    a, b = where(assign_threads(threads, 4)), where(assign_threads(threads, 5))
    moved = sum(1 for k in a if a[k] != b[k])
    assert all(b[k] == 4 for k in a if a[k] != b[k]), "sessions only ever move to the new replica"
    assert 0.08 * len(a) < moved < 0.35 * len(a)
