"""Data-parallel sharding of the session list across replicas: independent replicas,
no collective on the data path.  Each rank keeps the conversation threads that the reference's CHWBL
ring (internal/loadbalancer/balance_chwbl.go:14-84, key = first user message's first 100 runes)
assigns to it; torch.distributed is used only to agree on the start time and to sum / max the
per-rank results at the end."""
from __future__ import annotations

from .router import PREFIX_HASH, Router


def first_user_prefix(thread: dict, n: int = 100) -> str:
    for m in thread.get("messages", ()):
        if m.get("role") == "user":
            return (m.get("content") or "")[:n]      # python slices by code point == Go []rune
    return ""


def assign_threads(threads: list, world: int, replication: int = 256) -> list:
    """rank -> list of threads.  Static assignment: every pick is released at once, so the ring
    position alone decides (bounded-load never triggers) — the same key always maps to the same rank."""
    r = Router(replication)
    r.reconcile_endpoints({f"gpu-{i}": dict(address=f"gpu:{i}") for i in range(world)})
    out = [[] for _ in range(world)]
    for t in threads:
        addr, done = r.await_best_address(PREFIX_HASH, "", first_user_prefix(t), 125, timeout_s=0)
        done()
        out[int(addr.split(":")[1])].append(t)
    r.close()
    return out


def aggregate(dist, tokens: float, seconds: float, device=None):
    """Whole-job throughput over ranks: tokens add, time is the max over ranks."""
    import torch
    t = torch.tensor([tokens], dtype=torch.float64, device=device)
    s = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(s, op=dist.ReduceOp.MAX)
    return t.item(), s.item()


def gather_ranks(dist, values, device=None):
    """values: list of floats of this rank -> list over ranks of those lists (per-rank reporting: tokens add, the
    job's time is the max over ranks, and every rank's own figures are printed next to the aggregate)."""
    import torch
    v = torch.tensor(list(values), dtype=torch.float64, device=device)
    out = [torch.zeros_like(v) for _ in range(dist.get_world_size())]
    dist.all_gather(out, v)
    return [o.tolist() for o in out]
