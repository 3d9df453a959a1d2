"""TEST INFRASTRUCTURE — pure-Python restatement of the reference's endpoint group.

Follows /root/reference/internal/loadbalancer/group.go:25-150, balance_chwbl.go:14-162 and
balance_least_load.go:3-23 line by line (single-threaded: the blocking/broadcast part is reduced
to "returns None when the Go code would block").  Pinned by replaying the reference's own tables
(load_balancer_test.go:131-427, :19-129) in tests/test_router.py; the C++ router
(kubeai_b200/csrc/router.cc) is then checked against this and against the same tables.
"""
from __future__ import annotations

import bisect

from .weights import xxh64

LEAST_LOAD, PREFIX_HASH = 0, 1


class Group:
    def __init__(self, replication: int):
        self.endpoints = {}          # name -> dict(address, adapters:set, in_flight:int)
        self.total_in_flight = 0
        self.replication = replication
        self.hashes = {}             # hash -> endpoint name
        self.sorted = []

    # balance_chwbl.go:140-150
    @staticmethod
    def hash(s: str) -> int:
        return xxh64(s.encode("utf-8"), 0)

    @staticmethod
    def replica_input(name: str, i: int) -> str:
        return f"{name}{i}"

    # group.go:108-137
    def reconcile(self, observed: dict):
        for name, ep in observed.items():
            if name in self.endpoints:
                self.endpoints[name]["adapters"] = set(ep.get("adapters", ()))
            else:
                self.endpoints[name] = dict(address=ep["address"], adapters=set(ep.get("adapters", ())), in_flight=0)
                for i in range(self.replication):      # balance_chwbl.go:86-97
                    h = self.hash(self.replica_input(name, i))
                    self.hashes[h] = name
                    self.sorted.append(h)
                self.sorted.sort()
        for name in list(self.endpoints):
            if name not in observed:
                for i in range(self.replication):      # balance_chwbl.go:99-105
                    h = self.hash(self.replica_input(name, i))
                    self.hashes.pop(h, None)
                    j = bisect.bisect_left(self.sorted, h)
                    if j < len(self.sorted) and self.sorted[j] == h:
                        del self.sorted[j]
                del self.endpoints[name]

    # balance_chwbl.go:152-162
    @staticmethod
    def load_ok(load, total, n, factor):
        if total == 0:
            return True
        return float(load) <= (float(total + 1) / float(n)) * factor

    # balance_chwbl.go:14-84
    def chwbl(self, key: str, factor: float, adapter: str):
        if not self.sorted:
            return None
        h = self.hash(key)
        i = bisect.bisect_left(self.sorted, h)
        if i >= len(self.sorted):
            i = 0
        default = None
        for _ in range(len(self.sorted)):
            name = self.hashes[self.sorted[i]]
            ep = self.endpoints[name]
            if adapter == "" or adapter in ep["adapters"]:
                if default is None:
                    default = name
                if self.load_ok(ep["in_flight"], self.total_in_flight, len(self.endpoints), factor):
                    return name
            i += 1
            if i >= len(self.sorted):
                i = 0
        return default

    # balance_least_load.go:3-23 (Go map order is random; ties are broken by name here)
    def least_load(self, adapter: str):
        best, mn = None, 0
        for name in sorted(self.endpoints):
            ep = self.endpoints[name]
            if adapter and adapter not in ep["adapters"]:
                continue
            if best is None or ep["in_flight"] < mn:
                best, mn = name, ep["in_flight"]
        return best

    # group.go:53-88 — returns (address, name) or None where the reference would block
    def pick(self, strategy, adapter="", prefix="", mean_load_pct=125):
        if not self.endpoints:
            return None
        if strategy == PREFIX_HASH:
            name = self.chwbl(adapter + prefix, mean_load_pct / 100.0, adapter)
        elif strategy == LEAST_LOAD:
            name = self.least_load(adapter)
        else:
            raise ValueError(f"unknown load balancing strategy: {strategy}")
        if name is None:
            return None
        self.add_in_flight(name, 1)
        return self.endpoints[name]["address"], name

    def add_in_flight(self, name, d):        # group.go:147-150
        self.total_in_flight += d
        self.endpoints[name]["in_flight"] += d

    def done(self, name):
        self.total_in_flight -= 1
        if name in self.endpoints:
            self.endpoints[name]["in_flight"] -= 1
