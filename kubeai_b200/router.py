"""ctypes driver of the C++ router (b200_router_*), shaped like the Go LoadBalancer API
(internal/loadbalancer/load_balancer.go:191 AwaitBestAddress -> (addr, done, err))."""
from __future__ import annotations

import ctypes as C

from ._lib import B200Error, check, lib

LEAST_LOAD, PREFIX_HASH = 0, 1


class DeadlineExceeded(TimeoutError):
    """context.DeadlineExceeded of the Go API."""


class Router:
    def __init__(self, replication: int = 256):
        self._l = lib()
        self._h = C.c_void_p()
        check(self._l.b200_router_create(replication, C.byref(self._h)))

    def close(self):
        if self._h:
            self._l.b200_router_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reconcile_endpoints(self, endpoints: dict):
        """endpoints: name -> {"address": str, "adapters": iterable}"""
        names = list(endpoints)
        n = len(names)
        arr = lambda xs: (C.c_char_p * max(n, 1))(*xs) if n else (C.c_char_p * 1)()
        check(self._l.b200_router_set_endpoints(
            self._h, arr([s.encode() for s in names]), arr([endpoints[s]["address"].encode() for s in names]),
            arr([",".join(sorted(endpoints[s].get("adapters", ()))).encode() for s in names]), n))

    def await_best_address(self, strategy=LEAST_LOAD, adapter="", prefix="", mean_load_pct=125, timeout_s=0.001):
        buf = C.create_string_buffer(256)
        tok = C.c_uint64()
        p = prefix.encode("utf-8")
        rc = self._l.b200_router_pick(self._h, strategy, adapter.encode(), p, len(p), mean_load_pct,
                                      -1 if timeout_s is None else int(timeout_s * 1e6), buf, 256, C.byref(tok))
        if rc == -7:
            raise DeadlineExceeded(self._l.b200_last_error().decode())
        check(rc)
        token = tok.value
        return buf.value.decode(), (lambda: check(self._l.b200_router_done(self._h, token)))

    def add_in_flight(self, name: str, delta: int):
        check(self._l.b200_router_add_inflight(self._h, name.encode(), delta))

    def metrics(self) -> str:
        n = self._l.b200_router_metrics(self._h, None, 0)
        buf = C.create_string_buffer(n + 1)
        self._l.b200_router_metrics(self._h, buf, n + 1)
        return buf.value.decode()

    def in_flight(self, name: str | None = None):
        ep, tot = C.c_int64(), C.c_int64()
        check(self._l.b200_router_inflight(self._h, name.encode() if name else None, C.byref(ep), C.byref(tot)))
        return ep.value, tot.value
