"""Python driver for the engine part of the C ABI (tests, bench, smoke).  Mirrors the calls a cgo
shim would make; holds no model logic."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import Config, Sampling, Stats, StepInfo, Usage, check, lib

FINISH = {0: None, 1: "stop", 2: "length", 3: "aborted", 4: "error"}


def default_config(**kw) -> Config:
    cfg = Config()
    lib().b200_config_default(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def config_from_hf(model_dir: str, **kw) -> Config:
    """Engine config whose architecture comes from an HF config.json."""
    cfg = default_config(**kw)
    check(lib().b200_config_from_hf(str(model_dir).encode(), C.byref(cfg)))
    return cfg


def safetensors_list(path: str) -> dict:
    import json
    n = check(lib().b200_safetensors_list(str(path).encode(), None, 0))
    buf = C.create_string_buffer(n + 1)
    lib().b200_safetensors_list(str(path).encode(), buf, n + 1)
    return json.loads(buf.value)


def mini_config(**kw) -> Config:
    """The 2-layer test model of oracle/weights.py ModelCfg()."""
    base = dict(num_layers=2, hidden=512, q_heads=4, kv_heads=1, intermediate=1024, vocab=512,
                max_model_len=256, max_num_seqs=16, max_batched_tokens=256, num_kv_blocks=128,
                manual_step=1, seed=0)
    base.update(kw)
    return default_config(**base)


@dataclass
class PollResult:
    tokens: list
    finished: str | None
    usage: tuple


class Engine:
    def __init__(self, cfg: Config):
        self._l = lib()
        self._h = C.c_void_p()
        self.cfg = cfg
        check(self._l.b200_engine_create(C.byref(cfg), C.byref(self._h)))

    def close(self):
        if self._h:
            self._l.b200_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit(self, prompt_ids, max_tokens=16, ignore_eos=True, stop_ids=()) -> int:
        ids = np.ascontiguousarray(prompt_ids, dtype=np.int32)
        stops = np.ascontiguousarray(stop_ids, dtype=np.int32)
        sp = Sampling(max_tokens=max_tokens, temperature=0.0, ignore_eos=1 if ignore_eos else 0,
                      num_stop_ids=len(stops),
                      stop_ids=stops.ctypes.data_as(C.POINTER(C.c_int32)) if len(stops) else None)
        rid = C.c_uint64()
        check(self._l.b200_submit(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), len(ids), C.byref(sp),
                                  C.byref(rid)))
        return rid.value

    def poll(self, rid: int, cap: int = 4096) -> PollResult:
        buf = (C.c_int32 * cap)()
        n, fin, u = C.c_int32(), C.c_int32(), Usage()
        check(self._l.b200_poll(self._h, rid, buf, cap, C.byref(n), C.byref(fin), C.byref(u)))
        return PollResult(list(buf[: n.value]), FINISH[fin.value],
                          (u.prompt_tokens, u.cached_tokens, u.completion_tokens))

    def wait(self, rid: int, timeout_s: float = -1.0) -> bool:
        rc = self._l.b200_wait(self._h, rid, int(timeout_s * 1e6) if timeout_s >= 0 else -1)
        if rc == -7:
            return False
        check(rc)
        return True

    def abort(self, rid: int):
        check(self._l.b200_abort(self._h, rid))

    def release(self, rid: int):
        check(self._l.b200_release(self._h, rid))

    def step(self):
        info = StepInfo()
        rc = check(self._l.b200_engine_step(self._h, C.byref(info)))
        return rc == 1, info

    def run(self, max_steps: int, idle_timeout_us: int = 20000):
        """Up to max_steps pipelined iterations (b200_engine_run); returns the StepInfo of every step that ran."""
        infos = (StepInfo * max_steps)()
        n = C.c_int32()
        check(self._l.b200_engine_run(self._h, max_steps, idle_timeout_us, infos, C.byref(n)))
        return [infos[i] for i in range(n.value)]

    def stats(self) -> Stats:
        s = Stats()
        check(self._l.b200_stats_get(self._h, C.byref(s)))
        return s

    def replay(self, n: int, repeat: int = 1):
        ms, tok, smp, kvt, ln = C.c_double(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self._l.b200_engine_replay(self._h, n, repeat, C.byref(ms), C.byref(tok), C.byref(smp), C.byref(kvt),
                                         C.byref(ln)))
        return dict(ms=ms.value, tokens=tok.value, sampled=smp.value, kv_tokens=kvt.value, launches=ln.value)

    KERNEL_CLASSES = ["embed", "rmsnorm", "gemm_qkv", "rope_kvwrite", "attn_decode", "attn_prefill", "gemm_o",
                      "gemm_gate_up", "silu_mul", "gemm_down", "gemm_lm_head", "argmax"]

    def set_skip_mask(self, classes=()):
        mask = 0
        for c in classes:
            mask |= 1 << self.KERNEL_CLASSES.index(c)
        check(self._l.b200_engine_set_skip_mask(self._h, mask))

    def set_recording(self, on: bool):
        check(self._l.b200_engine_set_recording(self._h, 1 if on else 0))

    def profile(self, n: int) -> dict:
        """Per-kernel-class device time (us) and launch counts over the last n recorded steps."""
        k = len(self.KERNEL_CLASSES)
        us, ln = (C.c_double * k)(), (C.c_int64 * k)()
        check(self._l.b200_engine_profile(self._h, n, us, ln, k))
        return {name: dict(us=us[i], launches=ln[i]) for i, name in enumerate(self.KERNEL_CLASSES)}

    def profile_range(self, n: int, min_tokens: int, max_tokens: int):
        """profile() over the recorded steps with min_tokens <= T <= max_tokens; returns (per-class dict, totals dict)."""
        k = len(self.KERNEL_CLASSES)
        us, ln = (C.c_double * k)(), (C.c_int64 * k)()
        st, tk, sm, kv = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        check(self._l.b200_engine_profile_range(self._h, n, min_tokens, max_tokens, us, ln, k, C.byref(st), C.byref(tk),
                                                C.byref(sm), C.byref(kv)))
        return ({name: dict(us=us[i], launches=ln[i]) for i, name in enumerate(self.KERNEL_CLASSES)},
                dict(steps=st.value, tokens=tk.value, sampled=sm.value, kv_tokens=kv.value))

    def reset_prefix_cache(self):
        check(self._l.b200_engine_reset_prefix_cache(self._h))

    def tensor(self, name: str) -> np.ndarray:
        """bf16 bit patterns (uint16) of a named weight, copied to host."""
        nb, p = C.c_uint64(), C.c_void_p()
        check(self._l.b200_engine_tensor_info(self._h, name.encode(), C.byref(nb), C.byref(p)))
        out = np.empty(nb.value // 2, dtype=np.uint16)
        check(self._l.b200_engine_tensor_read(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), nb.value))
        return out

    def tensor_write(self, name: str, bits: np.ndarray):
        bits = np.ascontiguousarray(bits, dtype=np.uint16)
        check(self._l.b200_engine_tensor_write(self._h, name.encode(), bits.ctypes.data_as(C.c_void_p), bits.nbytes))

    def load_safetensors(self, path: str):
        check(self._l.b200_engine_load_safetensors(self._h, str(path).encode()))

    def forward_logits(self, ids) -> np.ndarray:
        """fp32 array [n, vocab] of the bf16 logits for one sequence (debug/parity)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.empty((len(ids), self.cfg.vocab), dtype=np.uint16)
        check(self._l.b200_engine_forward_logits(self._h, ids.ctypes.data_as(C.POINTER(C.c_int32)), len(ids),
                                                 out.ctypes.data_as(C.c_void_p)))
        return (out.astype(np.uint32) << 16).view(np.float32)

    def set_keep_logits(self, on: bool):
        check(self._l.b200_engine_set_keep_logits(self._h, 1 if on else 0))

    def read_logits(self, rows: int) -> np.ndarray:
        """fp32 array [rows, vocab] of the bf16 logits the last step sampled from (after set_keep_logits(True))."""
        out = np.empty((rows, self.cfg.vocab), dtype=np.uint16)
        check(self._l.b200_engine_read_logits(self._h, out.ctypes.data_as(C.c_void_p), rows))
        return (out.astype(np.uint32) << 16).view(np.float32)

    def generate(self, prompts, max_tokens=16, max_steps=100000):
        """Manual-step helper: run all prompts to completion, return the list of token lists."""
        rids = [self.submit(p, max_tokens=max_tokens) for p in prompts]
        outs = [[] for _ in rids]
        done = [False] * len(rids)
        for _ in range(max_steps):
            ran, _info = self.step()
            for i, r in enumerate(rids):
                if not done[i]:
                    pr = self.poll(r)
                    outs[i] += pr.tokens
                    done[i] = pr.finished is not None
            if all(done):
                break
            if not ran:
                raise RuntimeError("engine idle with unfinished requests")
        for r in rids:
            self.release(r)
        return outs
