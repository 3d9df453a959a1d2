#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 1500 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/ac_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/ac_tests.log
python - <<'PY'
# single-stream and 8-stream decode latency at long context: split-KV vs one CTA per (sequence, KV head)
import os, sys, time, numpy as np, torch
sys.path.insert(0, ".")
from kubeai_b200.engine import Engine, default_config
for nseq, ctx in ((1, 2000), (1, 500), (4, 2000), (8, 1000), (16, 2000)):
    row = {}
    for name, env in (("split", None), ("one_cta", "1")):
        os.environ.pop("B200_ATTN_SPLIT", None)
        if env: os.environ["B200_ATTN_SPLIT"] = env
        with Engine(default_config(manual_step=1, max_num_seqs=16, max_batched_tokens=2048, max_model_len=2048, kv_fraction=0.2)) as e:
            rng = np.random.default_rng(0)
            for i in range(nseq):
                e.submit(rng.integers(0, 128000, size=ctx - i).tolist(), max_tokens=40)
            us = []
            for _ in range(60):
                ran, info = e.step()
                if not ran: break
                if info.prefill_seqs == 0 and info.decode_seqs == nseq: us.append(info.device_us)
            row[name] = float(np.median(us[3:]))
    print(f"{nseq:2d} sequences x ctx {ctx:4d}: decode step {row['split']:7.1f} us with split-KV, {row['one_cta']:7.1f} us without ({row['one_cta'] / row['split']:.2f}x)")
PY
