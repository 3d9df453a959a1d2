"""Timeline of the pair-GEMM launches inside decode steps of the engine: %globaltimer stamps of every CTA of every
GEMM launch (B200_GEMM_TRACE_LAUNCHES blocks), printed as per-launch [first entry, first MMA data, last MMA, last exit]
relative to the step's first GEMM, plus the gap to the next GEMM.   usage: python scripts/step_trace.py"""
import ctypes as C, os, sys
NL = 129 * 2
os.environ["B200_GEMM_TRACE_LAUNCHES"] = str(NL)
import torch
sys.path.insert(0, ".")
from kubeai_b200 import lib
from kubeai_b200.engine import Engine, default_config

torch.manual_seed(0)
eng = Engine(default_config(manual_step=1, max_num_seqs=128, max_batched_tokens=2048, max_model_len=2048, kv_fraction=0.5))
for i in range(128):
    eng.submit(torch.randint(0, 128000, (400 + (i * 7) % 90,)).tolist(), max_tokens=64)
while True:                       # run until every sequence is decoding
    ran, info = eng.step()
    if info.prefill_seqs == 0 and info.decode_seqs == 128:
        break
for _ in range(5):
    eng.step()
tr = torch.zeros(NL, 148, 16, dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
lib().b200_op_gemm_trace(C.c_void_p(tr.data_ptr()))
eng.step(); eng.step()
torch.cuda.synchronize()
lib().b200_op_gemm_trace(None)
t = tr.cpu().double()[129:]       # the second traced step
t0 = t[0, :, 0][t[0, :, 0] > 0].min()
names = ["qkv", "o", "gate_up", "down"]
rows = []
for i in range(129):
    b = t[i]
    ok = b[:, 0] > 0
    ent, ext = b[ok, 0], b[ok, 7]
    mma0 = b[:, 10][b[:, 10] > 0]
    mma1 = b[:, 11][b[:, 11] > 0]
    dep = b[:, 12][b[:, 12] > 0]
    rows.append(((ent.min() - t0) / 1e3, (dep.min() - t0) / 1e3, (mma0.min() - t0) / 1e3, (mma0.max() - t0) / 1e3, (mma1.max() - t0) / 1e3, (ext.max() - t0) / 1e3))
print("launch kind      first-entry  dep-wait-passed  first-mma(min..max)  last-mma  last-exit | busy  gap-to-next-mma")
acc = {}
for i, r in enumerate(rows):
    kind = "lm_head" if i == 128 else names[i % 4]
    nxt = rows[i + 1][2] if i + 1 < len(rows) else float("nan")
    busy = r[4] - r[2]
    gap = nxt - r[4]
    a = acc.setdefault(kind, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += busy; a[2] += (gap if gap == gap else 0.0); a[3] += r[2] - r[0]
    if i < 12 or i >= 124:
        print(f"{i:4d} {kind:8s} {r[0]:10.2f} {r[1]:14.2f} {r[2]:12.2f}..{r[3]:7.2f} {r[4]:9.2f} {r[5]:9.2f} | {busy:6.2f} {gap:8.2f}")
print("per kind: launches, mean MMA-active span (first data -> last MMA), mean gap (last MMA -> next GEMM's first data), mean entry->first data")
for k, a in acc.items():
    print(f"  {k:8s} n={a[0]:3d}  busy {a[1] / a[0]:7.2f} us   gap after {a[2] / a[0]:7.2f} us   entry->data {a[3] / a[0]:6.2f} us")
print(f"step span (first GEMM entry -> lm_head exit): {rows[-1][5] - rows[0][0]:.1f} us")
eng.close()
