"""clock64 trace of CTA 0's first softmax warp over its first 16 key tiles (not a test): python tests/attn_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
from stable_audio_tools import _native as nat
if os.environ.get("SATB_LIB"):          # A/B of kernel variants built side by side (tools only)
    nat.LIB_PATH = os.path.abspath(os.environ["SATB_LIB"])
B, H, N = 8, 24, int(sys.argv[1]) if len(sys.argv) > 1 else 1025
q = torch.randn(B, N, H * 64, device="cuda").half(); k = torch.randn(B, N, H * 64, device="cuda").half()
v = torch.randn(B, N, H * 64, device="cuda").half(); o = torch.empty_like(q)
dbg = torch.zeros(16 * 12 + 4 * 296, dtype=torch.int64, device="cuda")
for _ in range(3):
    nat.check(nat.lib().satb_attention_trace(nat.ptr(q), nat.ptr(k), nat.ptr(v), nat.ptr(o), B, H, H, N, N, 0, nat.ptr(dbg), nat.stream_ptr()))
torch.cuda.synchronize()
allv = dbg.cpu()
d = allv[:192].view(16, 12)
t0 = int(d[0, 0])
names = ["begin", "s_full", "extras", "chunks", "lastexp", "epi", "o_done", "unit_top", "coords", "pre_sync", "stored", "p_ready"]
print("tile " + " ".join(f"{n:>8s}" for n in names))
for j in range(16):
    print(f"{j:4d} " + " ".join(f"{(int(d[j, i]) - t0) if int(d[j, i]) else 0:8d}" for i in range(12)))

res = allv[192:].view(296, 4)
t_min = int(res[:, 2][res[:, 2] > 0].min())
by_sm = {}
for c in range(296):
    sm, slot, a, b = (int(v) for v in res[c])
    if a:
        by_sm.setdefault(sm, []).append((c, slot, (a - t_min) / 1e3, (b - t_min) / 1e3))
overl = sum(1 for v in by_sm.values() if len(v) == 2 and max(v[0][2], v[1][2]) < min(v[0][3], v[1][3]))
print(f"{len(by_sm)} SMs used; SMs whose two CTAs overlap in time: {overl}")
durs = sorted((b - a, c) for v in by_sm.values() for c, slot, a, b in v)
ends = sorted(b for v in by_sm.values() for c, slot, a, b in v)
print("CTA duration us: min %.1f median %.1f max %.1f; last end %.1f; CTAs ending in the last 10 us: %d; longest CTAs %s" % (
    durs[0][0], durs[len(durs) // 2][0], durs[-1][0], ends[-1], sum(1 for e in ends if e > ends[-1] - 10), [c for _, c in durs[-8:]]))
for sm in sorted(by_sm)[:3]:
    print("  SM", sm, [(c, slot, round(a, 1), round(b, 1)) for c, slot, a, b in by_sm[sm]])
