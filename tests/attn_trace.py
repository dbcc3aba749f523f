"""Per-phase clock64 trace of one attention CTA (not a test): python tests/attn_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
from stable_audio_tools import _native as nat
B, H, N = 8, 24, 1025
q = torch.randn(B, N, H * 64, device="cuda").half(); k = torch.randn(B, N, H * 64, device="cuda").half()
v = torch.randn(B, N, H * 64, device="cuda").half(); o = torch.empty_like(q)
T = (N + 63) // 64
dbg = torch.zeros(T * 12, dtype=torch.int64, device="cuda")
for _ in range(3):
    nat.check(nat.lib().satb_attention_trace(nat.ptr(q), nat.ptr(k), nat.ptr(v), nat.ptr(o), B, H, H, N, N, 0, nat.ptr(dbg), nat.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    nat.check(nat.lib().satb_attention(nat.ptr(q), nat.ptr(k), nat.ptr(v), nat.ptr(o), B, H, H, N, N, 0, nat.stream_ptr()))
e1.record(); torch.cuda.synchronize()
print("kernel us", e0.elapsed_time(e1) * 100)
d = dbg.cpu().view(T, 12)
t0 = int(d[0, 0])
names = ["s_wait_begin", "s_full", "pv_free", "pass_done", "pairmax", "arrived", "m_wait_begin", "p_ready", "v_full", "pv_issued", "qk_issued"]
print("tile " + " ".join(f"{n:>12s}" for n in names))
for j in range(T):
    print(f"{j:4d} " + " ".join(f"{int(d[j, i]) - t0:12d}" for i in range(11)))
