"""clock64 trace of CTA 0's first softmax warp over its first 16 key tiles (not a test): python tests/attn_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
from stable_audio_tools import _native as nat
B, H, N = 8, 24, 1025
q = torch.randn(B, N, H * 64, device="cuda").half(); k = torch.randn(B, N, H * 64, device="cuda").half()
v = torch.randn(B, N, H * 64, device="cuda").half(); o = torch.empty_like(q)
dbg = torch.zeros(16 * 4, dtype=torch.int64, device="cuda")
for _ in range(3):
    nat.check(nat.lib().satb_attention_trace(nat.ptr(q), nat.ptr(k), nat.ptr(v), nat.ptr(o), B, H, H, N, N, 0, nat.ptr(dbg), nat.stream_ptr()))
torch.cuda.synchronize()
d = dbg.cpu().view(16, 4)
t0 = int(d[0, 0])
names = ["tile_begin", "s_full", "p_free", "p_ready"]
print("tile " + " ".join(f"{n:>12s}" for n in names))
for j in range(16):
    print(f"{j:4d} " + " ".join(f"{int(d[j, i]) - t0:12d}" for i in range(4)))
