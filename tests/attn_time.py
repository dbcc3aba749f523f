"""Timing / accuracy driver for the attention kernel (not a test): [SATB_ATTN_POLY=n] python tests/attn_time.py
SA-Open self-attention shape (8 rows x 24 heads x 1025 tokens), CUDA events, error vs torch fp32 softmax."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
from stable_audio_tools import _native as nat

B, H, N = 8, 24, int(sys.argv[1]) if len(sys.argv) > 1 else 1025
torch.manual_seed(0)
q = (2.0 * torch.randn(B, N, H * 64, device="cuda")).half()
k = torch.randn(B, N, H * 64, device="cuda").half()
v = torch.randn(B, N, H * 64, device="cuda").half()
o = torch.empty_like(q)


def run():
    nat.check(nat.lib().satb_attention(nat.ptr(q), nat.ptr(k), nat.ptr(v), nat.ptr(o), B, H, H, N, N, 0, nat.stream_ptr()))


for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1000 / 50
qh, kh, vh = (t[:2].float().view(2, N, H, 64).transpose(1, 2) for t in (q, k, v))
ref = torch.softmax(qh @ kh.transpose(-1, -2) / 8.0, dim=-1) @ vh
got = o[:2].float().view(2, N, H, 64).transpose(1, 2)
err = float((got - ref).norm() / ref.norm())
# box-speed reference
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    c = a @ b
e0.record()
for _ in range(20):
    c = a @ b
e1.record()
torch.cuda.synchronize()
gemm_tf = 20 * 2 * 8192 ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12
print("attention N=%d: %.1f us  (%.0f TF/s)  rel-L2 err %.2e  poly=%s  | box cublas bf16 %.0f TF/s"
      % (N, us, 4.0 * B * H * N * N * 64 / us / 1e6, err, os.environ.get("SATB_ATTN_POLY", "default"), gemm_tf), flush=True)
