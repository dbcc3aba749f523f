"""Timing driver (not a test): SA-Open-1.0 Oobleck decode of 1024 latents (47.55 s of audio), CUDA events.
usage: [SATB_*=...] python tests/dec_time.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
from stable_audio_tools import _native as nat
if os.environ.get("SATB_LIB"):          # A/B of kernel variants built side by side (tools only)
    nat.LIB_PATH = os.path.abspath(os.environ["SATB_LIB"])
from oracle import oobleck_oracle as oo
from stable_audio_tools.models.autoencoders import OobleckDecoder

dcfg = dict(out_channels=2, channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], latent_dim=64,
            use_snake=True, final_tanh=False)
dec = OobleckDecoder(**dcfg, operand_dtype=os.environ.get("DEC_DTYPE", "fp16"))
dec.load_state_dict(oo.make_oobleck_weights(oo.decoder_param_shapes(dcfg), seed=9,
                                            transposed=oo.decoder_transposed_prefixes(dcfg)))
dec = dec.cuda().eval()
z = torch.randn(1, 64, 1024).cuda()
for _ in range(3):
    y = dec(z)
torch.cuda.synchronize()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y = dec(z)
e1.record()
torch.cuda.synchronize()
# box-speed reference (boxes of the pool differ by several %): a cuBLAS bf16 GEMM and a device copy
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    c = a @ b
r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
r0.record()
for _ in range(20):
    c = a @ b
r1.record()
torch.cuda.synchronize()
gemm_tf = 20 * 2 * 8192 ** 3 / (r0.elapsed_time(r1) / 1e3) / 1e12
src = torch.empty(1 << 30, device="cuda", dtype=torch.uint8)
dst = torch.empty_like(src)
dst.copy_(src)
r0.record()
for _ in range(10):
    dst.copy_(src)
r1.record()
torch.cuda.synchronize()
copy_gbs = 10 * 2 * (1 << 30) / (r0.elapsed_time(r1) / 1e3) / 1e9
print("box: cublas bf16 %.0f TF/s, copy %.0f GB/s" % (gemm_tf, copy_gbs))
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SATB_") or k == "DEC_DTYPE")
print("decode_ms %.3f  [%s]" % (e0.elapsed_time(e1) / reps, tag), flush=True)
