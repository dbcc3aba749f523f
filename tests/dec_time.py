"""Timing driver (not a test): SA-Open-1.0 Oobleck decode of 1024 latents (47.55 s of audio), CUDA events.
usage: [SATB_*=...] python tests/dec_time.py [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
from oracle import oobleck_oracle as oo
from stable_audio_tools.models.autoencoders import OobleckDecoder

dcfg = dict(out_channels=2, channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], latent_dim=64,
            use_snake=True, final_tanh=False)
dec = OobleckDecoder(**dcfg)
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
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SATB_"))
print("decode_ms %.3f  [%s]" % (e0.elapsed_time(e1) / reps, tag), flush=True)
