"""Tiny driver for ncu captures: SA-Open-1.0 width, depth 2, B=4 CFG (8 rows x 1025 tokens), a few
forwards; plus one Oobleck decode of 256 latents.  Not a test."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import SAO_DIT, build_native_dit
from oracle import dit_oracle as do

what = sys.argv[1] if len(sys.argv) > 1 else "dit"
if what == "dit":
    cfg = dict(SAO_DIT, depth=2)
    m = build_native_dit(cfg, do.make_dit_weights(cfg, seed=5))
    B = 4
    x = torch.randn(B, 64, 1024).cuda(); t = torch.rand(B).cuda()
    c = torch.randn(B, 130, 768).cuda(); ge = torch.randn(B, 1536).cuda()
    for _ in range(3):
        y = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=7.0)
    torch.cuda.synchronize()
else:
    from oracle import oobleck_oracle as oo
    from stable_audio_tools.models.autoencoders import OobleckDecoder
    dcfg = dict(out_channels=2, channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], latent_dim=64,
                use_snake=True, final_tanh=False)
    dec = OobleckDecoder(**dcfg)
    dec.load_state_dict(oo.make_oobleck_weights(oo.decoder_param_shapes(dcfg), seed=9, transposed=oo.decoder_transposed_prefixes(dcfg)))
    dec = dec.cuda().eval()
    z = torch.randn(1, 64, int(sys.argv[2]) if len(sys.argv) > 2 else 256).cuda()
    for _ in range(2):
        y = dec(z)
    torch.cuda.synchronize()
print("done", float(y.abs().mean()))
