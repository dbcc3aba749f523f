"""Debug (not a test): which shared-memory size / carveout lets two attention CTAs share an SM."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
import torch
torch.zeros(1, device="cuda")
from stable_audio_tools import _native as nat
for carve in (-1, 100, 75, 50):
    print("carveout", carve, [(kb, nat.lib().satb_debug_attention_occupancy(kb * 1024, carve)) for kb in (16, 32, 48, 64, 72, 75, 80, 86, 96, 104, 112)], flush=True)
