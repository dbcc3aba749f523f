"""A/B driver (not a test): SA-Open decoder / encoder outputs with the fused ResidualUnit kernel vs the
two-launch path (SATB_RESUNIT=unfused), and each against the fp32 oracle.  Usage: python tests/resunit_ab.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

SAO_VAE = dict(channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], latent_dim=64, use_snake=True)
CASES = [("dec", 1, 8), ("dec", 2, 24), ("dec", 1, 320), ("enc", 1, 24), ("enc", 1, 160)]


def run_child(tag):
    from oracle import oobleck_oracle as oo
    from stable_audio_tools.models.autoencoders import OobleckDecoder, OobleckEncoder
    out = {}
    dcfg = dict(SAO_VAE, out_channels=2, final_tanh=False)
    dsd = oo.make_oobleck_weights(oo.decoder_param_shapes(dcfg), seed=11, transposed=oo.decoder_transposed_prefixes(dcfg))
    dec = OobleckDecoder(**dcfg)
    dec.load_state_dict(dsd)
    dec = dec.cuda().eval()
    ecfg = dict(SAO_VAE, in_channels=2, latent_dim=128)
    esd = oo.make_oobleck_weights(oo.encoder_param_shapes(ecfg), seed=12)
    enc = OobleckEncoder(**ecfg)
    enc.load_state_dict(esd)
    enc = enc.cuda().eval()
    for kind, B, n in CASES:
        torch.manual_seed(4)
        if kind == "dec":
            x = torch.randn(B, 64, n)
            y = dec(x.cuda()).cpu()
        else:
            x = 0.5 * torch.randn(B, 2, n * 2048).clamp(-1, 1)
            y = enc(x.cuda()).cpu()
        out[(kind, B, n)] = y
    torch.save(out, os.path.join(ROOT, "gpurun_out", f"resunit_{tag}.pt"))


def main():
    if len(sys.argv) > 1:
        run_child(sys.argv[1])
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for tag, env in (("fused", {}), ("unfused", {"SATB_RESUNIT": "unfused"})):
        subprocess.run([sys.executable, os.path.abspath(__file__), tag], env=dict(os.environ, **env), check=True, timeout=300)
    from oracle import oobleck_oracle as oo
    from helpers import rel_l2
    f = torch.load(os.path.join(ROOT, "gpurun_out", "resunit_fused.pt"))
    u = torch.load(os.path.join(ROOT, "gpurun_out", "resunit_unfused.pt"))
    dcfg = dict(SAO_VAE, out_channels=2, final_tanh=False)
    dsd = oo.make_oobleck_weights(oo.decoder_param_shapes(dcfg), seed=11, transposed=oo.decoder_transposed_prefixes(dcfg))
    ecfg = dict(SAO_VAE, in_channels=2, latent_dim=128)
    esd = oo.make_oobleck_weights(oo.encoder_param_shapes(ecfg), seed=12)
    for kind, B, n in CASES:
        torch.manual_seed(4)
        if kind == "dec":
            ref = oo.oobleck_decoder(torch.randn(B, 64, n), dsd, dcfg)
        else:
            ref = oo.oobleck_encoder(0.5 * torch.randn(B, 2, n * 2048).clamp(-1, 1), esd, ecfg)
        k = (kind, B, n)
        d = (f[k] - u[k]).abs()
        print(k, "fused-vs-oracle %.3e" % rel_l2(f[k], ref), "unfused-vs-oracle %.3e" % rel_l2(u[k], ref),
              "fused-vs-unfused max %.3e at %s" % (float(d.max()), tuple(int(i) for i in torch.nonzero(d == d.max())[0])),
              "n_diff", int((d > 0).sum()))
        if kind == "dec" and float(d.max()) > 0:
            bad = torch.nonzero(d.amax(dim=(0, 1)) > 1e-3 * float(ref.abs().max())).flatten()
            if len(bad):
                print("   first/last bad position", int(bad[0]), int(bad[-1]), "count", len(bad), "of", d.shape[-1])
    os.remove(os.path.join(ROOT, "gpurun_out", "resunit_fused.pt"))
    os.remove(os.path.join(ROOT, "gpurun_out", "resunit_unfused.pt"))


if __name__ == "__main__":
    main()
