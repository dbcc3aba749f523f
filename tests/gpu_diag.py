"""One-shot GPU diagnostic (not a pytest file): runs every primitive and the DiT on small
cases and prints an error table, so a single gpurun call tells which stage is wrong.
    python tests/gpu_diag.py
"""
import ctypes
import json
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "friendly-stable-audio-tools_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

from helpers import SAO_DIT, build_native_dit, load_golden, rel_l2, max_abs
from stable_audio_tools import _native as nat

results = []


def run(name, fn):
    t0 = time.time()
    try:
        torch.cuda.synchronize()
        r = fn()
        torch.cuda.synchronize()
        results.append((name, "ok", r, time.time() - t0))
    except Exception as e:  # noqa
        results.append((name, "FAIL", repr(e)[:300], time.time() - t0))
        traceback.print_exc()
    print(results[-1], flush=True)


def gemm_case(M, N, K, bf16=0):
    def f():
        torch.manual_seed(1)
        dt = torch.bfloat16 if bf16 else torch.float16
        a = torch.randn(M, K, device="cuda").to(dt)
        w = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
        c = torch.full((M, N), float("nan"), device="cuda")
        nat.check(nat.lib().satb_linear_f32out(nat.ptr(a), nat.ptr(w), nat.ptr(c), M, N, K, bf16, nat.stream_ptr()))
        torch.cuda.synchronize()
        ref = (a.double() @ w.double().T)
        nan = int(torch.isnan(c).sum())
        err = rel_l2(torch.nan_to_num(c), ref)
        # locate error structure
        d = (torch.nan_to_num(c).double() - ref).abs()
        row_bad = int((d.max(dim=1).values > 1e-2).sum())
        col_bad = int((d.max(dim=0).values > 1e-2).sum())
        return dict(rel=err, nan=nan, rows_bad=row_bad, cols_bad=col_bad)
    return f


def attn_case(B, H, Hkv, Nq, Nk):
    def f():
        from oracle.dit_oracle import attention_core
        torch.manual_seed(2)
        q = (torch.randn(B, Nq, H * 64) * 1.5).half()
        k = (torch.randn(B, Nk, Hkv * 64) * 1.5).half()
        v = torch.randn(B, Nk, Hkv * 64).half()
        o = torch.empty(B, Nq, H * 64, dtype=torch.float16, device="cuda")
        qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
        nat.check(nat.lib().satb_attention(nat.ptr(qd), nat.ptr(kd), nat.ptr(vd), nat.ptr(o), B, H, Hkv, Nq, Nk, 0, nat.stream_ptr()))
        heads = lambda t, h: t.float().view(t.shape[0], t.shape[1], h, 64).permute(0, 2, 1, 3)
        ref = attention_core(heads(q, H), heads(k, Hkv), heads(v, Hkv)).permute(0, 2, 1, 3).reshape(B, Nq, H * 64)
        return dict(rel=rel_l2(o.float().cpu(), ref))
    return f


def ln_case():
    x = (torch.randn(1025, 1536) * 3 + 0.5).cuda()
    g = (1 + 0.1 * torch.randn(1536)).cuda()
    out = torch.empty(1025, 1536, dtype=torch.float16, device="cuda")
    nat.check(nat.lib().satb_layernorm(nat.ptr(x), nat.ptr(g), None, nat.ptr(out), 1025, 1536, 0, nat.stream_ptr()))
    ref = torch.nn.functional.layer_norm(x, (1536,), g, None, 1e-5)
    return dict(rel=rel_l2(out.float(), ref))


def snake_case():
    from stable_audio_tools.models.blocks import SnakeBeta
    g = load_golden("snake_beta.npz")
    sn = SnakeBeta(24)
    with torch.no_grad():
        sn.alpha.copy_(torch.from_numpy(g["alpha"]))
        sn.beta.copy_(torch.from_numpy(g["beta"]))
    y = sn.cuda()(torch.from_numpy(g["x"]).cuda()).cpu()
    return dict(maxabs=max_abs(y, torch.from_numpy(g["y"])))


def dit_golden(name, dtype="fp16"):
    def f():
        from oracle import dit_oracle as do
        g = load_golden(name)
        cfg = json.loads(str(g["cfg"]))
        sd = do.make_dit_weights(cfg, seed=int(g["seed"]))
        m = build_native_dit(cfg, sd, operand_dtype=dtype)
        T = lambda k: torch.from_numpy(g[k]).cuda()
        x, t, c, ge, neg = T("x"), T("t"), T("cross"), T("glob"), T("neg")
        out = {}
        y, info = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=1.0, return_info=True)
        out["hidden"] = rel_l2(info["hidden_states"][-1].cpu(), torch.from_numpy(g["hidden_last"]))
        out["nocfg"] = rel_l2(y.cpu(), torch.from_numpy(g["y_nocfg"]))
        out["cfg7"] = rel_l2(m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=7.0).cpu(), torch.from_numpy(g["y_cfg7"]))
        out["phi"] = rel_l2(m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=4.0, scale_phi=0.7).cpu(), torch.from_numpy(g["y_cfg4_phi"]))
        out["neg"] = rel_l2(m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=3.0, negative_cross_attn_cond=neg).cpu(), torch.from_numpy(g["y_neg3"]))
        return out
    return f


def dit_full(depth, B, cfg_scale):
    def f():
        from oracle import dit_oracle as do
        cfg = dict(SAO_DIT, depth=depth)
        sd = do.make_dit_weights(cfg, seed=5)
        torch.manual_seed(1)
        x = torch.randn(B, 64, 1024); t = torch.rand(B) * 0.9 + 0.05
        c = torch.randn(B, 130, 768); c[:, 40:128] = 0.0
        ge = torch.randn(B, 1536)
        t0 = time.time()
        ref = do.dit_forward(sd, cfg, x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=cfg_scale)
        t_cpu = time.time() - t0
        m = build_native_dit(cfg, sd)
        y = m(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(), global_embed=ge.cuda(), cfg_scale=cfg_scale)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = m(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(), global_embed=ge.cuda(), cfg_scale=cfg_scale)
        e1.record(); torch.cuda.synchronize()
        return dict(rel=rel_l2(y.cpu(), ref), cpu_s=t_cpu, gpu_ms=e0.elapsed_time(e1) / 5)
    return f


def oobleck_golden(dtype="fp16"):
    def f():
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import test_gpu_oobleck as tg
        g, ae = tg._build(dtype)
        y = ae.decoder(torch.from_numpy(g["z"]).cuda()).cpu()
        h = ae.encoder(torch.from_numpy(g["a"]).cuda()).cpu()
        return dict(dec=rel_l2(y, torch.from_numpy(g["audio"])), enc=rel_l2(h, torch.from_numpy(g["h"])))
    return f


def oobleck_full(L):
    def f():
        from oracle import oobleck_oracle as oo
        from stable_audio_tools.models.autoencoders import OobleckDecoder
        dcfg = dict(out_channels=2, channels=128, c_mults=[1, 2, 4, 8, 16], strides=[2, 4, 4, 8, 8], latent_dim=64,
                    use_snake=True, final_tanh=False)
        dsd = oo.make_oobleck_weights(oo.decoder_param_shapes(dcfg), seed=9, transposed=oo.decoder_transposed_prefixes(dcfg))
        dec = OobleckDecoder(**dcfg)
        dec.load_state_dict(dsd)
        dec = dec.cuda().eval()
        torch.manual_seed(3)
        z = torch.randn(1, 64, L)
        out = {}
        if L <= 32:
            t0 = time.time()
            ref = oo.oobleck_decoder(z, dsd, dcfg)
            out["cpu_s"] = time.time() - t0
            y = dec(z.cuda())
            out["rel"] = rel_l2(y.cpu(), ref)
        zc = z.cuda()
        y = dec(zc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            y = dec(zc)
        e1.record(); torch.cuda.synchronize()
        out["gpu_ms"] = e0.elapsed_time(e1) / 3
        out["finite"] = bool(torch.isfinite(y).all())
        return out
    return f


def dit_timing(B, depth=24):
    def f():
        from oracle import dit_oracle as do
        cfg = dict(SAO_DIT, depth=depth)
        sd = do.make_dit_weights(cfg, seed=5)
        m = build_native_dit(cfg, sd)
        torch.manual_seed(1)
        x = torch.randn(B, 64, 1024).cuda(); t = (torch.rand(B) * 0.9 + 0.05).cuda()
        c = torch.randn(B, 130, 768).cuda(); ge = torch.randn(B, 1536).cuda()
        for _ in range(3):
            y = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=7.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            y = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=7.0)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        tf = 2.272e12 * 2 * B * depth / 24
        return dict(ms=ms, tflops=tf / ms / 1e9, finite=bool(torch.isfinite(y).all()))
    return f


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), os.cpu_count(), "cpus", flush=True)
    run("snake", snake_case)
    run("layernorm", ln_case)
    for shp in [(128, 256, 64), (128, 64, 64), (128, 128, 64), (128, 256, 128), (128, 256, 1536), (256, 512, 256),
                (1025, 1536, 1536), (8200, 4608, 1536), (333, 128, 768), (129, 384, 200), (2050, 64, 1536)]:
        run(f"gemm{shp}", gemm_case(*shp))
    run("gemm_bf16", gemm_case(1025, 1536, 1536, 1))
    for shp in [(2, 4, 4, 1025, 1025), (1, 24, 12, 1025, 130), (2, 2, 1, 64, 1), (1, 3, 3, 65, 191)]:
        run(f"attn{shp}", attn_case(*shp))
    run("dit_prepend_small", dit_golden("dit_prepend_small.npz"))
    run("dit_adaln_small", dit_golden("dit_adaln_small.npz"))
    run("dit_prepend_small_bf16", dit_golden("dit_prepend_small.npz", "bf16"))
    run("dit_full_d2_cfg", dit_full(2, 1, 7.0))
    run("dit_full_d1_nocfg", dit_full(1, 2, 1.0))
    run("oobleck_small_fp16", oobleck_golden("fp16"))
    run("oobleck_small_bf16", oobleck_golden("bf16"))
    run("oobleck_full_L8", oobleck_full(8))
    run("oobleck_full_L32", oobleck_full(32))
    run("oobleck_full_L1024", oobleck_full(1024))
    run("dit_timing_B1", dit_timing(1))
    run("dit_timing_B4", dit_timing(4))
    print("\n==== SUMMARY ====")
    for r in results:
        print(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w") as f:
        json.dump([[r[0], r[1], str(r[2]), r[3]] for r in results], f, indent=1)
