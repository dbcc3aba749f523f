"""GPU parity of the native DiffusionTransformer against (a) golden outputs of the real
reference module and (b) the CPU oracle, through the drop-in module (ctypes -> C ABI).

Tolerances: the native path uses fp16 operands (the reference GPU path's own autocast
dtype, inference/sampling.py:210) with fp32 accumulation and an fp32 residual stream; the
golden / oracle values are fp32 end to end.  Gate: rel-L2 <= 2e-3 on the DiT output for
fp16 operands, <= 1.5e-2 for bf16.  Classifier-free guidance returns
u + (c - u) * s: the difference of two nearly equal forwards is amplified by s relative to the
output, so those cases are gated at tol * max(1, s / 1.5)."""
import json

import pytest
import torch

from helpers import SAO_DIT, build_native_dit, load_golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu

TOL = {"fp16": 2e-3, "bf16": 1.5e-2}


def tol(dtype, cfg_scale=1.0):
    return TOL[dtype] * max(1.0, cfg_scale / 1.5)


def _golden_case(name):
    from oracle import dit_oracle as do
    g = load_golden(name)
    cfg = json.loads(str(g["cfg"]))
    sd = do.make_dit_weights(cfg, seed=int(g["seed"]))
    wsum = float(sum(v.double().abs().sum() for v in sd.values()))
    assert abs(wsum - float(g["wsum"])) <= 1e-6 * abs(wsum), "synthetic weight RNG drifted from the golden run"
    return g, cfg, sd


@pytest.mark.parametrize("name", ["dit_prepend_small.npz", "dit_adaln_small.npz", "dit_patch2_small.npz",
                                  "dit_qknorm_small.npz"])
@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_dit_small_vs_reference_golden(name, dtype):
    g, cfg, sd = _golden_case(name)
    m = build_native_dit(cfg, sd, operand_dtype=dtype)
    T = lambda k: torch.from_numpy(g[k]).cuda()
    x, t, c, ge, neg = T("x"), T("t"), T("cross"), T("glob"), T("neg")
    cases = {
        "y_nocfg": dict(cfg_scale=1.0),
        "y_cfg7": dict(cfg_scale=7.0),
        "y_cfg4_phi": dict(cfg_scale=4.0, scale_phi=0.7),
        "y_neg3": dict(cfg_scale=3.0, negative_cross_attn_cond=neg),
    }
    for key, kw in cases.items():
        y = m(x, t, cross_attn_cond=c, global_embed=ge, **kw).cpu()
        err = rel_l2(y, torch.from_numpy(g[key]))
        assert err < tol(dtype, kw["cfg_scale"]), f"{name} {key} {dtype}: rel l2 {err}"
    y, info = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=1.0, return_info=True)
    hid = info["hidden_states"][-1].cpu()
    err = rel_l2(hid, torch.from_numpy(g["hidden_last"]))
    assert err < TOL[dtype], f"{name} hidden {dtype}: rel l2 {err}"


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_dit_input_concat_and_prepend_cond_vs_reference_golden(dtype):
    """input_concat_cond (half-length: nearest-neighbour resize; 16 extra channels folded through the 1x1 pre-conv into
    project_in) and prepend_cond (3 to_prepend_embed tokens in front of the global token, zeros on the unconditional
    CFG rows) against the real DiffusionTransformer (dit.py:157-173,185-195,281-311)."""
    g, cfg, sd = _golden_case("dit_concat_prepend_small.npz")
    m = build_native_dit(cfg, sd, operand_dtype=dtype)
    T = lambda k: torch.from_numpy(g[k]).cuda()
    x, t = T("x"), T("t")
    kw = dict(cross_attn_cond=T("cross"), global_embed=T("glob"), input_concat_cond=T("concat"))
    full = dict(kw, prepend_cond=T("prepend"), prepend_cond_mask=torch.ones(2, 3, dtype=torch.bool, device="cuda"))
    cases = [("y_nocfg", full, dict(cfg_scale=1.0)), ("y_cfg5", full, dict(cfg_scale=5.0)),
             ("y_cfg3_phi", full, dict(cfg_scale=3.0, scale_phi=0.5)), ("y_concat_only", kw, dict(cfg_scale=4.0)),
             ("y_cfg5", full, dict(cfg_scale=5.0))]          # again after a no-prepend call: P 4 -> 1 -> 4 re-reserves
    for key, cond, extra in cases:
        y = m(x, t, **cond, **extra)
        assert y.shape == x.shape
        err = rel_l2(y.cpu(), torch.from_numpy(g[key]))
        assert err < tol(dtype, extra["cfg_scale"]), f"{key} {dtype}: rel l2 {err}"
    y, info = m(x, t, cfg_scale=1.0, return_info=True, **full)
    assert info["hidden_states"][-1].shape == (2, 200 + 4, 256)
    m.cuda_graph = True                                       # the captured forward carries the prepend tokens too
    y1 = m(x, t, cfg_scale=5.0, **full)
    assert rel_l2(y1.cpu(), torch.from_numpy(g["y_cfg5"])) < tol(dtype, 5.0)
    with pytest.raises(ValueError):
        m(x, t, cross_attn_cond=T("cross"), global_embed=T("glob"))          # the model needs its concat input


def test_dit_prepend_cond_only_guidance_and_patched_concat_vs_oracle():
    """(a) CFG is on when only prepend_cond is given (dit.py:270): no cross-attention rows at all; (b) patch_size 2
    together with input_concat_cond: the concat channels are patched with x ("b c (t p) -> b (c p) t" over all
    io + concat channels, dit.py:206-207) and only io_channels come back."""
    from oracle import dit_oracle as do
    base = dict(io_channels=64, embed_dim=256, depth=2, num_heads=4, cond_token_dim=0, global_cond_dim=256,
                transformer_type="continuous_transformer")
    g = torch.Generator().manual_seed(3)
    x, t = torch.randn(2, 64, 96, generator=g), torch.rand(2, generator=g)
    ge, pc = torch.randn(2, 256, generator=g), torch.randn(2, 5, 32, generator=g)
    cfg = dict(base, prepend_cond_dim=32)
    sd = do.make_dit_weights(cfg, seed=21)
    m = build_native_dit(cfg, sd)
    ref = do.dit_forward(sd, cfg, x, t, global_embed=ge, prepend_cond=pc, cfg_scale=3.0)
    y = m(x.cuda(), t.cuda(), global_embed=ge.cuda(), prepend_cond=pc.cuda(), cfg_scale=3.0).cpu()
    assert rel_l2(y, ref) < tol("fp16", 3.0)
    cfg = dict(base, cond_token_dim=128, project_cond_tokens=False, input_concat_dim=8, patch_size=2)
    sd = do.make_dit_weights(cfg, seed=22)
    m = build_native_dit(cfg, sd)
    c, ic = torch.randn(2, 7, 128, generator=g), torch.randn(2, 8, 96, generator=g)
    for kw in (dict(cfg_scale=1.0), dict(cfg_scale=4.0), dict(cfg_scale=4.0, scale_phi=0.6)):
        ref = do.dit_forward(sd, cfg, x, t, cross_attn_cond=c, global_embed=ge, input_concat_cond=ic, **kw)
        y = m(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(), global_embed=ge.cuda(), input_concat_cond=ic.cuda(), **kw).cpu()
        assert y.shape == x.shape
        assert rel_l2(y, ref) < tol("fp16", kw["cfg_scale"]), kw


def test_dit_repeat_call_is_deterministic_and_cache_safe():
    g, cfg, sd = _golden_case("dit_prepend_small.npz")
    m = build_native_dit(cfg, sd)
    T = lambda k: torch.from_numpy(g[k]).cuda()
    x, t, c, ge = T("x"), T("t"), T("cross"), T("glob")
    y1 = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=7.0)
    y2 = m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=7.0)
    assert torch.equal(y1, y2)
    c2 = c.clone() * 0.5            # new conditioning tensor -> cache must be invalidated
    y3 = m(x, t, cross_attn_cond=c2, global_embed=ge, cfg_scale=7.0)
    assert not torch.equal(y1, y3)
    c2.mul_(2.0)                    # in-place edit bumps the version -> re-prepared
    y4 = m(x, t, cross_attn_cond=c2, global_embed=ge, cfg_scale=7.0)
    assert rel_l2(y4.cpu(), y1.cpu()) < 1e-6


@pytest.mark.parametrize("depth,B,cfg_scale", [(2, 1, 7.0), (1, 2, 1.0), (1, 4, 3.0)])   # last = the bench shape: 8 rows
def test_dit_full_width_vs_oracle(depth, B, cfg_scale):
    """SA-Open-1.0 width (D=1536, 24 heads, 130x768 context, N=1025 tokens) at reduced depth,
    against the CPU oracle computed live (fp32)."""
    from oracle import dit_oracle as do
    cfg = dict(SAO_DIT, depth=depth)
    sd = do.make_dit_weights(cfg, seed=5)
    torch.manual_seed(1)
    x = torch.randn(B, 64, 1024)
    t = torch.rand(B) * 0.9 + 0.05
    c = torch.randn(B, 130, 768)
    c[:, 40:128] = 0.0   # padded T5 rows are exact zeros in the real pipeline (conditioners.py:343-344)
    ge = torch.randn(B, 1536)
    ref = do.dit_forward(sd, cfg, x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=cfg_scale)
    m = build_native_dit(cfg, sd)
    y = m(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(), global_embed=ge.cuda(), cfg_scale=cfg_scale).cpu()
    err = rel_l2(y, ref)
    assert err < tol("fp16", cfg_scale), f"rel l2 {err}"


@pytest.mark.parametrize("B,L,M", [(3, 33, 1), (1, 129, 130), (5, 64, 3)])
def test_dit_ragged_shapes_vs_oracle(B, L, M):
    """Odd batch sizes, sequence lengths that are not tile multiples, 1-token contexts."""
    from oracle import dit_oracle as do
    cfg = dict(io_channels=64, embed_dim=256, depth=2, num_heads=4, cond_token_dim=128, global_cond_dim=256,
               project_cond_tokens=False, transformer_type="continuous_transformer")
    sd = do.make_dit_weights(cfg, seed=7)
    g = torch.Generator().manual_seed(B * 1000 + L)
    x, t = torch.randn(B, 64, L, generator=g), torch.rand(B, generator=g)
    c, ge = torch.randn(B, M, 128, generator=g), torch.randn(B, 256, generator=g)
    m = build_native_dit(cfg, sd)
    for scale in (1.0, 6.0):
        ref = do.dit_forward(sd, cfg, x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=scale)
        y = m(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(), global_embed=ge.cuda(), cfg_scale=scale).cpu()
        assert rel_l2(y, ref) < tol("fp16", scale)


def test_dit_sa2_length_one_block_vs_oracle():
    """SA-2.0 sequence length (6144 latents + prepend = 6145 tokens, BASELINE configs[4]) at full width,
    one block, no CFG: 97 key tiles per attention row."""
    from oracle import dit_oracle as do
    cfg = dict(SAO_DIT, depth=1)
    sd = do.make_dit_weights(cfg, seed=8)
    g = torch.Generator().manual_seed(2)
    x, t = torch.randn(1, 64, 6144, generator=g), torch.tensor([0.3])
    c, ge = torch.randn(1, 130, 768, generator=g), torch.randn(1, 1536, generator=g)
    ref = do.dit_forward(sd, cfg, x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=1.0)
    y = build_native_dit(cfg, sd)(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(), global_embed=ge.cuda(), cfg_scale=1.0).cpu()
    assert rel_l2(y, ref) < 2e-3


def test_dit_full_width_bf16_vs_oracle():
    from oracle import dit_oracle as do
    cfg = dict(SAO_DIT, depth=1)
    sd = do.make_dit_weights(cfg, seed=9)
    g = torch.Generator().manual_seed(3)
    x, t = torch.randn(1, 64, 1024, generator=g), torch.tensor([0.6])
    c, ge = torch.randn(1, 130, 768, generator=g), torch.randn(1, 1536, generator=g)
    ref = do.dit_forward(sd, cfg, x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=1.0)
    y = build_native_dit(cfg, sd, operand_dtype="bf16")(x.cuda(), t.cuda(), cross_attn_cond=c.cuda(),
                                                        global_embed=ge.cuda(), cfg_scale=1.0).cpu()
    assert rel_l2(y, ref) < 1.5e-2



def test_dit_full_size_cfg_is_linear_in_the_scale_and_deterministic():
    """BASELINE-size property test (SA-Open-1.0, all 24 blocks, 1025 tokens; no oracle at this size):
    out(s) = u + (c - u) s is affine in the guidance scale, so out(5) - out(1) == 2 (out(3) - out(1)), the
    no-CFG output equals out(1), and repeated calls are bit-identical."""
    from oracle import dit_oracle as do
    sd = do.make_dit_weights(SAO_DIT, seed=10)
    m = build_native_dit(SAO_DIT, sd)
    g = torch.Generator().manual_seed(4)
    x, t = torch.randn(1, 64, 1024, generator=g).cuda(), torch.tensor([0.45]).cuda()
    c, ge = torch.randn(1, 130, 768, generator=g).cuda(), torch.randn(1, 1536, generator=g).cuda()
    run = lambda s: m(x, t, cross_attn_cond=c, global_embed=ge, cfg_scale=s).float()
    y1, y3, y5, y3b = run(1.0), run(3.0), run(5.0), run(3.0)
    assert torch.isfinite(y5).all()
    assert torch.equal(y3, y3b)
    d31, d51 = y3 - y1, y5 - y1
    assert rel_l2(d51, 2.0 * d31) < 1e-3
    assert float(d31.norm()) > 1e-3 * float(y1.norm())          # guidance really changes the output


def test_dit_cuda_graph_call_equals_the_eager_call():
    """cuda_graph = True: a denoiser call replayed from a captured CUDA graph gives the bits of the
    eager call, for new inputs (replay), new conditioning (re-capture) and after an eager call in between."""
    g, cfg, sd = _golden_case("dit_prepend_small.npz")
    m = build_native_dit(cfg, sd)
    T = lambda k: torch.from_numpy(g[k]).cuda()
    x, t, c, ge = T("x"), T("t"), T("cross"), T("glob")
    eager = lambda xx, tt, cc: m(xx, tt, cross_attn_cond=cc, global_embed=ge, cfg_scale=7.0).clone()
    from stable_audio_tools import _native
    y0 = eager(x, t, c)
    m.cuda_graph = True
    n0 = _native.launch_count()
    y1 = eager(x, t, c)                                   # capture + replay
    n_capture = _native.launch_count() - n0
    assert torch.equal(y0, y1)
    n0 = _native.launch_count()
    x2, t2 = x * 0.5 + 0.1, (t * 0.7).contiguous()
    y2 = eager(x2, t2, c)                                 # replay with new inputs
    assert _native.launch_count() - n0 > 10               # the replayed launches are accounted for
    c2 = (c * 0.9).contiguous()
    y3 = eager(x2, t2, c2)                                # new conditioning: prepared again, captured again
    m.cuda_graph = False
    assert torch.equal(y2, eager(x2, t2, c)) and torch.equal(y3, eager(x2, t2, c2))
    m.cuda_graph = True
    assert torch.equal(y3, eager(x2, t2, c2))             # eager call in between dropped the graph: captured again
    assert n_capture > 0
