"""Oracle pinned against outputs of the reference's own (pure-torch) merge code and self-checks
(SURVEY.md 8(c)): golden replay, parameter counts, shape table, algebraic identities."""
import os

import pytest
import torch

from oracle import merge as om
from oracle.controllora import ControlLoRAModel
from oracle.schedulers import DDIMScheduler, UniPCMultistepScheduler, alphas_cumprod
from oracle.sd15 import ControlNetModel, LoRALinearLayer, SD15Config, UNet2DConditionModel, count_params
from oracle.step import build_models, cfg_combine, denoise, fused_step, synthetic_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "merge_block_golden.pt")
TINY = SD15Config(block_out_channels=(32, 64, 64, 64), cross_attention_dim=32)


def test_merge_block_matches_reference_golden():
    cases = torch.load(GOLD)
    assert len(cases) == 3
    for case in cases:
        c, h, w, b = case["shape"]
        blk = om.ControlNetBlock(c, (h, w), 6)
        blk.load_state_dict(case["state_dict"])
        inter = om.interleave_tensors(case["residuals"])
        assert torch.equal(inter, case["interleaved"])
        with torch.no_grad():
            out = blk(inter)
            closed = om.closed_form_block(blk, case["residuals"])
        assert torch.allclose(out, case["out"], rtol=0, atol=1e-6)
        assert torch.allclose(closed, case["out"], rtol=1e-5, atol=2e-6)


def test_parameter_counts_match_published():
    with torch.device("meta"):
        unet = UNet2DConditionModel()
        cn = ControlNetModel()
        lora = ControlLoRAModel(SD15Config(), lora_linear_rank=32)
        multi = om.EdgeStyleMultiControlNetModel([lora] * 6, SD15Config(), (64, 64))
    assert count_params(unet) == 859_520_964
    assert count_params(cn) == 361_279_120
    assert lora.lora_param_count() == 6_354_944
    assert sum(isinstance(m, LoRALinearLayer) for m in lora.modules()) == 82
    assert sum(p.numel() for n, p in multi.named_parameters() if n.startswith("multi_")) == 53_902_720


def test_residual_shape_table():
    # /root/reference/model/edgestyle_onnx_pipeline.py:244-258
    want = [(320, 64, 64)] * 3 + [(320, 32, 32)] + [(640, 32, 32)] * 2 + [(640, 16, 16)] + [(1280, 16, 16)] * 2 \
        + [(1280, 8, 8)] * 3 + [(1280, 8, 8)]
    assert om.residual_shapes(SD15Config(), 64, 64) == want


def test_lora_fuse_equals_unfused_and_state_dict_filter():
    m = build_models(TINY, (8, 8), rank=4)
    inp = synthetic_inputs(TINY, 1, 8, 8)
    x = torch.cat([inp.latents] * 2)
    t = torch.tensor(501)
    d0, m0 = m.lora_agnostic(x, t, inp.prompt_embeds, inp.conds[0], 1.0)
    keys = m.lora_agnostic.state_dict().keys()
    assert all(k.split(".")[0] not in ControlLoRAModel._skip_layers or ".lora_layer." in k for k in keys)
    assert any(".lora_layer.down.weight" in k for k in keys)
    # tied: same storage as the UNet
    assert m.lora_agnostic.down_blocks[0].resnets[0].conv1.weight is m.unet.down_blocks[0].resnets[0].conv1.weight
    m.lora_agnostic.fuse_lora()
    d1, m1 = m.lora_agnostic(x, t, inp.prompt_embeds, inp.conds[0], 1.0)
    for a, b in zip(d0 + [m0], d1 + [m1]):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-4)


def test_zero_zero_convs_make_unet_cond_independent():
    m = build_models(TINY, (8, 8), rank=4)
    for net in (m.lora_agnostic, m.lora_clothes, m.openpose):
        for p in list(net.controlnet_down_blocks.parameters()) + list(net.controlnet_mid_block.parameters()):
            p.zero_()
    inp = synthetic_inputs(TINY, 1, 8, 8)
    x = torch.cat([inp.latents] * 2)
    t = torch.tensor(951)
    e0 = fused_step(m, x, t, inp.prompt_embeds, inp.conditioning_scale, inp.conds)
    e1 = fused_step(m, x, t, inp.prompt_embeds, inp.conditioning_scale, [c * 3 + 1 for c in inp.conds])
    assert torch.equal(e0, e1)


def test_ddim_schedule_constants():
    s = DDIMScheduler()
    ts = s.set_timesteps(20)
    assert ts.tolist() == list(range(951, 0, -50))
    ac = alphas_cumprod()
    assert abs(ac[0].item() - 0.99915) < 1e-6
    assert abs(ac[951].item() - 0.0081550) < 1e-6
    # DDIM with eps == true noise recovers x0 direction: one step from t to t' keeps x0 fixed
    x0 = torch.randn(1, 4, 8, 8)
    eps = torch.randn(1, 4, 8, 8)
    a_t, a_p = s.coefficients(951)
    x_t = a_t.sqrt() * x0 + (1 - a_t).sqrt() * eps
    x_p = s.step(eps, 951, x_t)
    assert torch.allclose(x_p, a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps, atol=1e-4)


def test_unipc_consistency():
    # with the exact eps of a fixed x0, any consistent solver must land on the analytic trajectory
    s = UniPCMultistepScheduler()
    ts = s.set_timesteps(10)
    assert len(ts) == 10 and ts[0] > ts[-1]
    x0 = torch.randn(1, 4, 4, 4)
    noise = torch.randn(1, 4, 4, 4)
    a0, s0 = s._alpha_sigma(s.sigmas[0])
    x = a0 * x0 + s0 * noise
    for i, t in enumerate(ts):
        a, sg = s._alpha_sigma(s.sigmas[i])
        eps = (x - a * x0) / sg
        x = s.step(eps, t, x)
    a, sg = s._alpha_sigma(s.sigmas[-1])
    assert torch.allclose(x, a * x0 + sg * noise, atol=1e-3)


def test_fp64_vs_fp32_oracle_and_cfg():
    m = build_models(TINY, (8, 8), rank=4)
    inp = synthetic_inputs(TINY, 1, 8, 8)
    x = torch.cat([inp.latents] * 2)
    t = torch.tensor(951)
    e32 = fused_step(m, x, t, inp.prompt_embeds, inp.conditioning_scale, inp.conds)
    m.unet.double()
    m.controlnet.double()
    e64 = fused_step(m, x.double(), t, inp.prompt_embeds.double(), inp.conditioning_scale,
                     [c.double() for c in inp.conds])
    assert (e32.double() - e64).abs().max() < 1e-4
    g = cfg_combine(e32, 4.5)
    u, c = e32.chunk(2)
    assert torch.allclose(g, u + 4.5 * (c - u))
    gv = cfg_combine(e32, torch.tensor([3.0]))
    assert torch.allclose(gv, u + 3.0 * (c - u))


def test_denoise_teacher_forced_equals_free_running():
    m = build_models(TINY, (8, 8), rank=4)
    inp = synthetic_inputs(TINY, 1, 8, 8)
    lat, trace = denoise(m, inp, 3, 4.5, return_eps=True)
    lat2 = denoise(m, inp, 3, 4.5, override_latents=[tr[0] for tr in trace])
    assert torch.equal(lat, lat2)


def test_vae_oracle_published_counts_and_identities():
    """oracle/vae.py (parity unpinned): published SD1.5 VAE parameter count and diffusers key names; the deterministic
    algebra the product path relies on (quant_conv folded into conv_out; V bias folded into the attention output bias)
    holds in fp64."""
    from oracle.vae import AutoencoderKL, VaeConfig

    full = AutoencoderKL()
    assert sum(p.numel() for p in full.parameters()) == 83_653_863
    keys = set(full.state_dict())
    for k in ("encoder.down_blocks.0.downsamplers.0.conv.weight", "encoder.mid_block.attentions.0.to_out.0.bias",
              "decoder.up_blocks.2.upsamplers.0.conv.bias", "decoder.up_blocks.3.resnets.0.conv_shortcut.weight",
              "encoder.mid_block.attentions.0.group_norm.weight", "quant_conv.weight", "post_quant_conv.bias"):
        assert k in keys, k
    assert len(keys) == 248
    torch.manual_seed(0)
    m = AutoencoderKL(VaeConfig(block_out_channels=(32, 32, 64, 64))).double().eval()
    x = torch.randn(1, 3, 32, 32, dtype=torch.float64)
    with torch.no_grad():
        d = m.encode(x).latent_dist
        assert d.mean.shape == (1, 4, 4, 4) and torch.equal(d.mode(), d.mean)
        n = torch.randn(1, 4, 4, 4, dtype=torch.float64)
        assert torch.allclose(d.sample(noise=n), d.mean + torch.exp(0.5 * d.logvar) * n)
        # a 1x1 after a conv is a conv
        h = torch.randn(1, 64, 4, 4, dtype=torch.float64)
        wq = m.quant_conv.weight.reshape(8, 8)
        w = torch.einsum("om,mikl->oikl", wq, m.encoder.conv_out.weight)
        b = wq @ m.encoder.conv_out.bias + m.quant_conv.bias
        assert torch.allclose(torch.nn.functional.conv2d(h, w, b, padding=1), m.quant_conv(m.encoder.conv_out(h)), atol=1e-10)
        # rows of softmax sum to 1: the V bias moves into the output bias
        a = m.encoder.mid_block.attentions[0]
        t = a.group_norm(h).view(1, 64, 16).transpose(1, 2)
        p = torch.softmax(a.to_q(t) @ a.to_k(t).transpose(1, 2) * 64 ** -0.5, -1)
        folded = (p @ (t @ a.to_v.weight.t())) @ a.to_out[0].weight.t() + (a.to_out[0].bias + a.to_out[0].weight @ a.to_v.bias)
        assert torch.allclose(h + folded.transpose(1, 2).reshape(1, 64, 4, 4), a(h), atol=1e-10)
        assert m.decode(d.mode()).sample.shape == (1, 3, 32, 32)


def test_vae_conditioning_embedder_matches_reference_golden():
    """`VAEControlNetConditioningEmbedding` + `_tie_weights` executed from the reference's own source text
    (tests/golden/make_golden_vae_cond.py; /root/reference/model/controllora.py:28-56) vs oracle.controllora."""
    import os

    from torch import nn

    from oracle.controllora import VAEControlNetConditioningEmbedding, _tie_weights
    from oracle.vae import AutoencoderKL, VaeConfig

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "vae_cond_golden.pt"))
    vae = AutoencoderKL(VaeConfig(block_out_channels=tuple(g["chans"]), layers_per_block=1)).eval()
    vae.load_state_dict(g["vae"])
    conv_in = nn.Conv2d(4, 16, 3, padding=1)
    emb = VAEControlNetConditioningEmbedding(conv_unet=conv_in, autoencoder=vae)
    assert float(conv_in.weight.detach().abs().max()) == 0.0  # zero_module (:36)
    unet_conv_in = nn.Conv2d(4, 16, 3, padding=1)
    unet_conv_in.load_state_dict(g["conv_in"])
    _tie_weights(unet_conv_in, conv_in)
    assert emb.conv_vae_out.weight is unet_conv_in.weight
    torch.manual_seed(g["seed"])
    with torch.no_grad():
        out = emb(g["image"])
    assert torch.allclose(out, g["out"], atol=1e-6, rtol=1e-6)
    # and piecewise: the draw is `randn(mean.shape)` from the global RNG, scaled by 0.18215, through the tied conv
    torch.manual_seed(g["seed"])
    with torch.no_grad():
        d = vae.encode(g["image"]).latent_dist
        noise = torch.randn(d.mean.shape)
        want = torch.nn.functional.conv2d(d.sample(noise=noise) * 0.18215, unet_conv_in.weight, unet_conv_in.bias, padding=1)
    assert torch.allclose(want, g["out"], atol=1e-6, rtol=1e-6)


def test_multi_controlnet_forward_matches_reference_golden():
    """`EdgeStyleMultiControlNetModel.forward` executed from the reference's own source text on stub nets
    (tests/golden/make_golden_multi_forward.py; edgestyle_multicontrolnet.py:116-171) vs oracle.merge: routing of the
    conditioning images / scales to the nets, level-wise zip, channel interleave, per-level ControlNetBlocks."""
    import os

    from torch import nn

    from oracle.merge import EdgeStyleMultiControlNetModel, closed_form_block, residual_shapes
    from oracle.sd15 import SD15Config

    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "multi_forward_golden.pt"))
    cfg = SD15Config(block_out_channels=(8, 16, 32, 32))
    assert [(c, s, s) for c, s in zip(g["ch"] + [32], g["hw"] + [1])] == residual_shapes(cfg, 8, 8)

    class StubNet(nn.Module):
        def __init__(self, base):
            super().__init__()
            self.base = base

        def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, conditioning_scale, guess_mode=False):
            shift = controlnet_cond.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)
            outs = [(b + shift) * conditioning_scale for b in self.base]
            return outs[:-1], outs[-1]

    multi = EdgeStyleMultiControlNetModel([StubNet(b) for b in g["bases"]], cfg, (8, 8))
    for blk, sd in zip(list(multi.multi_controlnet_down_blocks) + [multi.multi_controlnet_mid_block], g["blocks"]):
        blk.load_state_dict(sd)
    with torch.no_grad():
        down, mid = multi(torch.zeros(2, 4, 8, 8), torch.tensor(1), torch.zeros(2, 77, 8), g["images"], g["scales"])
    assert len(down) == 12
    for a, b in zip(list(down) + [mid], g["down"] + [g["mid"]]):
        assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    # the closed form the CUDA merge kernel implements (no interleaved tensor), per level
    blocks = list(multi.multi_controlnet_down_blocks) + [multi.multi_controlnet_mid_block]
    with torch.no_grad():
        for li, blk in enumerate(blocks):
            res = []
            for k in range(6):
                shift = g["images"][k].mean(dim=(1, 2, 3)).view(-1, 1, 1, 1)
                res.append((g["bases"][k][li] + shift) * g["scales"][k])
            want = (g["down"] + [g["mid"]])[li]
            assert torch.allclose(closed_form_block(blk, res), want, atol=2e-5, rtol=1e-4), li


def test_controlnet_forward_matches_reference_golden():
    """`CachedControlNetModel.forward` executed from the reference's own source text over the oracle's sub-modules
    (tests/golden/make_golden_controlnet_forward.py; controllora.py:59-287) vs oracle.sd15.ControlNetModel.forward:
    timestep forms, embedder skipping, skip order, zero-convs, conditioning_scale, guess_mode gains."""
    import importlib.util
    import os

    from oracle.sd15 import ControlNetModel

    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_golden_controlnet_forward",
                                                  os.path.join(here, "make_golden_controlnet_forward.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = torch.load(os.path.join(here, "controlnet_forward_golden.pt"))
    net = ControlNetModel(SD15Config(**g["cfg"])).eval()
    gen.seeded_fill(net, g["weight_seed"])
    cs = gen.cases()
    assert len(cs) == len(g["outs"]) == 4
    with torch.no_grad():
        for c, want in zip(cs, g["outs"]):
            down, mid = net(c["sample"], c["timestep"], c["ehs"], c["cond"], c["scale"], guess_mode=c["guess_mode"])
            assert len(down) == 12
            for a, b in zip(list(down) + [mid], want["down"] + [want["mid"]]):
                assert a.shape == b.shape and torch.allclose(a, b, atol=1e-6, rtol=1e-5)
    assert float(g["outs"][0]["mid"].abs().max()) > 1e-3          # the fixture is not trivially zero
    assert float(g["outs"][3]["mid"].abs().max()) == 0.0          # conditioning_scale 0


# ------------------------------------------------------------------------------------------------------------------
# Cross-check of the oracle's diffusers internals against an independent port of the same diffusers modules: TVM's
# relax frontend (get_timestep_embedding / Timesteps / TimestepEmbedding / Attention), executed from its source text by
# tests/golden/make_golden_tvm_port.py (the only other implementation of these modules on the box).
# ------------------------------------------------------------------------------------------------------------------
TVM_GOLD = os.path.join(os.path.dirname(__file__), "golden", "tvm_port_golden.pt")


def test_tvm_port_timestep_embedding():
    from oracle.sd15 import TimestepEmbedding, timestep_sinusoid

    g = torch.load(TVM_GOLD)
    ts = g["timesteps"]
    emb = timestep_sinusoid(ts["t"], ts["dim"])  # flip_sin_to_cos=True, freq_shift=0 (SD1.5 config)
    assert emb.shape == ts["emb"].shape
    assert torch.allclose(emb, ts["emb"], rtol=0, atol=2e-5)
    te = g["time_embedding"]
    mod = TimestepEmbedding(te["w1"].shape[1], te["w1"].shape[0])
    mod.load_state_dict({"linear_1.weight": te["w1"], "linear_1.bias": te["b1"], "linear_2.weight": te["w2"],
                         "linear_2.bias": te["b2"]})
    with torch.no_grad():
        y = mod(te["x"])
    assert torch.allclose(y, te["y"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("name", ["self_attention", "cross_attention"])
def test_tvm_port_attention(name):
    """Bias-free to_q/k/v, to_out[0] with bias, scale = dim_head ** -0.5, heads split as [B, N, heads, head_dim]."""
    from oracle.sd15 import Attention

    a = torch.load(TVM_GOLD)[name]
    assert a["qkv_bias"] is False and abs(a["scale"] - 20 ** -0.5) < 1e-12
    ctx = a["encoder_hidden_states"]
    att = Attention(a["to_q"].shape[1], 8, None if ctx is None else ctx.shape[-1])
    sd = {"to_q.weight": a["to_q"], "to_k.weight": a["to_k"], "to_v.weight": a["to_v"],
          "to_out.0.weight": a["to_out_w"], "to_out.0.bias": a["to_out_b"]}
    assert set(att.state_dict()) == set(sd)  # in particular: no q/k/v bias parameters in the oracle either
    att.load_state_dict(sd)
    with torch.no_grad():
        y = att(a["hidden_states"], ctx)
    assert torch.allclose(y, a["y"], rtol=1e-4, atol=1e-5)


def test_tvm_port_golden_is_reproducible():
    """When TVM's sources are on the box, regenerating the vectors from their text gives the committed fixture."""
    import importlib.util

    path = os.path.join(os.path.dirname(__file__), "golden", "make_golden_tvm_port.py")
    spec = importlib.util.spec_from_file_location("make_golden_tvm_port", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(os.path.join(mod.NN, "op.py")):
        pytest.skip("TVM sources not present on this box")
    _, menv = mod._make_env()
    g = torch.load(TVM_GOLD)
    ts = menv["Timesteps"](320, flip_sin_to_cos=True, downscale_freq_shift=0)
    assert torch.equal(ts(mod.Tensor(g["timesteps"]["t"]))._expr, g["timesteps"]["emb"])
