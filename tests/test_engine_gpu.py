"""Single-step and multi-step parity of the CUDA engine against the fp32 oracle (BASELINE.json gates:
per-step noise-prediction cosine >= 0.999 and max-abs <= 2e-2; 20-step DDIM latent PSNR >= 40 dB)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mk(cfg, h, w, rank, images=1, dtype=torch.float16, use_graph=False):
    from edgestyle_b200.engine import DenoiseEngine
    from oracle.step import build_models, synthetic_inputs

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    m = build_models(cfg, (h, w), rank=rank)
    inp = synthetic_inputs(cfg, images, h, w)
    eng = DenoiseEngine(cfg, m.unet.state_dict(), [m.lora_agnostic.state_dict(), m.lora_clothes.state_dict()],
                        m.openpose.state_dict(), m.controlnet.merge_state_dict(), rows=2 * images, h=h, w=w,
                        dtype=dtype, use_graph=use_graph)
    m.unet.to(DEV)
    m.controlnet.to(DEV)
    inp.latents = inp.latents.to(DEV)
    inp.prompt_embeds = inp.prompt_embeds.to(DEV)
    inp.conds = [c.to(DEV) for c in inp.conds]
    eng.set_prompt(inp.prompt_embeds)
    eng.set_conditioning(inp.conds)
    return m, inp, eng


def _metrics(got, want):
    got, want = got.float().flatten(), want.float().flatten()
    cos = torch.nn.functional.cosine_similarity(got, want, dim=0).item()
    return cos, (got - want).abs().max().item()


def _psnr(got, want):
    mse = (got.float() - want.float()).pow(2).mean().item()
    peak = want.abs().max().item()
    return 10 * math.log10(peak * peak / max(mse, 1e-30))


@pytest.mark.parametrize("h,w,images,scale", [(16, 16, 1, [1.0] * 6), (8, 16, 2, [1.0, 0.5, 2.0, 1.0, 0.0, 1.5])])
def test_step_parity_small(h, w, images, scale):
    from oracle.sd15 import SD15Config
    from oracle.step import fused_step

    cfg = SD15Config(block_out_channels=(64, 128, 256, 256), cross_attention_dim=96)
    m, inp, eng = _mk(cfg, h, w, rank=4, images=images)
    x = torch.cat([inp.latents] * 2)
    t = torch.tensor(951, device=DEV)
    want = fused_step(m, x, t, inp.prompt_embeds, scale, inp.conds)
    got = eng.step(x, t, scale)
    cos, mx = _metrics(got, want)
    assert cos >= 0.999 and mx <= 2e-2, (cos, mx)


def test_step_parity_full_size_and_graph():
    """BASELINE config 2 shapes: SD1.5 widths, 64x64 latent, CFG batch 2, rank-32 LoRA, fp16."""
    from oracle.sd15 import SD15Config
    from oracle.step import fused_step

    cfg = SD15Config()
    m, inp, eng = _mk(cfg, 64, 64, rank=32, use_graph=True)
    x = torch.cat([inp.latents] * 2)
    for tval in (951, 501, 1):
        t = torch.tensor(tval, device=DEV)
        want = fused_step(m, x, t, inp.prompt_embeds, inp.conditioning_scale, inp.conds)
        got = eng.step(x, t, inp.conditioning_scale)
        cos, mx = _metrics(got, want)
        print(f"t={tval}: cos={cos:.6f} max_abs={mx:.4g} |eps|max={want.abs().max().item():.3f}")
        assert cos >= 0.999 and mx <= 2e-2, (tval, cos, mx)


def test_ddim20_psnr_full_size():
    from oracle.schedulers import DDIMScheduler
    from oracle.sd15 import SD15Config
    from oracle.step import denoise

    cfg = SD15Config()
    m, inp, eng = _mk(cfg, 64, 64, rank=32, use_graph=True)
    want = denoise(m, inp, 20, 4.5)
    sch = DDIMScheduler()
    ts = sch.set_timesteps(20)
    lat = inp.latents.clone().float()
    for t in ts:
        eng.step(torch.cat([lat] * 2), torch.tensor(float(t), device=DEV), inp.conditioning_scale)
        a_t, a_p = sch.coefficients(int(t))
        eng.cfg_ddim_update(lat, float(a_t), float(a_p), 4.5)
    psnr = _psnr(lat, want)
    print(f"20-step DDIM latent PSNR = {psnr:.2f} dB")
    assert psnr >= 40.0, psnr
