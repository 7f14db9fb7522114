"""SURVEY.md 8(f) row N2 on the GPU: the VAE encoder (VAEControlNetConditioningEmbedding, controllora.py:38-42) and
decoder (edgestyle_pipeline.py:552-557) on the sm_100a kernels against the fp32 oracle (`oracle/vae.py`), plus the two
small kernels they add.  Tolerance (fp16 storage, fp32 accumulation): cosine >= 0.999 and max-abs error <= 2e-2 of the
reference's largest magnitude, as for the per-step noise prediction."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _close(got, want, tol=2e-2, cos_min=0.999):
    got, want = got.float().cpu().flatten(), want.float().cpu().flatten()
    cos = F.cosine_similarity(got, want, dim=0).item()
    err = (got - want).abs().max().item()
    ok = cos >= cos_min and err <= tol * max(1.0, want.abs().max().item())
    return ok, f"cos {cos:.6f} max-abs {err:.4g} (ref max {want.abs().max().item():.4g})"


def _oracle(block_out_channels, seed=0):
    from oracle.vae import AutoencoderKL, VaeConfig

    torch.manual_seed(seed)
    ref = AutoencoderKL(VaeConfig(block_out_channels=tuple(block_out_channels))).eval()
    with torch.no_grad():  # move the norms off identity and the biases off ~0 so that every parameter matters
        for k, p in ref.named_parameters():
            if "norm" in k:
                p.add_(0.2 * torch.randn_like(p))
            elif k.endswith(".bias"):
                p.add_(0.1 * torch.randn_like(p))
    return ref


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("rows,cols", [(5, 64), (130, 4096), (3, 12288)])
def test_softmax_rows(dtype, rows, cols):
    from edgestyle_b200 import ops

    g = torch.Generator().manual_seed(rows + cols)
    s = (torch.randn(rows, cols, generator=g) * 6).to(DEV)
    s[0, :8] = 40.0  # a dominant block: checks the max subtraction
    p = torch.full((rows, cols), 7.0, device=DEV, dtype=dtype)
    ops.softmax_rows(s, p, 0.5)
    want = torch.softmax(s * 0.5, dim=-1)
    assert (p.float() - want).abs().max().item() <= (4e-3 if dtype == torch.bfloat16 else 6e-4)
    assert (p.float().sum(-1) - 1).abs().max().item() <= (2e-2 if dtype == torch.bfloat16 else 3e-3)
    # pitched views: scores and probabilities as column slices of wider buffers
    big_s = torch.randn(rows, cols + 8, generator=g).to(DEV)
    big_p = torch.zeros(rows, cols + 16, device=DEV, dtype=dtype)
    ops.softmax_rows(big_s[:, 4:4 + cols], big_p[:, 8:8 + cols])
    assert (big_p[:, 8:8 + cols].float() - torch.softmax(big_s[:, 4:4 + cols], -1)).abs().max().item() <= 4e-3
    assert big_p[:, :8].abs().max().item() == 0 and big_p[:, 8 + cols:].abs().max().item() == 0


def test_gaussian_sample_and_padded_im2col():
    from edgestyle_b200 import ops

    g = torch.Generator().manual_seed(3)
    n, L, h, w = 3, 4, 5, 7
    mom = torch.randn(n * h * w, 16, generator=g).to(DEV)
    mom[:, 4:8] *= 20  # exercises both clamps of logvar (-30, 20)
    noise = torch.randn(n, L, h, w, generator=g).to(DEV)
    out = torch.empty(n, L, h, w, device=DEV)
    ops.gaussian_sample(mom, noise, out, 0.18215)
    m = mom.view(n, h, w, 16).permute(0, 3, 1, 2)
    want = (m[:, :4] + torch.exp(0.5 * m[:, 4:8].clamp(-30, 20)) * noise) * 0.18215
    assert torch.allclose(out, want, rtol=1e-4, atol=1e-5)
    ops.gaussian_sample(mom, None, out)
    assert torch.equal(out, m[:, :4].contiguous())
    # Downsample2D(padding=0): zero pad right/bottom only, stride 2 -- vector (c % 8 == 0) and scalar channel counts
    for c, ld in ((16, 16), (3, 8)):
        n, h, w = 2, 8, 12
        x = torch.randn(n * h * w, ld, generator=g).to(DEV).half()
        ldo = 9 * c if c % 8 == 0 else 32
        col = torch.full((n * (h // 2) * (w // 2), ldo), 5.0, device=DEV, dtype=torch.float16)
        ops.im2col3x3_pad(x, col, n, h, w, c, 2, 0, 1)
        xi = F.pad(x[:, :c].float().view(n, h, w, c).permute(0, 3, 1, 2), (0, 1, 0, 1))
        u = F.unfold(xi, 3, stride=2)
        want = u.view(n, c, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * c)
        assert torch.equal(col[:, :9 * c].float(), want)
        assert col[:, 9 * c:].abs().sum().item() == 0
    # pad (1, 1) reproduces es_im2col3x3
    x = torch.randn(2 * 8 * 8, 16, generator=g).to(DEV).half()
    a = torch.empty(2 * 4 * 4, 144, device=DEV, dtype=torch.float16)
    b = torch.empty_like(a)
    ops.im2col3x3(x, a, 2, 8, 8, 16, 2)
    ops.im2col3x3_pad(x, b, 2, 8, 8, 16, 2, 1, 1)
    assert torch.equal(a, b)


@pytest.mark.parametrize("chans,n,H,W,dtype", [((32, 64, 128, 128), 2, 64, 128, torch.float16),
                                               ((32, 64, 128, 128), 1, 64, 64, torch.bfloat16),
                                               ((128, 256, 512, 512), 1, 256, 256, torch.float16)])
def test_vae_encode_decode_parity(chans, n, H, W, dtype):
    """Small widths with a non-square image, and the real SD1.5 widths at 256x256 (1024 attention tokens of width 512)."""
    from edgestyle_b200.vae import AutoencoderKL, VaeConfig

    ref = _oracle(chans)
    vae = AutoencoderKL(VaeConfig(block_out_channels=chans), ref.state_dict(), dtype=dtype)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(n, 3, H, W, generator=g) * 2 - 1
    noise = torch.randn(n, 4, H // 8, W // 8, generator=g)
    tol = 2e-2 if dtype == torch.float16 else 8e-2
    cos_min = 0.999 if dtype == torch.float16 else 0.998
    with torch.no_grad():
        wd = ref.encode(x).latent_dist
        z = wd.sample(noise=noise)
        wimg = ref.decode(z).sample
    gd = vae.encode(x.to(DEV)).latent_dist
    for name, got, want in (("mode", gd.mode(), wd.mode()), ("sample", gd.sample(noise=noise.to(DEV)), z),
                            ("logvar", gd.logvar, wd.logvar)):
        ok, msg = _close(got, want, tol, cos_min)
        assert ok, f"encode {name}: {msg}"
    gimg = vae.decode(z.to(DEV)).sample
    assert gimg.shape == wimg.shape
    ok, msg = _close(gimg, wimg, tol, cos_min)
    assert ok, f"decode: {msg}"
    # generator-driven sampling stays inside the distribution: (sample - mean) / std is the drawn noise
    s1 = gd.sample(generator=torch.Generator(DEV).manual_seed(5))
    s2 = gd.sample(generator=torch.Generator(DEV).manual_seed(5))
    assert torch.equal(s1, s2) and not torch.equal(s1, gd.mode())
    # round trip through the checkpoint format
    import tempfile

    with tempfile.TemporaryDirectory() as d:
        vae.save_pretrained(d)
        again = AutoencoderKL.from_pretrained(d, torch_dtype=dtype)
    # same weights, same schedule; not bit-identical run to run (GroupNorm statistics are accumulated with fp32 atomics)
    assert (again.decode(z.to(DEV)).sample - gimg).abs().max().item() <= tol * max(1.0, gimg.abs().max().item())


def test_controllora_vae_conditioning_and_pipeline_decode():
    """preprocess_image of a ControlLoRA net (VAE encode -> sample * 0.18215 -> conv_vae_out == the tied UNet conv_in,
    controllora.py:36-42, 624) with handed-in noise, and the pipeline's output_type="pt" / "np" tail (VAE decode +
    denormalise, edgestyle_pipeline.py:552-572) against the oracle."""
    from edgestyle_b200 import config as C
    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      EdgeStyleStableDiffusionControlNetPipeline, UNet2DConditionModel)
    from edgestyle_b200.vae import AutoencoderKL, VaeConfig
    from oracle.sd15 import SD15Config
    from oracle.step import build_models, synthetic_inputs

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ocfg = SD15Config(block_out_channels=(64, 128, 256, 256), cross_attention_dim=96)
    h = w = 16
    m = build_models(ocfg, (h, w), rank=4)
    inp = synthetic_inputs(ocfg, 1, h, w)
    cfg = C.UNetConfig.from_any(ocfg)
    chans = (32, 64, 128, 128)
    rvae = _oracle(chans, seed=1)
    vae = AutoencoderKL(VaeConfig(block_out_channels=chans), rvae.state_dict())
    unet = UNet2DConditionModel(cfg, m.unet.state_dict())
    agn = ControlLoRAModel(cfg, m.lora_agnostic.state_dict(), 4, unet=unet)
    clo = ControlLoRAModel(cfg, m.lora_clothes.state_dict(), 4, unet=unet)
    pose = CachedControlNetModel(cfg, m.openpose.state_dict())
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], m.controlnet.merge_state_dict(), (h, w))
    with pytest.raises(RuntimeError):
        agn.preprocess_image(torch.zeros(1, 3, 8 * h, 8 * w, device=DEV))  # no autoencoder yet
    agn.set_autoencoder(vae)
    clo.set_autoencoder(vae)
    g = torch.Generator().manual_seed(9)
    img = torch.rand(1, 3, 8 * h, 8 * w, generator=g) * 2 - 1
    noise = torch.randn(2, 4, h, w, generator=g)
    with torch.no_grad():
        d = rvae.encode(torch.cat([img] * 2)).latent_dist
        want = F.conv2d(d.sample(noise=noise) * 0.18215, m.unet.conv_in.weight.float().cpu(),
                        m.unet.conv_in.bias.float().cpu(), padding=1)
    got = agn.preprocess_image(img.to(DEV), repeats=2, noise=noise.to(DEV))
    assert got.shape == want.shape == (2, 64, h, w)
    ok, msg = _close(got, want)
    assert ok, f"VAE conditioning embedding: {msg}"
    assert not torch.equal(got[0], got[1])  # the two CFG rows are independent samples (edgestyle_pipeline.py:657-662)
    # pipeline: raw images for all six nets, decoded output
    pipe = EdgeStyleStableDiffusionControlNetPipeline(vae=vae, unet=unet, controlnet=multi, use_graph=False)
    raw = [torch.rand(1, 3, 8 * h, 8 * w, generator=g) * 2 - 1 for _ in range(6)]
    kw = dict(image=raw, prompt_embeds=inp.prompt_embeds[1:], negative_prompt_embeds=inp.prompt_embeds[:1],
              latents=inp.latents, num_inference_steps=2, guidance_scale=4.5)
    torch.manual_seed(123)
    lat = pipe(output_type="latent", **kw).images
    torch.manual_seed(123)
    out = pipe(output_type="pt", **kw)
    assert out.nsfw_content_detected is None and out.images.shape == (1, 3, 8 * h, 8 * w)
    with torch.no_grad():
        wimg = (rvae.decode(lat.float().cpu() / 0.18215).sample / 2 + 0.5).clamp(0, 1)
    ok, msg = _close(out.images, wimg)
    assert ok, f"pipeline decode: {msg}"
    torch.manual_seed(123)
    arr = pipe(output_type="np", **kw).images
    assert arr.shape == (1, 8 * h, 8 * w, 3) and arr.dtype.name == "float32"
    assert abs(arr - out.images.cpu().permute(0, 2, 3, 1).numpy()).max() <= 2e-2  # two runs: fp32-atomic statistics
    with pytest.raises(ValueError):
        EdgeStyleStableDiffusionControlNetPipeline(unet=unet, controlnet=multi)(output_type="pt", **kw)
