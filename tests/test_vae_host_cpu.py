"""Host logic of edgestyle_b200/vae.py without a GPU: weight packing (tap-major conv matrices, quant_conv folded into
conv_out, V bias folded into the output bias, padded 1x1s), the layer schedule and the scratch-buffer reuse are run
against a torch restatement of the C-ABI semantics (`include/edgestyle_b200.h`) standing in for `ops`, and compared
with the oracle (`oracle/vae.py`).  The kernels themselves are covered by the `-m gpu` tests."""
import types

import pytest
import torch
import torch.nn.functional as F

from oracle.vae import AutoencoderKL as OracleVAE
from oracle.vae import VaeConfig as OracleCfg


def _fake_ops():
    o = types.SimpleNamespace()
    o.calls = []

    def set_gemm_workspace(nbytes=0, device="cpu"):
        return None

    def gemm(a, b, n, *, out, taps=1, whn=None, bias=None, residual=None, alpha=1.0, c1=None, block_n=0, act=0,
             a2=None, b2=None, gn_ws=None, gn_groups=0, rows_per_img=0, **kw):
        assert not kw, kw
        c1 = c1 if c1 is not None else a.shape[1]
        assert b.is_contiguous() and b.shape[1] == taps * c1 and b.shape[0] >= n
        assert c1 % 8 == 0 and a.stride(0) % 8 == 0 and (out.dtype == torch.float32 or out.stride(0) % 8 == 0)
        A = a[:, :c1].float()
        if taps == 9:
            w, h, ni = whn
            assert w * h * ni == a.shape[0] and (w >= 128 or 128 % w == 0)
            x = A.view(ni, h, w, c1).permute(0, 3, 1, 2)
            wt = b[:n].float().view(n, 3, 3, c1).permute(0, 3, 1, 2)
            acc = F.conv2d(x, wt, padding=1).permute(0, 2, 3, 1).reshape(-1, n)
            rpi = w * h
        else:
            acc = A @ b[:n].float().t()
            rpi = rows_per_img
        if a2 is not None:  # second K source (1x1 / centre tap): the resnet shortcut
            assert b2.is_contiguous() and b2.shape == (n, a2.shape[1]) and a2.shape[0] == a.shape[0]
            acc = acc + a2.float() @ b2.float().t()
        if bias is not None:
            acc = acc + bias[:n]
        acc = alpha * acc
        if residual is not None:
            acc = acc + residual[:, :n].float()
        out[:, :n] = acc.to(out.dtype)
        if gn_ws is not None:  # (sum, sumsq) per (image, group) of the OUTPUT, accumulated into a zeroed slot
            assert rpi > 0 and rpi % 32 == 0 and n % gn_groups == 0 and out.stride(0) == n
            assert float(gn_ws.abs().max()) == 0.0, "statistics slot must be zeroed by the caller"
            og = out[:, :n].float().view(-1, rpi, gn_groups, n // gn_groups)
            gn_ws.view(-1, gn_groups, 2).copy_(torch.stack([og.sum(dim=(1, 3)), (og * og).sum(dim=(1, 3))], dim=-1))
        o.calls.append("gemm")
        return out

    def groupnorm(x0, out, gamma, beta, ws, n_img, hw, groups, eps, silu, stats_ready=False, **kw):
        assert not kw
        C = x0.shape[1]
        assert C % 8 == 0 and C % groups == 0 and ws.numel() >= n_img * groups * 2
        x = x0.float().view(n_img, hw, C).permute(0, 2, 1)
        if stats_ready:  # normalise with the statistics the producer left in ws (catches stale / mismatched slots)
            st = ws.view(n_img, groups, 2)
            cnt = hw * (C // groups)
            mean = st[..., 0] / cnt
            var = (st[..., 1] / cnt - mean * mean).clamp_min(0)
            a = torch.rsqrt(var + eps).repeat_interleave(C // groups, dim=1) * gamma
            sh = beta - mean.repeat_interleave(C // groups, dim=1) * a
            y = x * a[:, :, None] + sh[:, :, None]
        else:
            y = F.group_norm(x, groups, gamma, beta, eps)
        if silu:
            y = F.silu(y)
        out.copy_(y.permute(0, 2, 1).reshape(n_img * hw, C).to(out.dtype))
        o.calls.append("groupnorm_apply" if stats_ready else "groupnorm")
        return out

    def nchw_to_nhwc(src, dst):
        n, c, h, w = src.shape
        dst.zero_()
        dst[:, :c] = src.permute(0, 2, 3, 1).reshape(-1, c).to(dst.dtype)
        return dst

    def im2col3x3_pad(src, dst, n, h, w, c, stride, pad_lo, pad_hi):
        x = src[:, :c].float().view(n, h, w, c).permute(0, 3, 1, 2)
        x = F.pad(x, (pad_lo, pad_hi, pad_lo, pad_hi))
        cols = F.unfold(x, 3, stride=stride)                      # [n, c*9, L] with row = ch*9 + tap
        L = cols.shape[-1]
        cols = cols.view(n, c, 9, L).permute(0, 3, 2, 1).reshape(n * L, 9 * c)  # column = tap*c + ch
        assert dst.shape == cols.shape, (dst.shape, cols.shape)
        dst.copy_(cols.to(dst.dtype))
        return dst

    def upsample2x(src, dst, n, h, w):
        c = src.shape[1]
        x = src.view(n, h, w, c).permute(0, 3, 1, 2)
        y = F.interpolate(x.float(), scale_factor=2.0, mode="nearest").to(src.dtype)
        dst.copy_(y.permute(0, 2, 3, 1).reshape(-1, c))
        return dst

    def softmax_rows(s, p, scale=1.0):
        assert s.dtype == torch.float32 and s.shape[1] % 4 == 0
        p.copy_(torch.softmax(s * scale, dim=-1).to(p.dtype))
        return p

    def gaussian_sample(moments, noise, out, scale=1.0):
        n, L = out.shape[:2]
        hw = out[0, 0].numel()
        m = moments[:, :2 * L].view(n, hw, 2 * L).permute(0, 2, 1).reshape(n, 2 * L, *out.shape[2:])
        z = m[:, :L]
        if noise is not None:
            z = z + torch.exp(0.5 * m[:, L:].clamp(-30, 20)) * noise
        out.copy_(z * scale)
        return out

    for f in (set_gemm_workspace, gemm, groupnorm, nchw_to_nhwc, im2col3x3_pad, upsample2x, softmax_rows,
              gaussian_sample):
        setattr(o, f.__name__, f)
    return o


@pytest.fixture()
def vae_pair(monkeypatch):
    from edgestyle_b200 import vae as V

    torch.manual_seed(0)
    cfg = OracleCfg(block_out_channels=(32, 64, 256, 256))
    ref = OracleVAE(cfg).eval()
    with torch.no_grad():  # default init leaves the norms at identity and the biases tiny: perturb them
        for k, p in ref.named_parameters():
            if "norm" in k:
                p.add_(0.2 * torch.randn_like(p))
            elif k.endswith(".bias"):
                p.add_(0.1 * torch.randn_like(p))
    fake = _fake_ops()
    monkeypatch.setattr(V, "ops", fake)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    mine = V.AutoencoderKL(V.VaeConfig(block_out_channels=(32, 64, 256, 256)), ref.state_dict(), dtype=torch.float32,
                           device="cpu")
    return ref, mine, fake


def test_spec_matches_oracle_state_dict():
    from edgestyle_b200.vae import VaeConfig, vae_spec

    ref = OracleVAE(OracleCfg())
    sd = ref.state_dict()
    spec = vae_spec(VaeConfig())
    assert list(spec) == list(spec.keys()) and set(spec) == set(sd)
    for k, shp in spec.items():
        assert tuple(sd[k].shape) == tuple(shp), k
    assert sum(v.numel() for v in sd.values()) == 83_653_863  # published SD1.5 VAE parameter count


def test_encode_schedule_and_packing(vae_pair):
    ref, mine, fake = vae_pair
    x = torch.randn(2, 3, 32, 64)
    noise = torch.randn(2, 4, 4, 8)
    with torch.no_grad():
        want = ref.encode(x).latent_dist
    got = mine.encode(x).latent_dist
    assert torch.allclose(got.mode(), want.mode(), atol=2e-4, rtol=1e-4)
    assert torch.allclose(got.sample(noise=noise, scale=0.18215), want.sample(noise=noise) * 0.18215, atol=2e-4, rtol=1e-4)
    assert torch.allclose(got.logvar, want.logvar, atol=2e-4, rtol=1e-4)
    rep = got.repeat(2).sample(noise=torch.cat([noise, -noise]))
    assert torch.allclose(rep[:2] + rep[2:], 2 * want.mode(), atol=4e-4, rtol=1e-4)
    # a second call reuses the scratch buffers and must give the same answer (no stale state)
    again = mine.encode(x).latent_dist.mode()
    assert torch.equal(again, got.mode())
    # every GroupNorm whose images hold whole 32-row groups took its statistics from the producing GEMM
    assert fake.calls.count("groupnorm_apply") > 0


def test_decode_schedule_and_packing(vae_pair):
    ref, mine, fake = vae_pair
    z = torch.randn(2, 4, 4, 8)
    with torch.no_grad():
        want = ref.decode(z).sample
    got = mine.decode(z).sample
    assert got.shape == want.shape == (2, 3, 32, 64)
    assert torch.allclose(got, want, atol=5e-4, rtol=1e-4), (got - want).abs().max()


def test_no_cuda_no_vae():
    from edgestyle_b200 import vae as V
    from edgestyle_b200.ext import EdgeStyleNativeError

    ref = OracleVAE(OracleCfg(block_out_channels=(32, 32, 32, 32)))
    if not torch.cuda.is_available():
        with pytest.raises(EdgeStyleNativeError):
            V.AutoencoderKL(V.VaeConfig(block_out_channels=(32, 32, 32, 32)), ref.state_dict())
    with pytest.raises(KeyError):
        V.AutoencoderKL(V.VaeConfig(block_out_channels=(32, 32, 32, 32)), {})
