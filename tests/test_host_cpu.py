"""Host-side logic that needs no GPU: parameter specs vs the oracle, weight packing, schedulers, the
reference-surface error behaviour of the model mirrors."""
import pytest
import torch

from edgestyle_b200 import config as C
from edgestyle_b200.schedulers import DDIMScheduler
from oracle.schedulers import DDIMScheduler as OracleDDIM
from oracle.sd15 import SD15Config
from oracle.step import build_models

TINY = SD15Config(block_out_channels=(64, 128, 128, 128), cross_attention_dim=64)


@pytest.fixture(scope="module")
def models():
    return build_models(TINY, (8, 8), rank=4)


def test_specs_match_oracle_state_dicts(models):
    cfg = C.UNetConfig.from_any(TINY)
    usd = models.unet.state_dict()
    spec = C.unet_spec(cfg)
    assert set(spec) == set(usd) and all(tuple(usd[k].shape) == v for k, v in spec.items())
    lsd = models.lora_agnostic.state_dict()
    ls = dict(C.lora_spec(cfg, 4))
    ls.update(C.controlnet_extra_spec(cfg, False))
    assert set(ls) <= set(lsd) and all(tuple(lsd[k].shape) == v for k, v in ls.items())
    assert all(k.startswith("controlnet_cond_embedding.") for k in set(lsd) - set(ls))
    psd = models.openpose.state_dict()
    ps = dict(C.encoder_spec(cfg))
    ps.update(C.controlnet_extra_spec(cfg, True))
    assert set(ps) == set(psd)
    msd = models.controlnet.merge_state_dict()
    ms = C.merge_spec(cfg, 8, 8)
    assert set(ms) == set(msd) and all(tuple(msd[k].shape) == v for k, v in ms.items())
    assert len(C.lora_linear_names(C.UNetConfig())) == 82  # the reference's 82 LoRA'd Linear layers


def test_full_size_spec_parameter_counts():
    cfg = C.UNetConfig()
    n = lambda spec: sum(int(torch.tensor(s).prod()) for s in spec.values())
    assert n(C.unet_spec(cfg)) == 859_520_964
    cn = dict(C.encoder_spec(cfg))
    cn.update(C.controlnet_extra_spec(cfg, True))
    assert n(cn) == 361_279_120
    assert n(C.lora_spec(cfg, 32)) == 6_354_944
    assert n(C.merge_spec(cfg, 64, 64)) == 53_902_720
    assert C.residual_shapes(cfg, 64, 64)[3] == (320, 32, 32) and C.residual_shapes(cfg, 96, 128)[-1] == (1280, 12, 16)


def test_ddim_host_tables_match_oracle():
    a, b = DDIMScheduler(), OracleDDIM()
    for n in (20, 25, 50):
        ta, tb = a.set_timesteps(n), b.set_timesteps(n)
        assert ta.tolist() == tb.tolist()
        for t in ta:
            ca, cb = a.coefficients(int(t)), b.coefficients(int(t))
            assert abs(ca[0] - float(cb[0])) < 1e-7 and abs(ca[1] - float(cb[1])) < 1e-7


def test_merge_block_repack_is_a_pure_permutation(models):
    from edgestyle_b200.engine import pack_merge_block

    blk = models.controlnet.multi_controlnet_down_blocks[4]
    sd = blk.state_dict()
    c = blk.third_conv.weight.shape[0]
    h, w = blk.second_normalization.weight.shape[1:]
    p = pack_merge_block(sd, c, h, w, torch.float32, "cpu")
    g1 = sd["first_normalization.weight"]  # [3C, H, W], channel index c*3 + pair
    assert torch.equal(p["g1"][5, 2, 7], g1[7 * 3 + 2].reshape(-1)[5])
    assert torch.equal(p["w1"][1, :, 3], sd["first_conv.weight"][3 * 3 + 1, :, 0, 0])  # [pair][net][channel]
    assert torch.equal(p["w2"][2, 3], sd["second_conv.weight"][3, 2, 0, 0])
    assert torch.equal(p["g2"][6, 9], sd["second_normalization.weight"][9].reshape(-1)[6])


def test_model_mirrors_reference_surface(models):
    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      UNet2DConditionModel)

    cfg = C.UNetConfig.from_any(TINY)
    unet = UNet2DConditionModel(cfg, models.unet.state_dict())
    agn = ControlLoRAModel(cfg, models.lora_agnostic.state_dict(), lora_linear_rank=4)
    clo = ControlLoRAModel(cfg, models.lora_clothes.state_dict(), lora_linear_rank=4)
    pose = CachedControlNetModel(cfg, models.openpose.state_dict())
    # state_dict filter semantics of controllora.py:600-606
    assert all(k.split(".")[0] not in ControlLoRAModel._skip_layers or ".lora_layer." in k for k in agn.state_dict())
    with pytest.raises(NotImplementedError):
        EdgeStyleMultiControlNetModel([agn, pose, clo, pose, agn, pose], latent_hw=(8, 8))  # nets 2/4 must be one object
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], models.controlnet.merge_state_dict(), (8, 8))
    with pytest.raises(RuntimeError):
        multi.unet()  # not tied yet (app.py:95-97)
    agn.tie_weights(unet)
    clo.tie_weights(unet)
    assert multi.unet() is unet
    with pytest.raises(KeyError):  # a conv-LoRA net needs the LoRAConv2dLayer tensors of every convolution
        ControlLoRAModel(cfg, models.lora_agnostic.state_dict(), lora_linear_rank=4, lora_conv2d_rank=4)
    fresh = ControlLoRAModel.from_unet(unet, lora_linear_rank=4)
    assert all(v.abs().max() == 0 for k, v in fresh.state_dict().items() if k.endswith("up.weight"))
    with pytest.raises(KeyError):
        CachedControlNetModel(cfg, {})
    x = torch.zeros(2, 4, 8, 8)
    with pytest.raises(ValueError):
        multi.forward(x, 1, torch.zeros(2, 77, 64), [x] * 5, [1.0] * 5)
    # no CUDA here: the engine must refuse to run rather than fall back
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            multi.forward(x, 1, torch.zeros(2, 77, 64), [torch.zeros(2, 64, 8, 8)] * 6, [1.0] * 6)


def test_pipeline_rejects_out_of_scope_arguments(models):
    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      EdgeStyleStableDiffusionControlNetPipeline, UNet2DConditionModel)

    cfg = C.UNetConfig.from_any(TINY)
    unet = UNet2DConditionModel(cfg, models.unet.state_dict())
    agn = ControlLoRAModel(cfg, models.lora_agnostic.state_dict(), 4, unet=unet)
    clo = ControlLoRAModel(cfg, models.lora_clothes.state_dict(), 4, unet=unet)
    pose = CachedControlNetModel(cfg, models.openpose.state_dict())
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], models.controlnet.merge_state_dict(), (8, 8))
    pipe = EdgeStyleStableDiffusionControlNetPipeline(unet=unet, controlnet=multi)
    pe = torch.zeros(1, 77, 64)
    conds = [torch.zeros(2, 64, 8, 8)] * 6
    with pytest.raises(NotImplementedError):
        pipe(prompt="a photo", image=conds, prompt_embeds=pe, negative_prompt_embeds=pe, output_type="latent")
    with pytest.raises(ValueError):  # decoded output needs the pipeline's vae (edgestyle_pipeline.py:552-557)
        pipe(image=conds, prompt_embeds=pe, negative_prompt_embeds=pe, output_type="pil")
    with pytest.raises(ValueError):
        pipe(image=conds, prompt_embeds=pe, negative_prompt_embeds=pe, output_type="jpeg")
    with pytest.raises(ValueError, match="does not support custom"):  # retrieve_timesteps (edgestyle_pipeline.py:698-706)
        pipe(image=conds, prompt_embeds=pe, negative_prompt_embeds=pe, output_type="latent", timesteps=[900, 500, 100])
    with pytest.raises(ValueError, match="list of generators"):      # prepare_latents (:613-617)
        pipe(image=conds, prompt_embeds=pe, negative_prompt_embeds=pe, output_type="latent", num_images_per_prompt=2,
             generator=[torch.Generator().manual_seed(0)])
    with pytest.raises(NotImplementedError):
        pipe(image=conds, prompt_embeds=pe, negative_prompt_embeds=pe, output_type="latent", clip_skip=1)
    with pytest.raises(ValueError):
        pipe(image=conds[:4], prompt_embeds=pe, negative_prompt_embeds=pe, output_type="latent")


def test_checkpoint_directory_roundtrip(models, tmp_path):
    """N1: the reference's on-disk format -- merge blocks at the top level, one controlnet_{idx}/ per distinct
    ControlLoRA with ONLY LoRA + non-tied tensors, shared objects restored from load_pattern."""
    import os

    from safetensors.torch import load_file

    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      UNet2DConditionModel)

    cfg = C.UNetConfig.from_any(TINY)
    unet = UNet2DConditionModel(cfg, models.unet.state_dict())
    agn = ControlLoRAModel(cfg, models.lora_agnostic.state_dict(), 4, unet=unet)
    clo = ControlLoRAModel(cfg, models.lora_clothes.state_dict(), 4, unet=unet)
    pose = CachedControlNetModel(cfg, models.openpose.state_dict())
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], models.controlnet.merge_state_dict(), (8, 8))
    d = str(tmp_path / "edgestyle")
    pattern = [0, None, 1, None, 1, None]
    multi.save_pretrained(d, save_pattern=pattern)
    unet.save_pretrained(str(tmp_path / "unet"))
    pose.save_pretrained(str(tmp_path / "openpose"))
    assert sorted(os.listdir(d)) == ["controlnet_0", "controlnet_1", "diffusion_pytorch_model.safetensors"]
    top = load_file(os.path.join(d, "diffusion_pytorch_model.safetensors"))
    assert all(k.startswith("multi_controlnet_") for k in top)
    sub = load_file(os.path.join(d, "controlnet_1", "diffusion_pytorch_model.safetensors"))
    assert all(k.split(".")[0] not in ControlLoRAModel._skip_layers or ".lora_layer." in k for k in sub)
    unet2 = UNet2DConditionModel.from_pretrained(str(tmp_path / "unet"))
    pose2 = CachedControlNetModel.from_pretrained(str(tmp_path / "openpose"))
    multi2 = EdgeStyleMultiControlNetModel.from_pretrained(d, load_pattern=pattern, controlnet_class=ControlLoRAModel,
                                                           static_controlnets=[None, pose2, None, pose2, None, pose2],
                                                           latent_hw=(8, 8))
    assert multi2.nets[2] is multi2.nets[4] and multi2.nets[0] is not multi2.nets[2]  # train_...py:849-856
    for n in (multi2.nets[0], multi2.nets[2]):
        n.tie_weights(unet2)
    assert multi2.unet() is unet2 and unet2.config == cfg
    for a, b in ((multi.state_dict(), multi2.state_dict()), (agn.state_dict(), multi2.nets[0].state_dict()),
                 (clo.state_dict(), multi2.nets[2].state_dict()), (unet.state_dict(), unet2.state_dict())):
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    with pytest.raises(ValueError):
        EdgeStyleMultiControlNetModel.from_pretrained(d, controlnet_class=ControlLoRAModel)  # load_pattern required
    with pytest.raises(ValueError):
        EdgeStyleMultiControlNetModel.from_pretrained(d, load_pattern=pattern, controlnet_class=ControlLoRAModel)


def test_fuse_matches_oracle_fuse_lora(models, tmp_path):
    """ControlLoRAModel.fuse() / fused_state_dict() (controllora.py:728-777) against the oracle's fuse_lora, the
    FusedControlLoRAModel checkpoint round trip, and multi.fuse() pinning the engine to fused weight copies."""
    import copy

    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      FusedControlLoRAModel, UNet2DConditionModel)

    cfg = C.UNetConfig.from_any(TINY)
    unet = UNet2DConditionModel(cfg, models.unet.state_dict())
    agn = ControlLoRAModel(cfg, models.lora_agnostic.state_dict(), lora_linear_rank=4)
    with pytest.raises(RuntimeError):
        agn.fused_state_dict()  # not tied yet
    agn.tie_weights(unet)
    before = {k: v.clone() for k, v in unet.state_dict().items()}
    fused = agn.fuse()
    assert isinstance(fused, FusedControlLoRAModel) and isinstance(fused, CachedControlNetModel) and not fused.uses_lora
    ref = copy.deepcopy(models.lora_agnostic)
    ref.fuse_lora()
    want = ref.full_state_dict()
    got = fused.state_dict()
    changed = 0
    for k, v in want.items():
        if k.startswith("controlnet_cond_embedding."):
            continue
        assert k in got, k
        assert torch.allclose(got[k].float(), v.float(), atol=1e-6), k
        changed += int(k in before and not torch.equal(got[k], before[k]))
    assert changed > 0, "the LoRA update must change at least one tied weight"
    assert all(".lora_layer." not in k for k in got)
    assert all(torch.equal(v, before[k]) for k, v in unet.state_dict().items())  # the UNet is left untouched
    half = agn.fused_state_dict(0.5)
    k = next(k for k in got if k in before and not torch.equal(got[k], before[k]))
    assert torch.allclose(half[k].float() - before[k].float(), 0.5 * (got[k].float() - before[k].float()), atol=1e-6)
    fused.save_pretrained(tmp_path / "fused")
    again = FusedControlLoRAModel.from_pretrained(tmp_path / "fused")
    assert all(torch.equal(again.state_dict()[k], v) for k, v in got.items())
    with pytest.raises(NotImplementedError):
        agn.fuse_lora()
    clo = ControlLoRAModel(cfg, models.lora_clothes.state_dict(), lora_linear_rank=4, unet=unet)
    pose = CachedControlNetModel(cfg, models.openpose.state_dict())
    multi = EdgeStyleMultiControlNetModel([agn, pose, clo, pose, clo, pose], models.controlnet.merge_state_dict(), (8, 8))
    multi.fuse()
    assert multi._fused and multi.nets[0] is agn  # nets stay ControlLoRA objects: conditioning stays cacheable


def test_checkpoint_directory_matches_reference_save_pretrained(models, tmp_path):
    """N1 against the reference's own `save_pretrained` / `state_dict` (edgestyle_multicontrolnet.py:173-282, executed
    from its source text on stub nets by tests/golden/make_golden_checkpoint_dir.py): the mirror reads the file the
    reference wrote, and writes the same file name, keys, tensors and `controlnet_{idx}` sub-directories."""
    import json
    import os
    import shutil

    from safetensors.torch import load_file

    from edgestyle_b200.model import (CachedControlNetModel, ControlLoRAModel, EdgeStyleMultiControlNetModel,
                                      UNet2DConditionModel)

    here = os.path.join(os.path.dirname(__file__), "golden")
    meta = json.load(open(os.path.join(here, "checkpoint_dir_golden.json")))
    golden = load_file(os.path.join(here, "checkpoint_dir_golden.safetensors"))
    assert sorted(golden) == meta["keys"] and meta["listing"] == ["diffusion_pytorch_model.safetensors"]
    cfg = C.UNetConfig.from_any(TINY)
    assert set(C.merge_spec(cfg, 8, 8)) == set(golden)
    assert all(tuple(golden[k].shape) == tuple(s) for k, s in C.merge_spec(cfg, 8, 8).items())
    unet = UNet2DConditionModel(cfg, models.unet.state_dict())
    agn = ControlLoRAModel(cfg, models.lora_agnostic.state_dict(), 4, unet=unet)
    clo = ControlLoRAModel(cfg, models.lora_clothes.state_dict(), 4, unet=unet)
    pose = CachedControlNetModel(cfg, models.openpose.state_dict())
    # a directory as the reference leaves it: its top-level file + one sub-directory per distinct pattern index
    d = tmp_path / "ref_written"
    d.mkdir()
    shutil.copy(os.path.join(here, "checkpoint_dir_golden.safetensors"), d / "diffusion_pytorch_model.safetensors")
    saved = [c for c in meta["calls"] if c[0] == "save_pretrained"]
    assert [c[2] for c in saved] == ["controlnet_0", "controlnet_1"] and all(c[3] is None for c in saved)  # VAE detached
    agn.save_pretrained(str(d / "controlnet_0"))
    clo.save_pretrained(str(d / "controlnet_1"))
    multi = EdgeStyleMultiControlNetModel.from_pretrained(str(d), load_pattern=meta["pattern"],
                                                          controlnet_class=ControlLoRAModel,
                                                          static_controlnets=[None, pose, None, pose, None, pose],
                                                          latent_hw=(8, 8))
    assert all(torch.equal(multi.state_dict()[k], v) for k, v in golden.items())
    out = tmp_path / "mirror_written"
    multi.save_pretrained(str(out), save_pattern=meta["pattern"])
    assert sorted(os.listdir(out)) == sorted(meta["listing"] + [c[2] for c in saved])
    mine = load_file(str(out / "diffusion_pytorch_model.safetensors"))
    assert set(mine) == set(golden) and all(torch.equal(mine[k], golden[k]) for k in golden)


def test_merge_level_groups_follow_the_decoder():
    """Merge launches are grouped in the order the decoder consumes the residual levels (mid, then skips 11..0); every
    level appears exactly once whatever the grouping."""
    from edgestyle_b200.engine import merge_level_groups

    by_level = merge_level_groups(13, 3, True, True, 3)
    assert by_level == [[12, 11, 10, 9], [8, 7, 6], [5, 4, 3], [2, 1, 0]]
    assert merge_level_groups(13, 3, True, False, 3) == [[12, 11, 10, 9, 8, 7, 6, 5, 4, 3], [2, 1, 0]]
    assert merge_level_groups(13, 3, True, False, 0) == [list(range(12, -1, -1))]
    assert merge_level_groups(13, 3, False, True, 3) == [list(range(12, -1, -1))]      # mode "residuals": one group
    for nlev, n_up in ((13, 3), (9, 2), (5, 2), (4, 3)):
        for args in ((True, True, 3), (True, False, 3), (True, False, 99), (False, False, 0)):
            flat = [li for grp in merge_level_groups(nlev, n_up, *args) for li in grp]
            assert sorted(flat) == list(range(nlev)) and all(grp for grp in merge_level_groups(nlev, n_up, *args))
            if args[0] and args[1]:
                assert flat == list(range(nlev - 1, -1, -1))
