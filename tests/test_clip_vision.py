"""CLIP image encoder (vision half; GeoWizard/geowizard/models/geowizard_pipeline.py:232-248).

  * not gpu: the ORACLE (oracle/clip_vision.py) is pinned against the installed `transformers`
    CLIPVisionModelWithProjection on shared random weights; the engine module's host logic (names, patch gather order,
    class / position rows, quick_gelu re-scaling, checkpoint round trip) runs on the CPU emulation of the kernels;
  * gpu: the CUDA path against the oracle — tiny, and the ViT-L/14 of the GeoWizard checkpoint (24 x 1024, 257 tokens);
    the bicubic antialiased resize against torch; the GeoWizard pipeline with the engine encoder end to end.
"""
import pytest
import torch

from oracle.clip_vision import (CLIPVisionCfg, clip_vision_forward, geowizard_img_embed, random_vision_state_dict,
                                tiny_vision_cfg)


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _kw(cfg):
    return dict(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
                num_attention_heads=cfg.num_attention_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                projection_dim=cfg.projection_dim, layer_norm_eps=cfg.layer_norm_eps, hidden_act=cfg.hidden_act)


def _hf(cfg, sd):
    tr = pytest.importorskip("transformers")
    m = tr.CLIPVisionModelWithProjection(tr.CLIPVisionConfig(**_kw(cfg))).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


def _pixels(batch, size, seed=0):
    return torch.randn(batch, 3, size, size, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("cfg", [tiny_vision_cfg(), tiny_vision_cfg(hidden_act="gelu"), CLIPVisionCfg(num_hidden_layers=2)])
def test_oracle_matches_transformers(cfg):
    sd = random_vision_state_dict(cfg, seed=3)
    x = _pixels(2, cfg.image_size, seed=1)
    with torch.no_grad():
        want = _hf(cfg, sd)(x)
    emb, last = clip_vision_forward(sd, cfg, x)
    assert _rel(emb, want.image_embeds) <= 2e-6 and _rel(last, want.last_hidden_state) <= 2e-6


def test_state_dict_names_are_transformers_names():
    from diffusion_e2e_ft_b200 import B200CLIPVisionModelWithProjection
    cfg = tiny_vision_cfg()
    eng = B200CLIPVisionModelWithProjection(**_kw(cfg))
    hf = _hf(cfg, random_vision_state_dict(cfg))
    want = {k: tuple(v.shape) for k, v in hf.state_dict().items() if "position_ids" not in k}
    assert {k: tuple(v.shape) for k, v in eng.state_dict().items()} == want
    eng.load_state_dict(hf.state_dict(), strict=True)
    with torch.device("meta"):
        full = B200CLIPVisionModelWithProjection()
    assert sum(p.numel() for p in full.parameters()) == 303_966_208 + 0      # OpenAI ViT-L/14 vision tower + 768 projection


def test_host_logic_on_cpu_emulation(monkeypatch, tmp_path):
    import cpu_emulation
    from diffusion_e2e_ft_b200 import B200CLIPVisionModelWithProjection, CLIPImageProcessorConfig
    cpu_emulation.install(monkeypatch)
    for cfg in (tiny_vision_cfg(), tiny_vision_cfg(hidden_act="gelu")):
        sd = random_vision_state_dict(cfg, seed=5)
        eng = B200CLIPVisionModelWithProjection(**_kw(cfg)).eval()
        eng.load_state_dict(sd)
        x = _pixels(2, cfg.image_size, seed=2)
        out = eng(x)
        emb, last = clip_vision_forward(sd, cfg, x)
        assert out.image_embeds.shape == (2, cfg.projection_dim) and out.last_hidden_state.shape == last.shape
        assert _rel(out.image_embeds, emb) <= 3e-3 and _rel(out.last_hidden_state, last) <= 3e-3
    eng.save_pretrained(str(tmp_path / "image_encoder"))
    again = B200CLIPVisionModelWithProjection.from_pretrained(str(tmp_path), subfolder="image_encoder", torch_dtype=torch.float16)
    assert again.dtype == torch.float16 and again.config["hidden_act"] == "gelu" and again.config["image_size"] == 56
    fe = CLIPImageProcessorConfig()
    assert fe.crop_size == {"height": 224, "width": 224} and len(fe.image_mean) == 3
    with pytest.raises(ValueError):
        eng(_pixels(1, 42))


def test_no_cpu_fallback():
    from diffusion_e2e_ft_b200 import B200CLIPVisionModelWithProjection
    cfg = tiny_vision_cfg()
    eng = B200CLIPVisionModelWithProjection(**_kw(cfg))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng(_pixels(1, cfg.image_size))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng.preprocess(torch.zeros(1, 3, 64, 64))


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("size,batch", [("tiny", 2), ("vitl14", 1), ("vitl14", 3)])
def test_cuda_clip_vision_matches_oracle(size, batch):
    from diffusion_e2e_ft_b200 import B200CLIPVisionModelWithProjection
    cfg = tiny_vision_cfg() if size == "tiny" else CLIPVisionCfg()
    sd = random_vision_state_dict(cfg, seed=11)
    eng = B200CLIPVisionModelWithProjection(**_kw(cfg)).eval()
    eng.load_state_dict(sd)
    eng = eng.cuda()
    x = _pixels(batch, cfg.image_size, seed=4)
    out = eng(x.cuda())
    torch.cuda.synchronize()
    sd_dev = {k: v.cuda() for k, v in sd.items()} if size != "tiny" else sd          # 24 fp32 layers: run the checker on the GPU
    torch.backends.cuda.matmul.allow_tf32 = False
    emb, last = clip_vision_forward(sd_dev, cfg, x.cuda() if size != "tiny" else x)
    r = (_rel(out.image_embeds.cpu(), emb.cpu()), _rel(out.last_hidden_state.cpu(), last.cpu()))
    print(size, batch, r)
    assert torch.isfinite(out.image_embeds).all()
    assert r[0] <= 3e-3 and r[1] <= 3e-3, r


@pytest.mark.gpu
def test_cuda_bicubic_antialiased_resize_and_preprocess():
    """geowizard_pipeline.py:236-245 on the device vs torch (F.interpolate bicubic antialias = what TF.resize calls)."""
    from diffusion_e2e_ft_b200 import B200CLIPVisionModelWithProjection
    from diffusion_e2e_ft_b200.ensemble import resize_bicubic_aa
    F = torch.nn.functional
    g = torch.Generator().manual_seed(9)
    for (h, w), size in (((480, 640), (224, 224)), ((768, 768), (224, 224)), ((100, 130), (224, 224)), ((224, 224), (224, 224))):
        x = torch.rand(2, 3, h, w, generator=g) * 2 - 1
        want = F.interpolate(x, size=size, mode="bicubic", antialias=True, align_corners=False)
        got = resize_bicubic_aa(x.cuda(), size)
        assert (got.cpu() - want).abs().max().item() <= 2e-5, ((h, w), (got.cpu() - want).abs().max().item())
    cfg = tiny_vision_cfg()
    eng = B200CLIPVisionModelWithProjection(**_kw(cfg)).cuda()
    rgb = torch.rand(2, 3, 96, 128, generator=g) * 2 - 1
    from oracle.clip_vision import CLIP_MEAN, CLIP_STD
    want = F.interpolate((rgb + 1) / 2, size=(56, 56), mode="bicubic", antialias=True, align_corners=False)
    want = (want - torch.tensor(CLIP_MEAN)[None, :, None, None]) / torch.tensor(CLIP_STD)[None, :, None, None]
    from diffusion_e2e_ft_b200 import CLIPImageProcessorConfig
    got = eng.preprocess(rgb.cuda(), CLIPImageProcessorConfig(56))
    assert (got.cpu() - want).abs().max().item() <= 2e-5


@pytest.mark.gpu
def test_cuda_geowizard_pipeline_with_the_engine_image_encoder():
    import engine_checks as EC
    r = EC.run_geowizard_with_image_encoder()
    print(r)
    assert r["img_embed_rel_l2"] <= 3e-3 and r["depth_rel_l2"] <= 3e-3 and r["normal_angle_deg"] <= 0.5, r
