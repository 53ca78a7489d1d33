"""CLIP text encoder (text half; Marigold/marigold/marigold_pipeline.py:355-369).

  * not gpu: the ORACLE (oracle/clip_text.py) is pinned against the installed `transformers` CLIPTextModel — the
    third-party implementation the reference calls (requirements.txt:7 pins 4.37.2; 5.5.0 is what this image has) — on
    shared random weights; the engine module's host logic (parameter names, fused q|k|v packing, the causal key-prefix
    loop, pooled output, checkpoint round trip, pipeline hook) runs on the CPU emulation of the kernels;
  * gpu: the CUDA path against the oracle, tiny and at the SD-2 text-encoder size (23 layers x 1024, 2 and 77 tokens).
"""
import pytest
import torch

from oracle.clip_text import CLIPTextCfg, clip_text_forward, random_state_dict, tiny_clip_cfg

BOS, EOS = 49406, 49407


def _ids(batch, L, seed=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1000, 40000, (batch, L), generator=g)
    ids[:, 0] = BOS
    ids[:, -1] = EOS
    if L > 4:
        ids[0, L // 2] = EOS                       # an early EOS: the pooled output must take the FIRST one
    return ids


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _hf(cfg, sd):
    tr = pytest.importorskip("transformers")
    c = tr.CLIPTextConfig(vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                          num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
                          max_position_embeddings=cfg.max_position_embeddings, hidden_act=cfg.hidden_act,
                          layer_norm_eps=cfg.layer_norm_eps, eos_token_id=cfg.eos_token_id, bos_token_id=cfg.bos_token_id,
                          pad_token_id=1)
    m = tr.CLIPTextModel(c).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


@pytest.mark.parametrize("cfg,L", [(tiny_clip_cfg(), 2), (tiny_clip_cfg(), 9), (tiny_clip_cfg(hidden_act="quick_gelu"), 5),
                                   (CLIPTextCfg(num_hidden_layers=2), 2), (CLIPTextCfg(num_hidden_layers=2), 77)])
def test_oracle_matches_transformers(cfg, L):
    sd = random_state_dict(cfg, seed=3)
    ids = _ids(2, L, seed=L)
    with torch.no_grad():
        want = _hf(cfg, sd)(ids)
    last, pooled = clip_text_forward(sd, cfg, ids)
    assert _rel(last, want.last_hidden_state) <= 2e-6, _rel(last, want.last_hidden_state)
    assert _rel(pooled, want.pooler_output) <= 2e-6


def test_state_dict_names_are_transformers_names():
    from diffusion_e2e_ft_b200 import B200CLIPTextModel
    cfg = tiny_clip_cfg()
    eng = B200CLIPTextModel(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2)
    hf = _hf(cfg, random_state_dict(cfg))
    want = {k: tuple(v.shape) for k, v in hf.state_dict().items() if "position_ids" not in k}
    assert {k: tuple(v.shape) for k, v in eng.state_dict().items()} == want
    sd = dict(hf.state_dict())
    sd["text_model.embeddings.position_ids"] = torch.arange(77)[None]        # transformers < 4.31 checkpoints carry it
    eng.load_state_dict(sd, strict=True)
    with torch.device("meta"):
        full = B200CLIPTextModel()
    assert sum(p.numel() for p in full.parameters()) == 340_387_840          # SD-2 text_encoder (OpenCLIP ViT-H, 23 layers)
    with pytest.raises(NotImplementedError):
        B200CLIPTextModel(hidden_size=768, num_attention_heads=8)            # head width 96
    with pytest.raises(NotImplementedError):
        B200CLIPTextModel(hidden_act="relu")


def test_host_logic_on_cpu_emulation(monkeypatch, tmp_path):
    import cpu_emulation
    from diffusion_e2e_ft_b200 import B200CLIPTextModel, EmptyPromptTokenizer, MarigoldPipeline
    cpu_emulation.install(monkeypatch)
    qcfg = tiny_clip_cfg(hidden_act="quick_gelu")                      # SD-1.x text encoders: SiLU epilogue on 1.702-scaled fc1
    qsd = random_state_dict(qcfg, seed=6)
    q = B200CLIPTextModel(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                          hidden_act="quick_gelu").eval()
    q.load_state_dict(qsd)
    assert _rel(q(_ids(2, 5))[0], clip_text_forward(qsd, qcfg, _ids(2, 5))[0]) <= 3e-3
    cfg = tiny_clip_cfg()
    sd = random_state_dict(cfg, seed=5)
    eng = B200CLIPTextModel(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2).eval()
    eng.load_state_dict(sd)
    for L in (2, 9):
        ids = _ids(2, L, seed=L)
        out = eng(ids)
        last, pooled = clip_text_forward(sd, cfg, ids)
        assert out[0].shape == (2, L, 128) and out.last_hidden_state is out[0]
        assert _rel(out[0], last) <= 3e-3 and _rel(out.pooler_output, pooled) <= 3e-3
    # non-causal attention would differ visibly: token 0 must not see token 1
    a = eng(torch.tensor([[BOS, EOS]]))[0][:, 0]
    b = eng(torch.tensor([[BOS, 1234]]))[0][:, 0]
    assert torch.equal(a, b)
    # checkpoint round trip in the transformers layout
    eng.save_pretrained(str(tmp_path / "text_encoder"))
    assert (tmp_path / "text_encoder" / "model.safetensors").exists()
    again = B200CLIPTextModel.from_pretrained(str(tmp_path), subfolder="text_encoder", torch_dtype=torch.float16)
    assert again.dtype == torch.float16 and again.config["num_hidden_layers"] == 2
    assert torch.equal(again.state_dict()["text_model.final_layer_norm.weight"],
                       eng.state_dict()["text_model.final_layer_norm.weight"].half())
    # the pipeline hook (marigold_pipeline.py:355-369)
    tok = EmptyPromptTokenizer()
    assert tok("", padding="do_not_pad", max_length=77, truncation=True, return_tensors="pt").input_ids.tolist() == [[BOS, EOS]]
    with pytest.raises(NotImplementedError):
        tok("a photo")
    pipe = MarigoldPipeline(torch.nn.Linear(1, 1), None, None, text_encoder=eng, tokenizer=tok)   # unet only gives .dtype
    pipe.encode_empty_text()
    assert pipe.empty_text_embed.shape == (1, 2, 128)
    assert _rel(pipe.empty_text_embed, clip_text_forward(sd, cfg, torch.tensor([[BOS, EOS]]))[0]) <= 3e-3


def test_no_cpu_fallback():
    from diffusion_e2e_ft_b200 import B200CLIPTextModel
    eng = B200CLIPTextModel(hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eng(torch.tensor([[BOS, EOS]]))


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("size,L,batch", [("tiny", 2, 1), ("tiny", 9, 2), ("sd2", 2, 1), ("sd2", 77, 2)])
def test_cuda_clip_text_matches_oracle(size, L, batch):
    from diffusion_e2e_ft_b200 import B200CLIPTextModel
    cfg = tiny_clip_cfg() if size == "tiny" else CLIPTextCfg()
    sd = random_state_dict(cfg, seed=11)
    kw = dict(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
              num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads)
    eng = B200CLIPTextModel(**kw).eval()
    eng.load_state_dict(sd)
    eng = eng.cuda()
    ids = _ids(batch, L, seed=L)
    out = eng(ids.cuda())
    torch.cuda.synchronize()
    last, pooled = clip_text_forward(sd, cfg, ids)
    r = _rel(out[0].cpu(), last)
    print(size, L, batch, r)
    assert torch.isfinite(out[0]).all()
    assert r <= 3e-3 and _rel(out.pooler_output.cpu(), pooled) <= 3e-3, r
    # fp16 module (what the pipelines hold): same function within the fp16 parameter rounding
    out16 = eng.half()(ids.cuda())
    assert out16[0].dtype == torch.float16 and _rel(out16[0].float().cpu(), last) <= 6e-3


@pytest.mark.gpu
def test_cuda_pipeline_encodes_the_empty_prompt_with_the_engine_encoder():
    """MarigoldPipeline.encode_empty_text -> B200CLIPTextModel -> ctx [1,2,C] feeding the UNet's constant-context path."""
    import engine_checks as EC
    r = EC.run_pipeline_with_text_encoder()
    print(r)
    assert r["ctx_rel_l2"] <= 3e-3 and r["depth_rel_l2"] <= 3e-3, r
