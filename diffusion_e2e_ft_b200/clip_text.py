"""CLIP text encoder on the engine kernels (the text half).

Drop-in for `transformers.CLIPTextModel` at the one place the hot path's callers use it —
Marigold/marigold/marigold_pipeline.py:355-369 (`encode_empty_text`: tokenizer("", padding="do_not_pad") ->
`self.text_encoder(text_input_ids)[0].to(self.dtype)`, once per process) and training/train.py's prompt encoding of
the same empty prompt.  Parameter names are transformers' (`text_model.embeddings.token_embedding.weight`,
`text_model.encoder.layers.N.self_attn.q_proj.weight`, ...), so `text_encoder/model.safetensors` of an SD-2
checkpoint loads with `load_state_dict`; `from_pretrained` reads `config.json` + weights like the other modules.

Arithmetic (transformers==4.37.2 models/clip/modeling_clip.py, restated in oracle/clip_text.py): token + position
embedding (b200_embed_tokens, fp32 stream) -> N x [LayerNorm -> fused q|k|v GEMM -> causal attention (head_dim 64:
the flash kernel, one launch per query position with that position's key prefix — the sequence is 2 tokens for the
empty prompt, 77 at most, and the encoder runs once per process) -> out_proj GEMM + residual -> LayerNorm -> fc1 GEMM
with the erf-GELU epilogue -> fc2 GEMM + residual] -> final LayerNorm.  No torch arithmetic; no CPU fallback.

`quick_gelu` towers (OpenAI CLIP: SD-1.x text encoders, the ViT-L/14 image encoder of clip_vision.py) run on the SiLU
epilogue with pre-scaled fc1 weights.  Not built: heads whose width is not 64.
"""
import json
import os
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import ops
from .modules import ConfigDict, Packed, _f16, _f32
from .ops import F16, F32

BOS_TOKEN_ID, EOS_TOKEN_ID = 49406, 49407       # CLIP BPE vocabulary: <|startoftext|>, <|endoftext|>


class EmptyPromptTokenizer:
    """The only tokenisation the hot path's callers perform is of the empty prompt (marigold_pipeline.py:361-368):
    `tokenizer("", padding="do_not_pad", max_length=model_max_length, truncation=True, return_tensors="pt")` ->
    input_ids [[BOS, EOS]].  The BPE vocabulary / merges files are not available offline, so any other prompt is
    refused instead of being mis-tokenised."""
    model_max_length = 77

    def __call__(self, text, padding="do_not_pad", max_length=None, truncation=True, return_tensors="pt"):
        texts = [text] if isinstance(text, str) else list(text)
        if any(t != "" for t in texts):
            raise NotImplementedError("EmptyPromptTokenizer only encodes the empty prompt (no BPE vocabulary offline); "
                                      "pass a transformers CLIPTokenizer for anything else")
        L = self.model_max_length if padding == "max_length" else 2
        ids = torch.full((len(texts), L), EOS_TOKEN_ID, dtype=torch.long)   # SD-2's pad token id is 0 (`!`) ...
        if padding == "max_length":
            ids[:, 2:] = 0                                                   # ... tokenizer/special_tokens_map.json
        ids[:, 0] = BOS_TOKEN_ID
        return SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids))


class _SelfAttn(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.k_proj, self.v_proj = nn.Linear(C, C), nn.Linear(C, C)
        self.q_proj, self.out_proj = nn.Linear(C, C), nn.Linear(C, C)


class _MLP(nn.Module):
    def __init__(self, C, I):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(C, I), nn.Linear(I, C)


QUICK_GELU_K = 1.702


class _EncoderLayer(nn.Module):
    """transformers CLIPEncoderLayer (pre-LN).  `act`: "gelu" -> the erf-GELU GEMM epilogue; "quick_gelu"
    (x * sigmoid(1.702 x), the OpenAI CLIP towers) -> the SiLU epilogue on fc1 weights / bias pre-scaled by 1.702
    (silu(1.702 z) = 1.702 * quick_gelu(z)), undone exactly by alpha = 1 / 1.702 on the fc2 accumulator."""

    def __init__(self, C, I, eps, act="gelu"):
        super().__init__()
        self.self_attn = _SelfAttn(C)
        self.layer_norm1 = nn.LayerNorm(C, eps=eps)
        self.mlp = _MLP(C, I)
        self.layer_norm2 = nn.LayerNorm(C, eps=eps)
        self.eps, self.act = eps, act
        self._pk = Packed()

    def packed(self):
        a, m = self.self_attn, self.mlp
        k = QUICK_GELU_K if self.act == "quick_gelu" else 1.0

        def build():
            return dict(ln1=(_f32(self.layer_norm1.weight), _f32(self.layer_norm1.bias)),
                        ln2=(_f32(self.layer_norm2.weight), _f32(self.layer_norm2.bias)),
                        wqkv=_f16(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                        bqkv=_f32(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                        wo=_f16(a.out_proj.weight), bo=_f32(a.out_proj.bias),
                        w1=_f16(m.fc1.weight.float() * k), b1=_f32(m.fc1.bias.float() * k),
                        w2=_f16(m.fc2.weight), b2=_f32(m.fc2.bias))
        return self._pk.get(list(self.parameters()), build)

    def run(self, h, B, L, heads, causal):
        """h: fp32 residual stream [B*L, C] -> the same after this layer."""
        pk = self.packed()
        C = h.shape[1]
        y = ops.layer_norm(h, *pk["ln1"], eps=self.eps)
        qkv = ops.linear(y, pk["wqkv"], pk["bqkv"]).view(B, L, 3 * C)
        if causal:
            o = torch.empty((B, L, C), dtype=F16, device=h.device)
            for i in range(L):                                   # query i sees keys 0..i
                ops.attention_d64(qkv[:, i:i + 1, :C], qkv[:, :i + 1, C:2 * C], qkv[:, :i + 1, 2 * C:], heads, 0.125,
                                  out=o[:, i:i + 1])
        else:
            o = ops.attention_d64(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, 0.125)
        h = ops.linear(o.view(B * L, C), pk["wo"], pk["bo"], residual=h, out_dtype=F32)
        y = ops.layer_norm(h, *pk["ln2"], eps=self.eps)
        if self.act == "quick_gelu":
            m = ops.linear(y, pk["w1"], pk["b1"], act=ops.ACT_SILU)
            return ops.linear(m, pk["w2"], pk["b2"], residual=h, out_dtype=F32, alpha=1.0 / QUICK_GELU_K)
        m = ops.linear(y, pk["w1"], pk["b1"], act=ops.ACT_GELU)
        return ops.linear(m, pk["w2"], pk["b2"], residual=h, out_dtype=F32)


class _Embeddings(nn.Module):
    def __init__(self, vocab, C, max_pos):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, C)
        self.position_embedding = nn.Embedding(max_pos, C)


class _Encoder(nn.Module):
    def __init__(self, n, C, I, eps, act="gelu"):
        super().__init__()
        self.layers = nn.ModuleList([_EncoderLayer(C, I, eps, act) for _ in range(n)])


class _TextTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        C = cfg["hidden_size"]
        self.embeddings = _Embeddings(cfg["vocab_size"], C, cfg["max_position_embeddings"])
        self.encoder = _Encoder(cfg["num_hidden_layers"], C, cfg["intermediate_size"], cfg["layer_norm_eps"],
                                cfg["hidden_act"])
        self.final_layer_norm = nn.LayerNorm(C, eps=cfg["layer_norm_eps"])


class CLIPTextOutput(tuple):
    """`out[0]` / `out.last_hidden_state`, `out[1]` / `out.pooler_output` (transformers BaseModelOutputWithPooling)."""

    def __new__(cls, last, pooled):
        o = super().__new__(cls, (last, pooled))
        o.last_hidden_state, o.pooler_output = last, pooled
        return o


_CFG_KEYS = ("vocab_size", "hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads",
             "max_position_embeddings", "layer_norm_eps", "hidden_act", "eos_token_id", "bos_token_id")


class B200CLIPTextModel(nn.Module):

    def __init__(self, vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                 num_attention_heads=16, max_position_embeddings=77, layer_norm_eps=1e-5, hidden_act="gelu",
                 eos_token_id=EOS_TOKEN_ID, bos_token_id=BOS_TOKEN_ID, **extra):
        super().__init__()
        if hidden_size != 64 * num_attention_heads:
            raise NotImplementedError(f"head width {hidden_size // num_attention_heads}: the attention kernel is d=64")
        if hidden_act not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"hidden_act={hidden_act!r}: gelu (SD-2) and quick_gelu (OpenAI CLIP) are built")
        self.config = ConfigDict(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                                 num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                                 max_position_embeddings=max_position_embeddings, layer_norm_eps=layer_norm_eps,
                                 hidden_act=hidden_act, eos_token_id=eos_token_id, bos_token_id=bos_token_id)
        if extra:
            self.config["_extra"] = dict(extra)
        self.text_model = _TextTransformer(self.config)
        self._pk = Packed()

    # -- nn.Module conveniences the pipeline reads (`self.text_encoder.device`, `.dtype`)
    @property
    def device(self):
        return self.text_model.final_layer_norm.weight.device

    @property
    def dtype(self):
        return self.text_model.final_layer_norm.weight.dtype

    def register_to_config(self, **kw):
        self.config.update(kw)

    # -- transformers directory layout: config.json + model.safetensors (or pytorch_model.bin)
    def save_pretrained(self, save_directory, safe_serialization=True, **unused):
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: v for k, v in self.config.items() if k != "_extra"}
        cfg.update(self.config.get("_extra", {}))
        cfg.update(architectures=["CLIPTextModel"], model_type="clip_text_model")
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        sd = {k: v.detach().to("cpu").contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, "model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, "pytorch_model.bin"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, **unused):
        d = pretrained_model_name_or_path if subfolder is None else os.path.join(pretrained_model_name_or_path, subfolder)
        if not os.path.isdir(d):
            raise FileNotFoundError(f"{d}: not a directory (local checkpoint folders only; there is no hub access)")
        with open(os.path.join(d, "config.json")) as f:
            raw = json.load(f)
        known = {k: raw[k] for k in _CFG_KEYS if k in raw}
        if known.get("eos_token_id") == 2:                      # legacy configs carry eos_token_id 2; the real id is 49407
            known["eos_token_id"] = EOS_TOKEN_ID
        model = cls(**known)
        extra = {k: v for k, v in raw.items() if k not in _CFG_KEYS and k not in ("architectures", "model_type")}
        if extra:
            model.config["_extra"] = extra
        safe, binp = os.path.join(d, "model.safetensors"), os.path.join(d, "pytorch_model.bin")
        if os.path.exists(safe):
            from safetensors.torch import load_file
            sd = load_file(safe)
        elif os.path.exists(binp):
            sd = torch.load(binp, map_location="cpu")
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin in {d}")
        model.load_state_dict(sd, strict=True)
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        return model.eval()

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = dict(state_dict)
        sd.pop("text_model.embeddings.position_ids", None)        # a persistent buffer in transformers < 4.31 checkpoints
        return super().load_state_dict(sd, strict=strict, **kw)

    @torch.no_grad()
    def forward(self, input_ids, attention_mask=None, **_ignored):
        """input_ids [B, L] int64 (L <= max_position_embeddings) -> CLIPTextOutput.  `attention_mask` is accepted and —
        as in the Stable Diffusion pipelines, which never pass one to this model — only the causal mask applies."""
        cfg = self.config
        ids = input_ids.to(self.device, torch.long).contiguous()
        ops._need_cuda(ids)
        B, L = ids.shape
        C, H = cfg["hidden_size"], cfg["num_attention_heads"]
        if L > cfg["max_position_embeddings"]:
            raise ValueError(f"sequence length {L} > max_position_embeddings {cfg['max_position_embeddings']}")
        emb = self.text_model.embeddings
        h = ops.embed_tokens(ids, emb.token_embedding.weight, emb.position_embedding.weight)
        for layer in self.text_model.encoder.layers:
            h = layer.run(h, B, L, H, causal=True)
        fl = self.text_model.final_layer_norm
        fin = self._pk.get([fl.weight, fl.bias], lambda: (_f32(fl.weight), _f32(fl.bias)))
        last16 = ops.layer_norm(h, *fin, eps=cfg["layer_norm_eps"]).view(B, L, C)
        last = last16.to(self.dtype)
        # pooled = hidden state at the first EOS of each row (index bookkeeping on 2..77 ids: host-side plumbing)
        eos = (ids == cfg["eos_token_id"]).int().argmax(dim=-1)
        return CLIPTextOutput(last, last[torch.arange(B, device=ids.device), eos])
