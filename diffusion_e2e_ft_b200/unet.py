"""B200UNet2DConditionModel — drop-in for diffusers' / GeoWizard's `UNet2DConditionModel` on the
single-step denoising path.

Call-compatible with the reference call sites
    Marigold/marigold/marigold_pipeline.py:452-454   unet(x, t, encoder_hidden_states=E).sample
    training/train.py:500                            unet(x, t, E, return_dict=False)[0]
    GeoWizard/.../geowizard_pipeline.py:319-321      unet(x, t.repeat(2), encoder_hidden_states=E, class_labels=C).sample
Control flow restates GeoWizard/geowizard/models/unet_2d_condition.py:845-1221; parameters keep the
diffusers `state_dict` names.  All arithmetic runs in libb200_e2eft.so.
"""
import torch
import torch.nn as nn

from . import ops
from .modules import (ConfigDict, ConvInSmall, ConvOutSmall, Downsample2D, Packed, ResnetBlock2D,
                      Transformer2DModel, Upsample2D, _f16, _f32)
from .checkpoint import PretrainedMixin
from .ops import F16, F32


class UNet2DConditionOutput:
    """Stand-in for diffusers' BaseOutput subclass: `.sample` plus tuple-style indexing."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _DownBlock(nn.Module):
    """CrossAttnDownBlock2D (unet_2d_blocks.py:1027-1185) / DownBlock2D (:1188-1273)."""

    def __init__(self, cin, cout, temb, n, heads, cross_dim, has_attn, add_down, groups, eps, joint):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(n)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, heads, cross_dim, groups, joint) for _ in range(n)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 1)]) if add_down else None


class _MidBlock(nn.Module):
    """UNetMidBlock2DCrossAttn (unet_2d_blocks.py:634-777)."""

    def __init__(self, ch, temb, heads, cross_dim, groups, eps, joint):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, cross_dim, groups, joint)])


class _UpBlock(nn.Module):
    """CrossAttnUpBlock2D (unet_2d_blocks.py:2201-2371) / UpBlock2D (:2374-2481)."""

    def __init__(self, cin, cout, cprev, temb, n, heads, cross_dim, has_attn, add_up, groups, eps, joint):
        super().__init__()
        rs = []
        for i in range(n):
            skip = cin if i == n - 1 else cout
            rin = cprev if i == 0 else cout
            rs.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(cout, heads, cross_dim, groups, joint) for _ in range(n)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None


_DEFAULTS = dict(
    in_channels=8, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
    up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
    layers_per_block=2, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
    norm_num_groups=32, norm_eps=1e-5, class_embed_type=None,
    projection_class_embeddings_input_dim=None, joint_attention=False,
    flip_sin_to_cos=True, freq_shift=0, sample_size=96, act_fn="silu", use_linear_projection=True)


class B200UNet2DConditionModel(PretrainedMixin, nn.Module):
    """`stream_dtype`: dtype of the residual stream inside the engine (fp32 = parity mode, fp16 = fast)."""
    _diffusers_class_name = "UNet2DConditionModel"
    _config_defaults = _DEFAULTS

    def __init__(self, stream_dtype=torch.float32, **config):
        super().__init__()
        cfg = ConfigDict(_DEFAULTS)
        unknown = set(config) - set(_DEFAULTS)
        if unknown:
            raise TypeError(f"unknown UNet config keys: {sorted(unknown)}")
        cfg.update(config)
        for k in ("block_out_channels", "down_block_types", "up_block_types", "attention_head_dim"):
            if isinstance(cfg[k], list):
                cfg[k] = tuple(cfg[k])
        if not cfg["flip_sin_to_cos"] or cfg["freq_shift"] != 0 or not cfg["use_linear_projection"]:
            raise NotImplementedError("engine supports the SD-2 embedding / linear-projection config only")
        self.config = cfg
        self.stream_dtype = stream_dtype
        boc = tuple(cfg["block_out_channels"])
        heads = tuple(cfg["attention_head_dim"])       # number of heads (unet_2d_condition.py:244-250)
        temb = boc[0] * 4
        g, eps, cd, J = cfg["norm_num_groups"], cfg["norm_eps"], cfg["cross_attention_dim"], cfg["joint_attention"]
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.class_embedding = (TimestepEmbedding(cfg["projection_class_embeddings_input_dim"], temb)
                                if cfg["class_embed_type"] == "projection" else None)
        n = cfg["layers_per_block"]
        downs, ch = [], boc[0]
        for i, t in enumerate(cfg["down_block_types"]):
            cin, ch = ch, boc[i]
            downs.append(_DownBlock(cin, ch, temb, n, heads[i], cd, t == "CrossAttnDownBlock2D",
                                    i != len(boc) - 1, g, eps, J))
        self.down_blocks = nn.ModuleList(downs)
        self.mid_block = _MidBlock(boc[-1], temb, heads[-1], cd, g, eps, J)
        rev, rheads = list(reversed(boc)), list(reversed(heads))
        ups, cout = [], rev[0]
        for i, t in enumerate(cfg["up_block_types"]):
            cprev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            ups.append(_UpBlock(cin, cout, cprev, temb, n + 1, rheads[i], cd, t == "CrossAttnUpBlock2D",
                                i != len(boc) - 1, g, eps, J))
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=eps)
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)
        self._pk = Packed()
        self._gradient_checkpointing = False

    # ------------------------------------------------------------------ diffusers-API shims
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        """No-op: the engine's attention is always the fused flash kernel (Marigold/run.py:284-287)."""

    def enable_gradient_checkpointing(self):
        self._gradient_checkpointing = True

    def register_to_config(self, **kw):
        """diffusers API used by the load hook (training/train.py:335): unknown keys are kept, not rejected."""
        self.config.update({k: v for k, v in kw.items() if k in _DEFAULTS})
        extra = {k: v for k, v in kw.items() if k not in _DEFAULTS and k != "_extra"}
        if extra:
            self.config.setdefault("_extra", {}).update(extra)

    def _resnets(self):
        for blk in self.down_blocks:
            yield from blk.resnets
        yield from self.mid_block.resnets
        for blk in self.up_blocks:
            yield from blk.resnets

    def _embed_packed(self):
        te, ce = self.time_embedding, self.class_embedding
        resnets = list(self._resnets())
        params = list(te.parameters()) + (list(ce.parameters()) if ce is not None else [])
        for r in resnets:
            params += [r.time_emb_proj.weight, r.time_emb_proj.bias]

        def build():
            d = dict(w1=_f16(te.linear_1.weight), b1=_f32(te.linear_1.bias),
                     w2=_f16(te.linear_2.weight), b2=_f32(te.linear_2.bias),
                     wall=_f16(torch.cat([r.time_emb_proj.weight for r in resnets], 0)),
                     ball=_f32(torch.cat([r.time_emb_proj.bias for r in resnets], 0)))
            if ce is not None:
                kin = ce.linear_1.weight.shape[1]
                kpad = (kin + 7) // 8 * 8
                w = torch.zeros(ce.linear_1.weight.shape[0], kpad, device=ce.linear_1.weight.device)
                w[:, :kin] = ce.linear_1.weight.detach()
                d.update(cw1=_f16(w), cb1=_f32(ce.linear_1.bias), cw2=_f16(ce.linear_2.weight),
                         cb2=_f32(ce.linear_2.bias), ckpad=kpad)
            offs, o = [], 0
            for r in resnets:
                offs.append(o)
                o += r.cout
            d["offs"] = offs
            return d
        return self._pk.get(params, build)

    # ------------------------------------------------------------------ forward
    # Exact single-step specialisations — pure wins on the reference's one-step path, all
    # parity-tested against the general path (tests/engine_checks.py:run_single_step_specialisations):
    #   * constant python-scalar timestep and no class labels (marigold_pipeline.py:452: `t` of the 1-step trailing
    #     schedule is always 999): temb and the 22 stacked `time_emb_proj` outputs are a function of the weights only
    #     -> computed once per weights version (unet_2d_condition.py:974-981), no embedding kernels per call;
    #   * zeros noise (marigold_pipeline.py:418-423): `sample` may carry only the leading channels, conv_in runs on
    #     those (the missing input channels are exact zeros);
    #   * one context shared by the batch (marigold_pipeline.py:428-432; detected without a device sync as a
    #     batch-broadcast view, `stride(0) == 0`, or batch 1): cross-attention collapses to two skinny GEMMs
    #     (modules.BasicTransformerBlock._packed_const_ctx).
    single_step_specialisations = True

    def _time_embedding(self, t, class_labels, B, dev):
        """[B, sum(cout)] fp32: every resnet's time_emb_proj(silu(temb (+class_emb))) (unet_2d_condition.py:957-1000)."""
        cfg = self.config
        ep = self._embed_packed()
        e = ops.timestep_embedding(t, cfg["block_out_channels"][0])
        e = ops.linear(e, ep["w1"], ep["b1"], act=ops.ACT_SILU)
        if self.class_embedding is not None:
            if class_labels is None:
                raise ValueError("class_labels should be provided when num_class_embeds > 0")
            cl = torch.zeros((B, ep["ckpad"]), dtype=F16, device=dev)
            cl[:, :class_labels.shape[1]] = class_labels
            c = ops.linear(cl, ep["cw1"], ep["cb1"], act=ops.ACT_SILU)
            c = ops.linear(c, ep["cw2"], ep["cb2"])
            e = ops.linear(e, ep["w2"], ep["b2"], residual=c, act=ops.ACT_SILU)     # silu(temb + class_emb)
        else:
            e = ops.linear(e, ep["w2"], ep["b2"], act=ops.ACT_SILU)                 # silu(temb)
        return ops.linear(e, ep["wall"], ep["ball"], out_dtype=F32)                  # all 22 time_emb_proj at once

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, return_dict=True, **unused):
        ops._need_cuda(sample)                                                      # sm_100a only, no CPU fallback
        if torch.is_grad_enabled() and (sample.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_train(sample, timestep, encoder_hidden_states, class_labels, return_dict)
        cfg, sdt = self.config, self.stream_dtype
        B, _, H, W = sample.shape
        dev = sample.device
        n_up = len(cfg["block_out_channels"]) - 1
        forward_size = (H % (2 ** n_up) != 0) or (W % (2 ** n_up) != 0)      # unet_2d_condition.py:920-930
        spec = self.single_step_specialisations
        if sample.shape[1] != cfg["in_channels"] and not (spec and sample.shape[1] < cfg["in_channels"]):
            raise ValueError(f"sample has {sample.shape[1]} channels, conv_in expects {cfg['in_channels']}")

        # ---- time / class embedding (unet_2d_condition.py:957-1000)
        ep = self._embed_packed()
        if spec and not torch.is_tensor(timestep) and self.class_embedding is None:
            cache = ep.setdefault("temb_cache", {})                 # lives and dies with the packed weights
            key = (float(timestep), B, str(dev))
            temb_all = cache.get(key)
            if temb_all is None:
                t = torch.full((B,), float(timestep), dtype=F32, device=dev)
                temb_all = cache[key] = self._time_embedding(t, None, B, dev)
                if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
                    cache.pop(key)                                   # graph-private memory must not outlive the capture
        else:
            if not torch.is_tensor(timestep):
                t = torch.full((B,), float(timestep), dtype=F32, device=dev)
            else:
                t = timestep.to(device=dev, dtype=F32).reshape(-1).expand(B).contiguous()
            temb_all = self._time_embedding(t, class_labels, B, dev)
        resnets = list(self._resnets())
        temb_of = {id(r): temb_all[:, o:o + r.cout] for r, o in zip(resnets, ep["offs"])}

        ehs = encoder_hidden_states
        const_ctx = None
        if spec and ehs.dim() == 3 and (ehs.shape[0] == 1 or ehs.stride(0) == 0):
            const_ctx = ehs[0]                                      # [S, Dctx] shared by every image
        ctx16 = ehs.to(F16).contiguous()

        # ---- down path
        if not hasattr(self, "_conv_in_run") or self._conv_in_run.conv is not self.conv_in:
            self._conv_in_run = ConvInSmall(self.conv_in)       # conv_in may be swapped (unet_prep.py:6-21)
        x = self._conv_in_run.run(sample if sample.dtype in (F16, F32) else sample.float(), sdt)
        skips = [x]
        for blk in self.down_blocks:
            for i, r in enumerate(blk.resnets):
                last = (i == len(blk.resnets) - 1) and blk.downsamplers is not None    # feeds the stride-2 conv
                x = r.run(x, temb_of[id(r)], None, sdt, f16_copy=last and blk.attentions is None)
                if blk.attentions is not None:
                    x = blk.attentions[i].run(x, ctx16, sdt, f16_copy=last, const_ctx=const_ctx)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0].run(x, sdt)
                skips.append(x)
        # ---- mid
        mb = self.mid_block
        x = mb.resnets[0].run(x, temb_of[id(mb.resnets[0])], None, sdt)
        x = mb.attentions[0].run(x, ctx16, sdt, const_ctx=const_ctx)
        x = mb.resnets[1].run(x, temb_of[id(mb.resnets[1])], None, sdt)
        # ---- up path
        for bi, blk in enumerate(self.up_blocks):
            for i, r in enumerate(blk.resnets):
                skip = skips.pop()
                last = (i == len(blk.resnets) - 1) and blk.upsamplers is not None and not forward_size
                x = r.run(x, temb_of[id(r)], skip, sdt, f16_copy=last and blk.attentions is None)
                if blk.attentions is not None:
                    x = blk.attentions[i].run(x, ctx16, sdt, f16_copy=last, const_ctx=const_ctx)
            if blk.upsamplers is not None:
                size = tuple(skips[-1].shape[1:3]) if forward_size else None
                x = blk.upsamplers[0].run(x, size, sdt)
        # ---- out
        if not hasattr(self, "_conv_out_run") or self._conv_out_run.conv is not self.conv_out:
            self._conv_out_run = ConvOutSmall(self.conv_norm_out, self.conv_out)
        out = self._conv_out_run.run(x)
        if out.dtype != sample.dtype:
            out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(out)


    # ------------------------------------------------------------------ differentiable forward (row a10)
    def _forward_train(self, sample, timestep, encoder_hidden_states, class_labels, return_dict):
        """Same graph as `forward`, every block executed as a torch.autograd.Function (autograd_blocks.py) so
        `loss.backward()` fills `.grad` of the parameters exactly like the reference's training/train.py:563."""
        from . import autograd_blocks as ab
        ck = bool(self._gradient_checkpointing)
        if self.stream_dtype != F32:
            raise NotImplementedError("training runs with the fp32 residual stream (stream_dtype=torch.float32)")
        cfg = self.config
        B, _, H, W = sample.shape
        dev = sample.device
        n_up = len(cfg["block_out_channels"]) - 1
        forward_size = (H % (2 ** n_up) != 0) or (W % (2 ** n_up) != 0)      # unet_2d_condition.py:920-930
        if not torch.is_tensor(timestep):
            t = torch.full((B,), float(timestep), dtype=F32, device=dev)
        else:
            t = timestep.to(device=dev, dtype=F32).reshape(-1).expand(B).contiguous()
        if self.class_embedding is not None and class_labels is None:
            raise ValueError("class_labels should be provided when num_class_embeds > 0")
        temb_all = ab.embed(self, t, class_labels)
        ep = self._embed_packed()
        resnets = list(self._resnets())
        temb_of = {id(r): temb_all[:, o:o + r.cout] for r, o in zip(resnets, ep["offs"])}
        ctx16 = encoder_hidden_states.detach().to(F16).contiguous()

        if not hasattr(self, "_conv_in_run") or self._conv_in_run.conv is not self.conv_in:
            self._conv_in_run = ConvInSmall(self.conv_in)
        x = ab.conv_in(self._conv_in_run, (sample if sample.dtype in (F16, F32) else sample.float()).detach())
        skips = [x]
        for blk in self.down_blocks:
            for i, r in enumerate(blk.resnets):
                last = (i == len(blk.resnets) - 1) and blk.downsamplers is not None
                x = ab.resnet(r, x, temb_of[id(r)], None, f16_copy=last and blk.attentions is None, ckpt=ck)
                if blk.attentions is not None:
                    x = ab.transformer(blk.attentions[i], x, ctx16, f16_copy=last, ckpt=ck)
                skips.append(x)
            if blk.downsamplers is not None:
                x = ab.downsample(blk.downsamplers[0], x)
                skips.append(x)
        mb = self.mid_block
        x = ab.resnet(mb.resnets[0], x, temb_of[id(mb.resnets[0])], ckpt=ck)
        x = ab.transformer(mb.attentions[0], x, ctx16, ckpt=ck)
        x = ab.resnet(mb.resnets[1], x, temb_of[id(mb.resnets[1])], ckpt=ck)
        for blk in self.up_blocks:
            for i, r in enumerate(blk.resnets):
                skip = skips.pop()
                last = (i == len(blk.resnets) - 1) and blk.upsamplers is not None and not forward_size
                x = ab.resnet(r, x, temb_of[id(r)], skip, f16_copy=last and blk.attentions is None, ckpt=ck)
                if blk.attentions is not None:
                    x = ab.transformer(blk.attentions[i], x, ctx16, f16_copy=last, ckpt=ck)
            if blk.upsamplers is not None:
                x = ab.upsample(blk.upsamplers[0], x, tuple(skips[-1].shape[1:3]) if forward_size else None)
        if not hasattr(self, "_conv_out_run") or self._conv_out_run.conv is not self.conv_out:
            self._conv_out_run = ConvOutSmall(self.conv_norm_out, self.conv_out)
        out = ab.conv_out(self._conv_out_run, x)
        if out.dtype != sample.dtype:
            out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNet2DConditionOutput(out)
