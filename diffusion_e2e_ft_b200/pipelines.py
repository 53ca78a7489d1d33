"""Diffusers-free restatement of the reference's host pipelines on top of the engine modules.

`diffusers` (DiffusionPipeline, BaseOutput, DDIMScheduler) and `matplotlib` are not importable in
this image, so the reference pipeline files cannot even be imported; these classes keep their
public surface — `__call__`, `single_infer`, `encode_rgb`, `decode_depth`, `decode_normal`, the
output dataclasses — and route the arithmetic through B200UNet2DConditionModel / B200AutoencoderKL.

  MarigoldPipeline                 <- Marigold/marigold/marigold_pipeline.py:113-538
  DepthNormalEstimationPipeline    <- GeoWizard/geowizard/models/geowizard_pipeline.py:67-401
  DDIMScheduler (1-step closed form) <- diffusers DDIMScheduler as used at marigold_pipeline.py:401-402,457-465
"""
import math
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np
import torch

from . import ops


# ------------------------------------------------------------------------------------ scheduler
class SchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    """Subset of diffusers' DDIMScheduler the reference touches: `set_timesteps`, `timesteps`,
    `step(...).prev_sample / .pred_original_sample`, `alphas_cumprod`, `config`.  scaled-linear betas,
    eta = 0, no clipping/thresholding (SD-2 config)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 prediction_type="v_prediction", timestep_spacing="trailing", steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.config = dict(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                           timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.num_inference_steps = None
        self.timesteps = None
        self._ac = [float(a) for a in self.alphas_cumprod]       # host copy: no device sync in step()

    def set_timesteps(self, num_inference_steps, device=None):
        T = self.config["num_train_timesteps"]
        self.num_inference_steps = num_inference_steps
        if self.config["timestep_spacing"] == "trailing":
            ts = np.round(np.arange(T, 0, -T / num_inference_steps)) - 1
        elif self.config["timestep_spacing"] == "leading":
            ts = (np.arange(0, num_inference_steps) * (T // num_inference_steps)).round()[::-1].copy()
            ts = ts + self.config["steps_offset"]
        else:
            raise ValueError(self.config["timestep_spacing"])
        self._host_timesteps = [int(t) for t in ts]
        self.timesteps = torch.tensor(self._host_timesteps, dtype=torch.long, device=device)

    def coefficients(self, t_index):
        """(t, t_prev, a_t, a_prev) for the i-th inference step, all host floats/ints."""
        t = self._host_timesteps[t_index]
        prev = t - self.config["num_train_timesteps"] // self.num_inference_steps
        a_t = self._ac[t]
        a_prev = self._ac[prev] if prev >= 0 else 1.0
        return t, prev, a_t, a_prev


# ------------------------------------------------------------------------------------ base
class PipelineBase:
    """Minimal `DiffusionPipeline` surface the reference relies on: register_modules, to, device, dtype."""

    def register_modules(self, **modules):
        self._module_names = list(modules)
        for k, v in modules.items():
            setattr(self, k, v)

    def to(self, *args, **kwargs):
        for k in self._module_names:
            m = getattr(self, k)
            if isinstance(m, torch.nn.Module):
                m.to(*args, **kwargs)
        return self

    @property
    def device(self):
        return next(self.unet.parameters()).device

    @property
    def dtype(self):
        return next(self.unet.parameters()).dtype

    # ---- CUDA-graph replay of a fixed-shape step (launch-bound inner loop: ~1800 kernels / step)
    use_cuda_graph = True

    def _weights_key(self):
        """Identity of every weight a captured graph has baked in: storage pointers + torch version counters of the
        parameters AND the engine's weights epoch (`modules.bump_weights_epoch`: the flat-buffer optimizer kernel
        updates parameters with raw device writes that version counters do not see — without the epoch a validation
        `single_infer` after a training step would replay a graph that points at freed packed-weight buffers)."""
        from .modules import _WEIGHTS_EPOCH
        ps = self.__dict__.get("_wk_params")
        mods = (id(self.unet), id(self.vae), id(getattr(self.unet, "conv_in", None)))
        if ps is None or self.__dict__.get("_wk_mods") != mods:
            ps = list(self.unet.parameters()) + list(self.vae.parameters())
            self.__dict__["_wk_params"], self.__dict__["_wk_mods"] = ps, mods
        return (_WEIGHTS_EPOCH[0], hash(tuple((p.data_ptr(), p._version) for p in ps)))

    def _graphed(self, key, fn, x):
        """Capture `fn(static_x)` once per (key, weights version) and replay it; returns a fresh tensor."""
        graphs = self.__dict__.setdefault("_graphs", {})
        wkey = self._weights_key()
        ent = graphs.get(key)
        if ent is None or ent["wkey"] != wkey:
            static_x = x.clone()
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):                         # warm-up: packs weights, sets func attributes
                    fn(static_x)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            before = (ops.STATS.launches, dict(ops.STATS.flops), dict(ops.STATS.count))
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = fn(static_x)
            delta = dict(launches=ops.STATS.launches - before[0],
                         flops={k: ops.STATS.flops[k] - before[1][k] for k in before[1]},
                         count={k: ops.STATS.count[k] - before[2][k] for k in before[2]})
            ent = dict(g=g, x=static_x, out=out, wkey=wkey, delta=delta)
            graphs[key] = ent
        ent["x"].copy_(x, non_blocking=True)
        ent["g"].replay()
        d = ent["delta"]
        ops.STATS.launches += d["launches"]
        for k in d["flops"]:
            ops.STATS.flops[k] += d["flops"][k]
            ops.STATS.count[k] += d["count"][k]
        return ent["out"].clone()


@dataclass
class MarigoldDepthOutput:
    depth_np: Optional[np.ndarray]
    depth_colored: Optional[object]
    uncertainty: Optional[np.ndarray]
    normal_np: Optional[np.ndarray]
    normal_colored: Optional[object]


from .ensemble import (ensemble_depths, ensemble_normals, minmax_normalise_, minmax_rows, normalise_rgb,  # noqa: E402
                       resize_bilinear_aa, resize_nearest)


def pyramid_noise_like(x, discount=0.9, generator=None):
    """Multi-resolution noise of marigold_pipeline.py:76-86 (also training/train.py:486-487): white noise plus
    bilinearly up-sampled coarser noise maps with geometrically decaying weights, renormalised to unit std.  The
    random draws (torch.randn, python `random`) are host-pipeline code as in the reference; they are not part of
    the measured hot path (E2E-FT uses zeros noise)."""
    import random
    b, c, w, h = x.shape
    u = torch.nn.Upsample(size=(w, h), mode="bilinear")
    noise = torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=generator)
    for i in range(10):
        r = random.random() * 2 + 2
        w, h = max(1, int(w / (r ** i))), max(1, int(h / (r ** i)))
        noise += u(torch.randn(b, c, w, h, device=x.device, dtype=x.dtype, generator=generator)) * discount ** i
        if w == 1 or h == 1:
            break
    return noise / noise.std()


def _resize_max_res(img, max_edge):
    """Marigold/marigold/util/image_util.py:79-108 resize_max_res: antialiased bilinear down-scale to a maximum edge
    length, on the device (csrc/postproc.cu)."""
    _, h, w = img.shape
    s = min(max_edge / w, max_edge / h)
    return resize_bilinear_aa(img, (int(h * s), int(w * s)))


class MarigoldPipeline(PipelineBase):
    rgb_latent_scale_factor = 0.18215
    depth_latent_scale_factor = 0.18215

    def __init__(self, unet, vae, scheduler, text_encoder=None, tokenizer=None, empty_text_embed=None):
        self.register_modules(unet=unet, vae=vae, scheduler=scheduler, text_encoder=text_encoder,
                              tokenizer=tokenizer)
        self.empty_text_embed = empty_text_embed          # [1, 2, 1024]; CLIP weights are not available offline

    @torch.no_grad()
    def __call__(self, input_image, denoising_steps: int = 10, ensemble_size: int = 10,
                 processing_res: int = 768, match_input_res: bool = True, resample_method: str = "bilinear",
                 batch_size: int = 0, color_map: Optional[str] = "Spectral", show_progress_bar: bool = True,
                 ensemble_kwargs=None, noise="gaussian", normals=False) -> MarigoldDepthOutput:
        assert processing_res >= 0 and ensemble_size >= 1
        if resample_method != "bilinear":
            raise NotImplementedError("the engine resizes with the reference's default (bilinear, antialiased)")
        if isinstance(input_image, torch.Tensor):
            rgb = input_image.squeeze()
        else:                                              # PIL.Image
            rgb = torch.from_numpy(np.asarray(input_image.convert("RGB")).copy()).permute(2, 0, 1)
        input_size = rgb.shape
        assert rgb.dim() == 3 and input_size[0] == 3, f"Wrong input shape {input_size}, expected [rgb, H, W]"
        # pre-processing on the device: the raw (uint8) image is uploaded once; resize + [0,255] ->
        # [-1,1] run as kernels (marigold_pipeline.py:237-247)
        was_u8 = rgb.dtype == torch.uint8
        rgb = rgb.to(self.device)
        rgb = rgb.to(torch.float32)                        # dtype cast at the API boundary
        if processing_res > 0:
            rgb = _resize_max_res(rgb, processing_res)
        rgb_norm = normalise_rgb(rgb, round_u8=was_u8 and processing_res > 0)
        lo, hi = minmax_rows(rgb_norm.view(1, -1))[0].tolist()
        assert lo >= -1.0 and hi <= 1.0
        rgb_norm = rgb_norm.to(self.dtype)
        duplicated = torch.stack([rgb_norm] * ensemble_size)
        bs = batch_size if batch_size > 0 else ensemble_size
        preds = []
        for i in range(0, ensemble_size, bs):
            preds.append(self.single_infer(duplicated[i:i + bs], denoising_steps, show_progress_bar,
                                           noise=noise, normals=normals).detach())
        preds = torch.concat(preds, dim=0).squeeze()
        pred_uncert = None
        if ensemble_size > 1:                              # test-time ensembling on the device (:288-294)
            if normals:
                pred, pred_uncert = ensemble_normals(preds)
            else:
                pred, pred_uncert = ensemble_depths(preds, **(ensemble_kwargs or {}))
        else:
            pred = preds
        pred = pred.to(torch.float32).contiguous()
        if normals:
            pred = ops.decode_post(pred[None], normals=True)[0]            # pred / (|pred| + 1e-5)   (:300-303)
        else:
            pred, mm = minmax_normalise_(pred)                             # (pred - min) / (max - min)   (:305-312)
            lo, hi = mm.tolist()
            if hi == lo:
                pred = torch.zeros_like(pred)
        if match_input_res:
            pred = resize_bilinear_aa(pred if normals else pred.unsqueeze(0),
                                      (input_size[-2], input_size[-1])).squeeze()
        pred = pred.cpu().numpy()
        if pred_uncert is not None:
            pred_uncert = pred_uncert.cpu().numpy() if torch.is_tensor(pred_uncert) else pred_uncert
        pred = pred.clip(-1.0, 1.0) if normals else pred.clip(0, 1)
        # colourising needs matplotlib (absent): the color_map=None path of marigold_pipeline.py:330-338
        return MarigoldDepthOutput(depth_np=None if normals else pred, depth_colored=None, uncertainty=pred_uncert,
                                   normal_np=pred if normals else None, normal_colored=None)

    def encode_empty_text(self):
        if self.text_encoder is None:
            raise RuntimeError("no text encoder: pass `empty_text_embed` ([1,2,1024]) to MarigoldPipeline")
        ids = self.tokenizer("", padding="do_not_pad", max_length=self.tokenizer.model_max_length,
                             truncation=True, return_tensors="pt").input_ids.to(self.text_encoder.device)
        self.empty_text_embed = self.text_encoder(ids)[0].to(self.dtype)

    @torch.no_grad()
    def single_infer(self, rgb_in, num_inference_steps: int, show_pbar: bool = False, noise="gaussian",
                     normals=False, generator=None):
        device = self.device
        rgb_in = rgb_in.to(device)
        # the kernels index each tensor with 32-bit element offsets: split batches whose largest activation
        # (256 channels at full resolution in the VAE decoder) would exceed 2^32 elements
        B, _, H, W = rgb_in.shape
        max_b = max(1, (2 ** 32 - 1) // (256 * H * W))
        if B > max_b:
            return torch.cat([self.single_infer(rgb_in[i:i + max_b], num_inference_steps, show_pbar, noise=noise,
                                                normals=normals, generator=generator)
                              for i in range(0, B, max_b)], dim=0)
        if (self.use_cuda_graph and noise == "zeros" and num_inference_steps == 1 and rgb_in.is_cuda
                and not torch.cuda.is_current_stream_capturing()):
            key = ("marigold", tuple(rgb_in.shape), rgb_in.dtype, bool(normals))
            return self._graphed(key, lambda x: self._single_infer_impl(x, 1, noise, normals, None), rgb_in)
        return self._single_infer_impl(rgb_in, num_inference_steps, noise, normals, generator)

    def _single_infer_impl(self, rgb_in, num_inference_steps, noise, normals, generator):
        device = rgb_in.device
        self.scheduler.set_timesteps(num_inference_steps)        # host-side only: graph-capture safe
        rgb_latent = self.encode_rgb(rgb_in)
        if noise == "gaussian":
            latent = torch.randn(rgb_latent.shape, device=device, dtype=rgb_latent.dtype, generator=generator)
        elif noise == "pyramid":
            latent = pyramid_noise_like(rgb_latent, generator=generator)
        elif noise == "zeros":
            latent = None                                   # exact zeros: never materialised
        else:
            raise ValueError(f"Unknown noise type: {noise}")
        if self.empty_text_embed is None:
            self.encode_empty_text()
        # one context for the whole batch (marigold_pipeline.py:428-432 `.repeat`s it): passed as a broadcast view so
        # the UNet can take its constant-context cross-attention path (same values, no copy)
        ctx = self.empty_text_embed.to(device).expand(rgb_latent.shape[0], -1, -1)
        zeros = None
        spec = getattr(self.unet, "single_step_specialisations", False)
        pt = self.scheduler.config["prediction_type"]
        for i in range(num_inference_steps):
            t, _, a_t, a_prev = self.scheduler.coefficients(i)
            if latent is None and spec:
                cur = None
                unet_input = rgb_latent                               # the zero half is never materialised: conv_in on 4 channels
            else:
                if latent is None:
                    zeros = torch.zeros_like(rgb_latent) if zeros is None else zeros
                    cur = zeros
                else:
                    cur = latent
                unet_input = torch.cat([rgb_latent, cur], dim=1)      # this order is important (:447-449)
            pred = self.unet(unet_input, t, encoder_hidden_states=ctx).sample
            sa, sb = math.sqrt(a_t), math.sqrt(1.0 - a_t)
            # x0 = c_x * x_t + c_m * model_out  (DDIM, eta = 0)
            if pt == "v_prediction":
                c_x, c_m = sa, -sb
            elif pt == "epsilon":
                c_x, c_m = 1.0 / sa, -sb / sa
            elif pt == "sample":
                c_x, c_m = 0.0, 1.0
            else:
                raise ValueError(pt)
            if i == num_inference_steps - 1:
                # last step: latent = pred_original_sample, fused with /scale + post_quant_conv + decoder
                dec = self.vae.decode_from_prediction(pred, c_m, noisy=latent, c_noisy=c_x)
                break
            # intermediate DDIM step (eta = 0): x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev) eps
            x_t = cur if cur is not None else torch.zeros_like(rgb_latent)
            x0 = c_x * x_t + c_m * pred
            eps = (x_t - sa * x0) / sb
            latent = math.sqrt(a_prev) * x0 + math.sqrt(1.0 - a_prev) * eps
        if normals:
            return ops.decode_post(dec.float().contiguous(), normals=True).to(dec.dtype)
        return ops.decode_post(dec.float().contiguous(), normals=False).to(dec.dtype)

    def encode_rgb(self, rgb_in):
        return self.vae.encode_scaled_mean(rgb_in)

    def decode_depth(self, depth_latent):
        z = self.vae.post_quant_conv(depth_latent, scale_in=1.0 / self.depth_latent_scale_factor)
        return self.vae.decoder(z).mean(dim=1, keepdim=True)

    def decode_normal(self, normal_latent):
        z = self.vae.post_quant_conv(normal_latent, scale_in=1.0 / self.depth_latent_scale_factor)
        return self.vae.decoder(z)


# ------------------------------------------------------------------------------------ GeoWizard
@dataclass
class DepthNormalPipelineOutput:
    depth_np: np.ndarray
    depth_colored: Optional[object]
    normal_np: np.ndarray
    normal_colored: Optional[object]
    uncertainty: Optional[np.ndarray] = None


class DepthNormalEstimationPipeline(PipelineBase):
    """GeoWizard joint depth+normal pipeline (geowizard_pipeline.py).  The image context comes from `image_encoder`
    (clip_vision.B200CLIPVisionModelWithProjection, or any object with the transformers interface) through
    `encode_img_embed` (:232-248), or is passed in as `img_embed` ([B or 1, 1, 768])."""

    latent_scale_factor = 0.18215

    def __init__(self, unet, vae, scheduler, image_encoder=None, feature_extractor=None):
        self.register_modules(unet=unet, vae=vae, scheduler=scheduler, image_encoder=image_encoder,
                              feature_extractor=feature_extractor)
        self.img_embed = None

    @staticmethod
    def class_embedding(domain, batch, device, dtype):
        """geowizard_pipeline.py:290-302 batched as train_depth_normal.py:684-704 -> [2B, 10]."""
        geo_class = torch.tensor([[0., 1.], [1., 0.]], device=device, dtype=dtype)
        geo = torch.cat([torch.sin(geo_class), torch.cos(geo_class)], dim=-1).repeat_interleave(batch, 0)
        dom = {"indoor": [1., 0., 0.], "outdoor": [0., 1., 0.], "object": [0., 0., 1.]}[domain]
        dom = torch.tensor([dom], device=device, dtype=dtype).repeat(2 * batch, 1)
        return torch.cat((geo, torch.cat([torch.sin(dom), torch.cos(dom)], dim=-1)), dim=-1)

    @torch.no_grad()
    def single_infer(self, input_rgb, num_inference_steps: int, domain: str, show_pbar: bool = False,
                     noise="zeros", img_embed=None):
        device = input_rgb.device
        B = input_rgb.shape[0]
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        if num_inference_steps != 1 or noise != "zeros":
            raise NotImplementedError("engine pipeline implements the E2E-FT setting: 1 step, zeros noise")
        rgb_latent = self.encode_RGB(input_rgb)
        geo_latent = torch.zeros_like(rgb_latent).repeat(2, 1, 1, 1)
        rgb_latent = rgb_latent.repeat(2, 1, 1, 1)
        emb = img_embed if img_embed is not None else self.img_embed
        if emb is None and self.image_encoder is not None:
            emb = self.encode_img_embed(input_rgb)
        if emb is None:
            raise RuntimeError("no image_encoder registered: pass img_embed ([B or 1,1,768])")
        ctx = emb.to(device)
        ctx = ctx.repeat(2, 1, 1) if ctx.shape[0] == B else ctx.repeat(2 * B, 1, 1)
        cls = self.class_embedding(domain, B, device, rgb_latent.dtype)
        t, _, a_t, _ = self.scheduler.coefficients(0)
        pred = self.unet(torch.cat([rgb_latent, geo_latent], dim=1), torch.full((2 * B,), t, device=device),
                         encoder_hidden_states=ctx, class_labels=cls).sample
        assert self.scheduler.config["prediction_type"] == "v_prediction"
        c_m = -math.sqrt(1.0 - a_t)
        d = self.vae.decode_from_prediction(pred[:B].contiguous(), c_m)
        n = self.vae.decode_from_prediction(pred[B:].contiguous(), c_m)
        depth = ops.decode_post(d.float().contiguous(), normals=False).to(d.dtype)
        normal = ops.decode_post(n.float().contiguous(), normals=True, sign=-1.0).to(n.dtype)   # :342 sign flip
        return depth, normal

    @torch.no_grad()
    def encode_img_embed(self, rgb):
        """geowizard_pipeline.py:232-248: bicubic-antialiased resize of (rgb + 1) / 2 to the crop size, CLIP mean / std,
        image_encoder(...).image_embeds.unsqueeze(1) -> [B, 1, 768] (one context row per input image)."""
        enc = self.image_encoder
        if hasattr(enc, "preprocess"):                                    # the engine encoder: device kernels
            x = enc.preprocess(rgb.float().contiguous(), self.feature_extractor)
        else:
            raise RuntimeError("image_encoder has no device `preprocess`; use B200CLIPVisionModelWithProjection or pass img_embed")
        return enc(x.to(enc.dtype)).image_embeds.unsqueeze(1).to(self.dtype)

    def encode_RGB(self, rgb_in):
        return self.vae.encode_scaled_mean(rgb_in)

    def decode_depth(self, depth_latent):
        z = self.vae.post_quant_conv(depth_latent, scale_in=1.0 / self.latent_scale_factor)
        return self.vae.decoder(z).mean(dim=1, keepdim=True)

    def decode_normal(self, normal_latent):
        z = self.vae.post_quant_conv(normal_latent, scale_in=1.0 / self.latent_scale_factor)
        return self.vae.decoder(z)

    @torch.no_grad()
    def __call__(self, input_image, denoising_steps: int = 1, ensemble_size: int = 1, processing_res: int = 768,
                 match_input_res: bool = True, domain: str = "indoor", color_map: Optional[str] = None,
                 show_progress_bar: bool = False, noise="zeros", img_embed=None) -> DepthNormalPipelineOutput:
        if isinstance(input_image, torch.Tensor):
            rgb = input_image.squeeze()
        else:
            rgb = torch.from_numpy(np.asarray(input_image.convert("RGB")).copy()).permute(2, 0, 1)
        input_size = rgb.shape
        was_u8 = rgb.dtype == torch.uint8
        rgb = rgb.to(self.device).to(torch.float32)
        if processing_res > 0:
            rgb = _resize_max_res(rgb, processing_res)
        rgb_norm = normalise_rgb(rgb, round_u8=was_u8 and processing_res > 0).to(self.dtype)
        dl, nl = [], []
        for _ in range(ensemble_size):                      # geowizard_pipeline.py:139-176 (batch size 1 per member)
            d, n = self.single_infer(rgb_norm[None], denoising_steps, domain, show_progress_bar, noise, img_embed)
            dl.append(d)
            nl.append(n)
        depth, normal = torch.cat(dl).squeeze(), torch.cat(nl).squeeze()
        uncert = None
        if ensemble_size > 1:                               # :179-188
            depth, uncert = ensemble_depths(depth)
            normal, _ = ensemble_normals(normal)
        depth, _ = minmax_normalise_(depth.to(torch.float32).contiguous())      # :192-194
        normal = normal.to(torch.float32)
        if match_input_res:
            # the reference resizes on the host (PIL for depth, cv2 INTER_NEAREST for normals, :201-206); here on the device
            depth = resize_bilinear_aa(depth[None], tuple(input_size[-2:]))[0]
            normal = resize_nearest(normal, tuple(input_size[-2:]))
        return DepthNormalPipelineOutput(depth_np=depth.cpu().numpy().clip(0, 1), depth_colored=None,
                                         normal_np=normal.cpu().numpy().clip(-1, 1), normal_colored=None,
                                         uncertainty=None if uncert is None else uncert.cpu().numpy())
