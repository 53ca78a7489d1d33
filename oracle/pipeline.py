"""ORACLE (test infrastructure) — restated host orchestration of the hot path.

* `DDIMOneStep`: the closed form `scheduler.set_timesteps(1)` (trailing) + `scheduler.step`
  reduces to on the single-step path (marigold_pipeline.py:401-402,457-465;
  training/train.py:509-518).  DDIMScheduler itself is third-party diffusers (absent);
  restated per SURVEY.md App. A.7.
* `marigold_single_infer`: Marigold/marigold/marigold_pipeline.py:371-478 (+ encode_rgb
  :481-498, decode_depth :501-519, decode_normal :522-538).
* `geowizard_single_infer`: GeoWizard/geowizard/models/geowizard_pipeline.py:251-344.
* `ensemble_normals`: marigold_pipeline.py:59-71.
* losses: training/util/loss.py:13-67.
"""
import math

import torch
import torch.nn.functional as F


class DDIMOneStep:
    """scaled-linear betas, v-prediction, trailing spacing (App. A.7)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 prediction_type="v_prediction", timestep_spacing="trailing"):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                               dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.timestep_spacing = timestep_spacing
        self.timesteps = None

    def set_timesteps(self, n, device=None):
        T = self.num_train_timesteps
        if self.timestep_spacing == "trailing":
            ts = torch.round(torch.arange(T, 0, -T / n)) - 1
        elif self.timestep_spacing == "leading":
            ts = (torch.arange(0, n) * (T // n)).flip(0)
        else:
            raise ValueError(self.timestep_spacing)
        self.timesteps = ts.long().to(device) if device is not None else ts.long()

    def pred_original_sample(self, model_output, t, sample):
        a = self.alphas_cumprod[int(t)].to(sample.device)
        sa, sb = a.sqrt().to(sample.dtype), (1 - a).sqrt().to(sample.dtype)
        if self.prediction_type == "v_prediction":
            return sa * sample - sb * model_output          # training/train.py:511-512
        if self.prediction_type == "epsilon":
            return (sample - sb * model_output) / sa        # :513-514
        if self.prediction_type == "sample":
            return model_output
        raise ValueError(self.prediction_type)


def encode_rgb(vae, rgb_in):
    h = vae.encoder(rgb_in)
    moments = vae.quant_conv(h)
    mean, _ = torch.chunk(moments, 2, dim=1)
    return mean * vae.config.scaling_factor


def decode_latent(vae, latent):
    z = vae.post_quant_conv(latent / vae.config.scaling_factor)
    return vae.decoder(z)


@torch.no_grad()
def marigold_single_infer(unet, vae, scheduler, rgb_in, empty_text_embed, noise="zeros",
                          normals=False, generator=None, return_latents=False):
    """One denoising step only (the E2E-FT path: denoising_steps=1)."""
    scheduler.set_timesteps(1)
    t = scheduler.timesteps[0]
    rgb_latent = encode_rgb(vae, rgb_in)
    if noise == "zeros":
        latent = torch.zeros_like(rgb_latent)
    elif noise == "gaussian":
        latent = torch.randn(rgb_latent.shape, generator=generator, dtype=rgb_latent.dtype)
    else:
        raise ValueError(noise)
    ctx = empty_text_embed.repeat(rgb_latent.shape[0], 1, 1)
    unet_input = torch.cat([rgb_latent, latent], dim=1)
    pred = unet(unet_input, t, encoder_hidden_states=ctx).sample
    x0 = scheduler.pred_original_sample(pred, t, latent)
    dec = decode_latent(vae, x0)
    if normals:
        out = dec / (torch.norm(dec, p=2, dim=1, keepdim=True) + 1e-5)
    else:
        out = (torch.clip(dec.mean(dim=1, keepdim=True), -1.0, 1.0) + 1.0) / 2.0
    if return_latents:
        return out, dict(rgb_latent=rgb_latent, unet_out=pred, x0=x0, decoded=dec)
    return out


def geowizard_class_embedding(domain, dtype=torch.float32, batch=1):
    """geowizard_pipeline.py:290-302, batched as train_depth_normal.py:684-704."""
    geo_class = torch.tensor([[0., 1.], [1., 0.]], dtype=dtype)
    geo_emb = torch.cat([torch.sin(geo_class), torch.cos(geo_class)], dim=-1)
    dom = {"indoor": [1., 0., 0.], "outdoor": [0., 1., 0.], "object": [0., 0., 1.]}[domain]
    dom = torch.tensor([dom], dtype=dtype).repeat(2 * batch, 1)
    dom_emb = torch.cat([torch.sin(dom), torch.cos(dom)], dim=-1)
    geo_emb = geo_emb.repeat_interleave(batch, 0)
    return torch.cat([geo_emb, dom_emb], dim=-1)                 # [2B, 10]


@torch.no_grad()
def geowizard_single_infer(unet, vae, scheduler, rgb_in, img_embed, domain="indoor", noise="zeros"):
    """Batched generalisation ([depth x B, normal x B]) of the one-image reference path."""
    B = rgb_in.shape[0]
    scheduler.set_timesteps(1)
    t = scheduler.timesteps[0]
    rgb_latent = encode_rgb(vae, rgb_in)
    assert noise == "zeros"
    geo_latent = torch.zeros_like(rgb_latent).repeat(2, 1, 1, 1)
    rgb_latent = rgb_latent.repeat(2, 1, 1, 1)
    ctx = img_embed.repeat(2, 1, 1) if img_embed.shape[0] == B else img_embed.repeat(2 * B, 1, 1)
    cls = geowizard_class_embedding(domain, rgb_in.dtype, B)
    pred = unet(torch.cat([rgb_latent, geo_latent], 1), t.repeat(2 * B), encoder_hidden_states=ctx,
                class_labels=cls).sample
    x0 = scheduler.pred_original_sample(pred, t, geo_latent)
    depth = decode_latent(vae, x0[:B]).mean(dim=1, keepdim=True)
    depth = (torch.clip(depth, -1.0, 1.0) + 1.0) / 2.0
    normal = decode_latent(vae, x0[B:])
    normal = normal / (torch.norm(normal, p=2, dim=1, keepdim=True) + 1e-5)
    return depth, -normal                                         # :342 sign flip


def ensemble_normals(preds):
    """marigold_pipeline.py:59-71 — returns (picked prediction, index)."""
    bsz, d, h, w = preds.shape
    n = preds / (torch.norm(preds, p=2, dim=1).unsqueeze(1) + 1e-5)
    phi = torch.atan2(n[:, 1], n[:, 0]).mean(dim=0)
    theta = torch.atan2(torch.norm(n[:, :2], p=2, dim=1), n[:, 2]).mean(dim=0)
    m = torch.zeros((d, h, w)).to(n)
    m[0] = torch.sin(theta) * torch.cos(phi)
    m[1] = torch.sin(theta) * torch.sin(phi)
    m[2] = torch.cos(theta)
    err = torch.acos(torch.clip(torch.cosine_similarity(m[None], n, dim=1), -0.999, 0.999))
    idx = torch.argmin(err.reshape(bsz, -1).sum(-1))
    return n[idx], int(idx)


def compute_scale_and_shift_masked(prediction, target, mask):
    """training/util/loss.py:31-47."""
    a_00 = torch.sum(mask * prediction * prediction, (1, 2))
    a_01 = torch.sum(mask * prediction, (1, 2))
    a_11 = torch.sum(mask, (1, 2))
    b_0 = torch.sum(mask * prediction * target, (1, 2))
    b_1 = torch.sum(mask * target, (1, 2))
    x_0, x_1 = torch.zeros_like(b_0), torch.zeros_like(b_1)
    det = a_00 * a_11 - a_01 * a_01
    valid = det > 0
    x_0[valid] = (a_11[valid] * b_0[valid] - a_01[valid] * b_1[valid]) / det[valid]
    x_1[valid] = (-a_01[valid] * b_0[valid] + a_00[valid] * b_1[valid]) / det[valid]
    return x_0, x_1


def ssi_loss(prediction, target, mask):
    """training/util/loss.py:17-29."""
    if mask.ndim == 4:
        mask = mask.squeeze(1)
    prediction, target = prediction.squeeze(1).float(), target.squeeze(1).float()
    scale, shift = compute_scale_and_shift_masked(prediction, target, mask)
    scaled = scale.view(-1, 1, 1) * prediction + shift.view(-1, 1, 1)
    return F.l1_loss(scaled[mask], target[mask])


def angular_loss(prediction, target, mask):
    """training/util/loss.py:56-67."""
    dot = torch.clamp(torch.sum(prediction.float() * target.float(), dim=1), -1.0, 1.0)
    return torch.acos(dot)[mask[:, 0]].mean()


def abs_rel(pred, gt):
    """Marigold/src/util/metric.py:34-44 (abs_relative_difference, no mask)."""
    return (torch.abs(pred - gt) / gt).mean()


def align_lstsq(pred, gt):
    """Marigold/src/util/alignment.py:38-47 — least-squares scale/shift of pred onto gt."""
    p, g = pred.reshape(-1, 1).double(), gt.reshape(-1, 1).double()
    A = torch.cat([p, torch.ones_like(p)], dim=1)
    x = torch.linalg.lstsq(A, g).solution
    return (pred.double() * x[0] + x[1]).to(pred.dtype)
