"""Generate the committed golden fixtures from the ORACLE (fp32, CPU, seeded synthetic weights).

The reference ships no golden vectors and cannot be imported here (diffusers/xformers absent), so
these pin the oracle restatement against silent drift and give the GPU tests a fixed target.
    python tests/golden/make_golden.py        -> tests/golden/*.pt   (a few hundred KB)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.unet import UNet2DConditionRef, tiny_config, seeded_init  # noqa: E402
from oracle.vae import AutoencoderKLRef, tiny_vae_config  # noqa: E402
from oracle import pipeline as P  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs(seed, *shape, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def build_tiny(kind="marigold"):
    if kind == "geowizard":
        cfg = tiny_config(class_embed_type="projection", projection_class_embeddings_input_dim=10,
                          cross_attention_dim=96, joint_attention=True)
    else:
        cfg = tiny_config()
    unet = seeded_init(UNet2DConditionRef(cfg), seed=1234).eval()
    vae = seeded_init(AutoencoderKLRef(tiny_vae_config()), seed=77).eval()
    return unet, vae


@torch.no_grad()
def main():
    torch.set_num_threads(8)
    out = {}
    unet, vae = build_tiny()
    # UNet alone, even and odd latent sizes, ctx 2 and 77 tokens
    for name, (h, w, s) in dict(unet_16x16_ctx2=(16, 16, 2), unet_15x20_ctx77=(15, 20, 77)).items():
        x = inputs(1, 2, 8, h, w)
        ctx = inputs(2, 2, s, 128, scale=0.5)
        out[name] = dict(y=unet(x, 999, ctx).sample)
    # VAE
    rgb = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1
    lat = P.encode_rgb(vae, rgb)
    out["vae_encode_64"] = dict(y=lat)
    z = inputs(4, 2, 4, 8, 8, scale=0.5)
    out["vae_decode_8"] = dict(y=P.decode_latent(vae, z))
    # pipeline
    sched = P.DDIMOneStep()
    ete = inputs(5, 1, 2, 128, scale=0.5)
    depth, lats = P.marigold_single_infer(unet, vae, sched, rgb, ete, return_latents=True)
    normals = P.marigold_single_infer(unet, vae, sched, rgb, ete, normals=True)
    out["marigold_depth_64"] = dict(y=depth, unet_out=lats["unet_out"], x0=lats["x0"])
    out["marigold_normals_64"] = dict(y=normals)
    # GeoWizard joint depth+normals
    gunet, _ = build_tiny("geowizard")
    emb = inputs(6, 2, 1, 96, scale=0.5)
    d, n = P.geowizard_single_infer(gunet, vae, sched, rgb, emb, domain="indoor")
    out["geowizard_64"] = dict(depth=d, normal=n)
    # weight fingerprints: detect RNG / init drift
    fp = lambda m: float(sum(p.double().abs().sum() for p in m.parameters()))
    out["fingerprint"] = dict(unet=fp(unet), vae=fp(vae), gunet=fp(gunet))
    torch.save(out, os.path.join(HERE, "golden_tiny.pt"))
    for k, v in out.items():
        print(k, {a: (tuple(b.shape) if torch.is_tensor(b) else b) for a, b in v.items()})


if __name__ == "__main__":
    main()
