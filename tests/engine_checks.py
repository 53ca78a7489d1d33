"""Graph-level parity: engine modules (CUDA, via the C ABI) vs the ORACLE on identical seeded weights
and inputs.  Shared by tests/test_engine_gpu.py, __graft_entry__.smoke() and tools/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from diffusion_e2e_ft_b200 import (B200AutoencoderKL, B200UNet2DConditionModel, DDIMScheduler,  # noqa: E402
                                   DepthNormalEstimationPipeline, MarigoldPipeline)
from oracle import pipeline as OP  # noqa: E402
import make_golden as MG  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "golden_tiny.pt")


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def mean_angle_deg(a, b):
    """DSINE/utils/utils.py:150-178 style per-pixel angular error between unit-normal maps (degrees)."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    cos = torch.clamp((a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1) + 1e-12), -1.0, 1.0)
    return torch.rad2deg(torch.acos(cos)).mean().item()


def absrel_protocol(pred_engine, pred_oracle, noise=0.05, seed=0):
    """north_star accuracy gate: `depth AbsRel within 1e-3 of the reference`.  No eval split offline, so
    build a synthetic ground truth the way the benchmark sees one: metric depth in [0.5, 10] m derived from
    the oracle prediction plus 5 % noise (so the oracle's own AbsRel is benchmark-like, ~5e-2), then score
    BOTH predictions with the reference protocol — least-squares scale/shift alignment
    (Marigold/src/util/alignment.py:38-47) then abs_relative_difference (Marigold/src/util/metric.py:34-44) —
    and report the difference."""
    pe, po = pred_engine.detach().float().cpu(), pred_oracle.detach().float().cpu()
    g = torch.Generator().manual_seed(seed)
    gt = 0.5 + 9.5 * po
    gt = (gt * (1.0 + noise * torch.randn(gt.shape, generator=g))).clamp(0.5, 10.0)
    ar_e = OP.abs_rel(OP.align_lstsq(pe, gt).clamp(0.5, 10.0), gt).item()
    ar_o = OP.abs_rel(OP.align_lstsq(po, gt).clamp(0.5, 10.0), gt).item()
    return dict(absrel_engine=ar_e, absrel_oracle=ar_o, absrel_delta=abs(ar_e - ar_o))


def engine_from_oracle(unet_ref, vae_ref, device, stream_dtype=torch.float32, dtype=torch.float32, vae_stream_dtype=None):
    cfg = unet_ref.config
    unet = B200UNet2DConditionModel(
        stream_dtype=stream_dtype, in_channels=cfg.in_channels, block_out_channels=cfg.block_out_channels,
        attention_head_dim=cfg.attention_head_dim, cross_attention_dim=cfg.cross_attention_dim,
        class_embed_type=cfg.class_embed_type,
        projection_class_embeddings_input_dim=cfg.projection_class_embeddings_input_dim,
        joint_attention=cfg.joint_attention)
    unet.load_state_dict(unet_ref.state_dict(), strict=True)
    vae = None
    if vae_ref is not None:
        vae = B200AutoencoderKL(stream_dtype=vae_stream_dtype or stream_dtype,
                                block_out_channels=vae_ref.config.block_out_channels)
        vae.load_state_dict(vae_ref.state_dict(), strict=True)
        vae = vae.to(device=device, dtype=dtype).eval().requires_grad_(False)
    return unet.to(device=device, dtype=dtype).eval().requires_grad_(False), vae


@torch.no_grad()
def run_marigold_tiny(device="cuda:0", stream_dtype=torch.float32):
    gold = torch.load(GOLD)
    unet_ref, vae_ref = MG.build_tiny()
    unet, vae = engine_from_oracle(unet_ref, vae_ref, device, stream_dtype)
    out = {}
    for name, (h, w, s) in dict(unet_16x16_ctx2=(16, 16, 2), unet_15x20_ctx77=(15, 20, 77)).items():
        y = unet(MG.inputs(1, 2, 8, h, w).to(device), 999, MG.inputs(2, 2, s, 128, scale=0.5).to(device)).sample
        out[name] = rel_l2(y, gold[name]["y"])
    rgb = (torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1).to(device)
    out["vae_encode"] = rel_l2(vae.encode_scaled_mean(rgb), gold["vae_encode_64"]["y"])
    z = MG.inputs(4, 2, 4, 8, 8, scale=0.5).to(device)
    out["vae_decode"] = rel_l2(vae.decoder(vae.post_quant_conv(z, scale_in=1 / 0.18215)), gold["vae_decode_8"]["y"])
    pipe = MarigoldPipeline(unet, vae, DDIMScheduler(), empty_text_embed=MG.inputs(5, 1, 2, 128, scale=0.5).to(device))
    depth = pipe.single_infer(rgb, 1, False, noise="zeros")
    normals = pipe.single_infer(rgb, 1, False, noise="zeros", normals=True)
    out["depth_rel_l2"] = rel_l2(depth, gold["marigold_depth_64"]["y"])
    out["normals_rel_l2"] = rel_l2(normals, gold["marigold_normals_64"]["y"])
    out["unet_rel_l2"] = max(out["unet_16x16_ctx2"], out["unet_15x20_ctx77"])
    out["normals_mean_angle_deg"] = mean_angle_deg(normals, gold["marigold_normals_64"]["y"])
    out.update(absrel_protocol(depth, gold["marigold_depth_64"]["y"]))
    return out


@torch.no_grad()
def run_single_step_specialisations(device="cuda:0", hw=(16, 16), full_width=False):
    """the exact single-step specialisations (cached constant-t embedding, conv_in on the 4 non-zero
    channels, constant-context cross-attention as two skinny GEMMs) against (a) the engine's own general path on the
    same inputs and (b) the fp32 oracle."""
    if full_width:
        from oracle.unet import UNet2DConditionRef, UNetConfig, seeded_init
        ref = seeded_init(UNet2DConditionRef(UNetConfig()), seed=4321).eval()
        dctx = 1024
    else:
        ref, _ = MG.build_tiny()
        dctx = 128
    unet, _ = engine_from_oracle(ref, None, device)
    lat = MG.inputs(21, 2, 4, *hw)
    ctx1 = MG.inputs(22, 1, 2, dctx, scale=0.5)
    x8 = torch.cat([lat, torch.zeros_like(lat)], 1)
    want = ref(x8, 999, ctx1.repeat(2, 1, 1)).sample
    unet.single_step_specialisations = False
    general = unet(x8.to(device), 999, ctx1.repeat(2, 1, 1).to(device)).sample
    unet.single_step_specialisations = True
    ctx_dev = ctx1.to(device)
    spec = unet(lat.to(device), 999, ctx_dev.expand(2, -1, -1)).sample            # 4-channel sample, broadcast context
    spec2 = unet(lat.to(device), 999, ctx_dev.expand(2, -1, -1)).sample           # second call: cached tables
    # a per-image (non-broadcast) context must take the general path and still be right
    ctx2 = MG.inputs(23, 2, 2, dctx, scale=0.5)
    want2 = ref(x8, 999, ctx2).sample
    got2 = unet(x8.to(device), 999, ctx2.to(device)).sample
    return dict(spec_vs_general=rel_l2(spec, general), spec_vs_oracle=rel_l2(spec, want),
                general_vs_oracle=rel_l2(general, want), repeat_call=rel_l2(spec2, spec),
                per_image_ctx_vs_oracle=rel_l2(got2, want2))


@torch.no_grad()
def run_geowizard_tiny(device="cuda:0", stream_dtype=torch.float32):
    gold = torch.load(GOLD)
    gunet_ref, vae_ref = MG.build_tiny("geowizard")
    unet, vae = engine_from_oracle(gunet_ref, vae_ref, device, stream_dtype)
    rgb = (torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1).to(device)
    emb = MG.inputs(6, 2, 1, 96, scale=0.5).to(device)
    pipe = DepthNormalEstimationPipeline(unet, vae, DDIMScheduler())
    d, n = pipe.single_infer(rgb, 1, "indoor", img_embed=emb)
    return dict(depth=rel_l2(d, gold["geowizard_64"]["depth"]), normal=rel_l2(n, gold["geowizard_64"]["normal"]),
                normal_mean_angle_deg=mean_angle_deg(n, gold["geowizard_64"]["normal"]))


@torch.no_grad()
def run_unet_fullwidth(device="cuda:0", latent=24, batch=1, stream_dtype=torch.float32):
    """Full SD-2 widths (320..1280, heads 5/10/20/20, ctx 1024) at a small latent so the CPU oracle
    finishes in seconds.  Exercises the production tile shapes."""
    from oracle.unet import UNet2DConditionRef, UNetConfig, seeded_init
    ref = seeded_init(UNet2DConditionRef(UNetConfig()), seed=4321).eval()
    unet, _ = engine_from_oracle(ref, None, device, stream_dtype)
    x = MG.inputs(11, batch, 8, latent, latent)
    ctx = MG.inputs(12, batch, 2, 1024, scale=0.5)
    want = ref(x, 999, ctx).sample
    got = unet(x.to(device), 999, ctx.to(device)).sample
    out = dict(rel_l2=rel_l2(got, want), max_abs=(got.cpu() - want).abs().max().item(), ref_std=want.std().item())
    # the reference's own fp16 GPU path (torch eager: cuBLAS/cuDNN/SDPA-math in fp16) against the same fp32 oracle
    ref16 = ref.half().to(device)
    y16 = ref16(x.half().to(device), 999, ctx.half().to(device)).sample
    out["torch_fp16_rel_l2"] = rel_l2(y16, want)
    return out


@torch.no_grad()
def run_full_size(device="cuda:0", res=768, batch=1, stream_dtype=torch.float32, vae_stream_dtype=None):
    """BASELINE.json full size: 768x768, SD-2 widths.  The oracle itself is run in fp32 on the GPU with plain
    torch ops (test infrastructure) so the comparison finishes in seconds; also checks size-independent
    properties (range of depth, unit normals, batch consistency)."""
    from oracle.unet import UNet2DConditionRef, UNetConfig, seeded_init
    from oracle.vae import AutoencoderKLRef, VAEConfig
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    uref = seeded_init(UNet2DConditionRef(UNetConfig()), seed=4321).eval()
    vref = seeded_init(AutoencoderKLRef(VAEConfig()), seed=99).eval()
    unet, vae = engine_from_oracle(uref, vref, device, stream_dtype, vae_stream_dtype=vae_stream_dtype)
    uref, vref = uref.to(device), vref.to(device)
    g = torch.Generator().manual_seed(7)
    rgb = (torch.rand(batch, 3, res, res, generator=g) * 2 - 1).to(device)
    ete = (torch.randn(1, 2, 1024, generator=g) * 0.5).to(device)
    want, lat = OP.marigold_single_infer(uref, vref, OP.DDIMOneStep(), rgb, ete, return_latents=True)
    pipe = MarigoldPipeline(unet, vae, DDIMScheduler(), empty_text_embed=ete)
    got = pipe.single_infer(rgb, 1, False, noise="zeros")
    out = dict(depth_rel_l2=rel_l2(got, want))
    out["rgb_latent_rel_l2"] = rel_l2(pipe.encode_rgb(rgb), lat["rgb_latent"])
    x = torch.cat([lat["rgb_latent"], torch.zeros_like(lat["rgb_latent"])], 1)
    out["unet_rel_l2"] = rel_l2(unet(x, 999, ete.repeat(batch, 1, 1)).sample, lat["unet_out"])
    out["decode_rel_l2"] = rel_l2(vae.decode_from_prediction(lat["unet_out"], -float((1 - OP.DDIMOneStep().alphas_cumprod[999]).sqrt())),
                                  lat["decoded"])
    out.update(absrel_protocol(got, want))
    out["depth_min"], out["depth_max"] = got.min().item(), got.max().item()
    nrm = pipe.single_infer(rgb, 1, False, noise="zeros", normals=True)
    out["normals_norm_err"] = (nrm.float().norm(dim=1) - 1).abs().max().item()
    # batch consistency: the same image twice in one batch gives the same answer as alone
    two = pipe.single_infer(torch.cat([rgb[:1], rgb[:1]]), 1, False, noise="zeros")
    out["batch_consistency"] = max(rel_l2(two[0:1], got[0:1]), rel_l2(two[1:2], got[0:1]))
    return out


@torch.no_grad()
def run_training_forward_tiny(device="cuda:0", modality="depth"):
    """Forward half of training/train.py:469-556 (bs=2, ctx 77 tokens) on the engine vs the oracle."""
    from diffusion_e2e_ft_b200.training import e2e_ft_forward
    unet_ref, vae_ref = MG.build_tiny()
    unet, vae = engine_from_oracle(unet_ref, vae_ref, device)
    g = torch.Generator().manual_seed(13)
    rgb = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    ctx = torch.randn(1, 77, 128, generator=g) * 0.5
    mask = torch.rand(2, 1, 64, 64, generator=g) > 0.2
    sched_o = OP.DDIMOneStep()
    lat = OP.encode_rgb(vae_ref, rgb)
    v = unet_ref(torch.cat([lat, torch.zeros_like(lat)], 1), 999, ctx.repeat(2, 1, 1)).sample
    dec = OP.decode_latent(vae_ref, sched_o.pred_original_sample(v, 999, torch.zeros_like(lat)))
    if modality == "depth":
        gt = torch.rand(2, 1, 64, 64, generator=g) * 9.9 + 0.1
        want = OP.ssi_loss(dec.mean(1, keepdim=True).clamp(-1, 1), gt, mask)
    else:
        gt = torch.nn.functional.normalize(torch.randn(2, 3, 64, 64, generator=g), dim=1)
        est = dec / (dec.norm(dim=1, keepdim=True) + 1e-5)
        want = OP.angular_loss(est.clamp(-1, 1), gt, mask)
    got, _ = e2e_ft_forward(unet, vae, DDIMScheduler(), rgb.to(device), gt.to(device), mask.to(device), ctx.to(device),
                            modality)
    return dict(loss_engine=got.item(), loss_oracle=want.item(), rel_err=abs(got.item() - want.item()) / abs(want.item()))


def run_unet_backward_tiny(device="cuda:0", hw=(16, 16), ctx_tokens=77, kind="marigold"):
    """Row a10: gradients of every UNet parameter through the engine's autograd blocks vs torch.autograd through
    the fp32 oracle, same weights / inputs / upstream gradient (bs=2).  Returns the forward error, the global
    relative L2 error over all parameter gradients and the worst single parameter."""
    unet_ref, _ = MG.build_tiny(kind)
    unet, _ = engine_from_oracle(unet_ref, None, device)
    unet.requires_grad_(True)
    unet_ref.requires_grad_(True)
    x = MG.inputs(1, 2, 8, *hw)
    dy = MG.inputs(7, 2, 4, *hw)
    kw_e, kw_r = {}, {}
    if kind == "geowizard":           # depth / normal halves, 1 image-embedding token, domain class embedding
        c = MG.inputs(2, 2, 1, 96, scale=0.5)
        cl = MG.inputs(3, 2, 10, scale=0.5)
        kw_e, kw_r = dict(class_labels=cl.to(device)), dict(class_labels=cl)
    else:
        c = MG.inputs(2, 2, ctx_tokens, 128, scale=0.5)
    y = unet(x.to(device), 999, c.to(device), **kw_e).sample
    (y * dy.to(device)).sum().backward()
    yr = unet_ref(x, 999, c, **kw_r).sample
    (yr * dy).sum().backward()
    ref = dict(unet_ref.named_parameters())
    num = den = 0.0
    worst, worst_name, missing = 0.0, None, []
    for n, p in unet.named_parameters():
        if p.grad is None:
            missing.append(n)
            continue
        gr = ref[n].grad
        e = rel_l2(p.grad, gr)
        if e > worst:
            worst, worst_name = e, n
        num += (p.grad.detach().float().cpu() - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    return dict(forward=rel_l2(y, yr), grad_global=(num / den) ** 0.5, grad_worst=worst, worst_name=worst_name,
                missing=missing, n_params=len(ref))


def run_training_step_tiny(device="cuda:0", modality="depth"):
    """Row a10, whole micro-step: rgb -> frozen VAE encode -> UNet -> x0 -> frozen VAE decode -> post-op -> task loss,
    then backward: gradients of every UNet parameter (engine autograd blocks) vs torch.autograd through the fp32
    oracle graph (training/train.py:469-563)."""
    from diffusion_e2e_ft_b200.training import LOSS_SCALE, e2e_ft_loss
    unet_ref, vae_ref = MG.build_tiny()
    unet, vae = engine_from_oracle(unet_ref, vae_ref, device)
    unet.requires_grad_(True)
    unet_ref.requires_grad_(True)
    vae_ref.requires_grad_(False)
    g = torch.Generator().manual_seed(13)
    rgb = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    ctx = torch.randn(1, 77, 128, generator=g) * 0.5
    mask = torch.rand(2, 1, 64, 64, generator=g) > 0.2
    sched_o = OP.DDIMOneStep()
    with torch.no_grad():
        lat = OP.encode_rgb(vae_ref, rgb)
    v = unet_ref(torch.cat([lat, torch.zeros_like(lat)], 1), 999, ctx.repeat(2, 1, 1)).sample
    dec = OP.decode_latent(vae_ref, sched_o.pred_original_sample(v, 999, torch.zeros_like(lat)))
    if modality == "depth":
        gt = torch.rand(2, 1, 64, 64, generator=g) * 9.9 + 0.1
        want = OP.ssi_loss(dec.mean(1, keepdim=True).clamp(-1, 1), gt, mask)
    else:
        gt = torch.nn.functional.normalize(torch.randn(2, 3, 64, 64, generator=g), dim=1)
        est = dec / (dec.norm(dim=1, keepdim=True) + 1e-5)
        want = OP.angular_loss(est.clamp(-1, 1), gt, mask)
    want.backward()
    got, _ = e2e_ft_loss(unet, vae, DDIMScheduler(), rgb.to(device), gt.to(device), mask.to(device), ctx.to(device),
                         modality)
    (got * LOSS_SCALE).backward()
    ref = dict(unet_ref.named_parameters())
    num = den = 0.0
    worst, worst_name, missing = 0.0, None, []
    for n, p in unet.named_parameters():
        if p.grad is None:
            missing.append(n)
            continue
        ge, gr = p.grad.detach().float().cpu() / LOSS_SCALE, ref[n].grad
        e = rel_l2(ge, gr)
        if e > worst and gr.norm() > 1e-3 * den ** 0.5:          # ignore parameters with a vanishing gradient
            worst, worst_name = e, n
        num += (ge - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    return dict(loss_engine=got.item(), loss_oracle=want.item(), loss_rel=abs(got.item() - want.item()) / abs(want.item()),
                grad_global=(num / den) ** 0.5, grad_worst=worst, worst_name=worst_name, missing=missing)


def run_training_loop_tiny(device="cuda:0", steps=3, lr=1e-4):
    """training/train.py:469-568 for a few iterations (depth recipe, bs=2): e2e_ft_loss -> backward -> gradient clip
    -> AdamW through FlatTrainer (flat buffers + fused CUDA optimizer) vs the oracle graph with
    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.  Reports the per-step losses and the cosine similarity /
    relative error of the accumulated parameter update."""
    from diffusion_e2e_ft_b200.training import FlatTrainer, e2e_ft_loss
    unet_ref, vae_ref = MG.build_tiny()
    unet, vae = engine_from_oracle(unet_ref, vae_ref, device)
    unet.requires_grad_(True)
    unet_ref.requires_grad_(True)
    vae_ref.requires_grad_(False)
    start = {n: p.detach().clone() for n, p in unet_ref.named_parameters()}
    opt = torch.optim.AdamW(unet_ref.parameters(), lr=lr, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)
    tr = FlatTrainer(unet, lr=lr, weight_decay=1e-2, max_grad_norm=1.0)
    g = torch.Generator().manual_seed(21)
    ctx = torch.randn(1, 77, 128, generator=g) * 0.5
    sched_o, sched = OP.DDIMOneStep(), DDIMScheduler()
    le, lo = [], []
    for _ in range(steps):
        rgb = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
        mask = torch.rand(2, 1, 64, 64, generator=g) > 0.2
        gt = torch.rand(2, 1, 64, 64, generator=g) * 9.9 + 0.1
        with torch.no_grad():
            lat = OP.encode_rgb(vae_ref, rgb)
        v = unet_ref(torch.cat([lat, torch.zeros_like(lat)], 1), 999, ctx.repeat(2, 1, 1)).sample
        dec = OP.decode_latent(vae_ref, sched_o.pred_original_sample(v, 999, torch.zeros_like(lat)))
        loss_o = OP.ssi_loss(dec.mean(1, keepdim=True).clamp(-1, 1), gt, mask)
        opt.zero_grad()
        loss_o.backward()
        torch.nn.utils.clip_grad_norm_(unet_ref.parameters(), 1.0)
        opt.step()
        loss_e, _ = e2e_ft_loss(unet, vae, sched, rgb.to(device), gt.to(device), mask.to(device), ctx.to(device), "depth")
        tr.backward(loss_e)
        tr.step()
        le.append(loss_e.item())
        lo.append(loss_o.item())
    dot = ne = no = 0.0
    for n, p in unet.named_parameters():
        de = p.detach().float().cpu() - start[n]
        do = dict(unet_ref.named_parameters())[n].detach() - start[n]
        dot += (de * do).sum().item()
        ne += de.pow(2).sum().item()
        no += do.pow(2).sum().item()
    return dict(loss_engine=le, loss_oracle=lo, update_cosine=dot / (ne * no) ** 0.5, update_norm_ratio=(ne / no) ** 0.5)


def run_checkpointing_tiny(device="cuda:0"):
    """Gradients with and without unet.enable_gradient_checkpointing() on the same weights / inputs, plus the
    checkpointed run against the oracle's autograd."""
    unet_ref, _ = MG.build_tiny()
    unet_ref.requires_grad_(True)
    x, c, dy = MG.inputs(1, 2, 8, 16, 16), MG.inputs(2, 2, 77, 128, scale=0.5), MG.inputs(7, 2, 4, 16, 16)
    (unet_ref(x, 999, c).sample * dy).sum().backward()
    grads = []
    for ck in (False, True):
        unet, _ = engine_from_oracle(unet_ref, None, device)
        unet.requires_grad_(True)
        if ck:
            unet.enable_gradient_checkpointing()
        y = unet(x.to(device), 999, c.to(device)).sample
        (y * dy.to(device)).sum().backward()
        grads.append({n: p.grad.detach().float().cpu() for n, p in unet.named_parameters()})
    num = den = num_o = den_o = 0.0
    worst, worst_name = 0.0, None
    ref = dict(unet_ref.named_parameters())
    for n in grads[0]:
        a, b = grads[0][n], grads[1][n]
        e = rel_l2(b, a)
        if e > worst:
            worst, worst_name = e, n
        num += (a - b).pow(2).sum().item()
        den += a.pow(2).sum().item()
        num_o += (b - ref[n].grad).pow(2).sum().item()
        den_o += ref[n].grad.pow(2).sum().item()
    return dict(global_rel_diff=(num / den) ** 0.5, worst_rel_diff=worst, worst_name=worst_name,
                ckpt_vs_oracle_global=(num_o / den_o) ** 0.5)


def run_training_step_geowizard_tiny(device="cuda:0"):
    """GeoWizard joint depth + normal micro-step (train_depth_normal.py:640-766): engine `e2e_ft_loss_geowizard` +
    backward vs torch.autograd through the fp32 oracle graph (joint attention, class-embedding projection, one decoder
    pass over both halves, 0.5 * SSI + angular on inverted normals)."""
    from diffusion_e2e_ft_b200.training import LOSS_SCALE, e2e_ft_loss_geowizard
    gunet_ref, vae_ref = MG.build_tiny("geowizard")
    unet, vae = engine_from_oracle(gunet_ref, vae_ref, device)
    unet.requires_grad_(True)
    gunet_ref.requires_grad_(True)
    vae_ref.requires_grad_(False)
    g = torch.Generator().manual_seed(17)
    B = 2
    rgb = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1
    emb = torch.randn(B, 1, 96, generator=g) * 0.5
    mask = torch.rand(B, 1, 64, 64, generator=g) > 0.2
    gt_d = torch.rand(B, 1, 64, 64, generator=g) * 9.9 + 0.1
    gt_n = torch.nn.functional.normalize(torch.randn(B, 3, 64, 64, generator=g), dim=1)
    sched_o = OP.DDIMOneStep()
    with torch.no_grad():
        lat = OP.encode_rgb(vae_ref, rgb)
    x = torch.cat([lat.repeat(2, 1, 1, 1), torch.zeros_like(lat).repeat(2, 1, 1, 1)], 1)
    cls = OP.geowizard_class_embedding("indoor", rgb.dtype, B)
    v = gunet_ref(x, torch.full((2 * B,), 999), encoder_hidden_states=emb.repeat(2, 1, 1), class_labels=cls).sample
    dec = OP.decode_latent(vae_ref, sched_o.pred_original_sample(v, 999, torch.zeros_like(v)))
    est_d = dec[:B].mean(1, keepdim=True).clamp(-1, 1)
    est_n = (dec[B:] / (dec[B:].norm(dim=1, keepdim=True) + 1e-5)).clamp(-1, 1)
    want = 0.5 * OP.ssi_loss(est_d, gt_d, mask) + OP.angular_loss(est_n, -gt_n, mask)
    want.backward()
    got, _, _ = e2e_ft_loss_geowizard(unet, vae, DDIMScheduler(), rgb.to(device), gt_d.to(device), gt_n.to(device),
                                      mask.to(device), emb.to(device), "indoor")
    (got * LOSS_SCALE).backward()
    ref = dict(gunet_ref.named_parameters())
    num = den = 0.0
    worst, worst_name, missing = 0.0, None, []
    for n, p in unet.named_parameters():
        if p.grad is None:
            missing.append(n)
            continue
        ge, gr = p.grad.detach().float().cpu() / LOSS_SCALE, ref[n].grad
        num += (ge - gr).pow(2).sum().item()
        den += gr.pow(2).sum().item()
    for n, p in unet.named_parameters():
        if p.grad is None:
            continue
        gr = ref[n].grad
        e = rel_l2(p.grad.detach().float().cpu() / LOSS_SCALE, gr)
        if e > worst and gr.norm() > 1e-3 * den ** 0.5:
            worst, worst_name = e, n
    return dict(loss_engine=got.item(), loss_oracle=want.item(), loss_rel=abs(got.item() - want.item()) / abs(want.item()),
                grad_global=(num / den) ** 0.5, grad_worst=worst, worst_name=worst_name, missing=missing)


@torch.no_grad()
def run_batch_consistency(device="cuda:0", res=384, batch=16):
    """bs-16 inference (BASELINE.json configs[4]) vs the same images one by one, SD-2 widths, the seeded weights of
    run_full_size.  Both sides are the engine: they differ only in tile shapes / summation order (halo vs per-tap conv,
    swapped epilogue, atomics order of the fused statistics), i.e. in which fp16 roundings flip."""
    from oracle.unet import UNet2DConditionRef, UNetConfig, seeded_init
    from oracle.vae import AutoencoderKLRef, VAEConfig
    uref = seeded_init(UNet2DConditionRef(UNetConfig()), seed=4321).eval()
    vref = seeded_init(AutoencoderKLRef(VAEConfig()), seed=99).eval()
    unet, vae = engine_from_oracle(uref, vref, device)
    g = torch.Generator().manual_seed(11)
    ete = (torch.randn(1, 2, 1024, generator=g) * 0.5).to(device)
    pipe = MarigoldPipeline(unet, vae, DDIMScheduler(), empty_text_embed=ete)
    rgb = (torch.rand(batch, 3, res, res, generator=g) * 2 - 1).to(device)
    full_d = pipe.single_infer(rgb, 1, False, noise="zeros")
    full_n = pipe.single_infer(rgb, 1, False, noise="zeros", normals=True)
    worst_d, worst_a = 0.0, 0.0
    for i in (0, batch // 2, batch - 1):
        worst_d = max(worst_d, rel_l2(full_d[i:i + 1], pipe.single_infer(rgb[i:i + 1], 1, False, noise="zeros")))
        worst_a = max(worst_a, mean_angle_deg(full_n[i:i + 1],
                                              pipe.single_infer(rgb[i:i + 1], 1, False, noise="zeros", normals=True)))
    return dict(depth_worst_vs_single=worst_d, normals_worst_angle_deg=worst_a,
                norm_err=(full_n.float().norm(dim=1) - 1).abs().max().item())


@torch.no_grad()
def run_pipeline_with_text_encoder(device="cuda:0"):
    """marigold_pipeline.py:355-369 on the engine: MarigoldPipeline.encode_empty_text -> B200CLIPTextModel (CUDA) ->
    ctx [1, 2, 128] -> the UNet's constant-context cross-attention; checked against the oracle text encoder feeding the
    oracle pipeline."""
    from diffusion_e2e_ft_b200 import B200CLIPTextModel, EmptyPromptTokenizer
    from oracle.clip_text import clip_text_forward, random_state_dict, tiny_clip_cfg
    from oracle.pipeline import DDIMOneStep
    unet_ref, vae_ref = MG.build_tiny()
    unet, vae = engine_from_oracle(unet_ref, vae_ref, device)
    cfg = tiny_clip_cfg()
    sd = random_state_dict(cfg, seed=21)
    enc = B200CLIPTextModel(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2).eval()
    enc.load_state_dict(sd)
    pipe = MarigoldPipeline(unet, vae, DDIMScheduler(), text_encoder=enc.to(device), tokenizer=EmptyPromptTokenizer())
    rgb = (torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1)
    depth = pipe.single_infer(rgb.to(device), 1, False, noise="zeros")
    ctx_ref = clip_text_forward(sd, cfg, torch.tensor([[49406, 49407]]))[0]
    want = OP.marigold_single_infer(unet_ref, vae_ref, DDIMOneStep(), rgb, ctx_ref)
    return dict(ctx_rel_l2=rel_l2(pipe.empty_text_embed, ctx_ref), depth_rel_l2=rel_l2(depth, want))


@torch.no_grad()
def run_geowizard_with_image_encoder(device="cuda:0"):
    """geowizard_pipeline.py:232-248,283-288 on the engine: rgb -> bicubic-AA resize + CLIP normalisation ->
    B200CLIPVisionModelWithProjection -> img_embed [B,1,proj] -> the joint depth+normal step; against the oracle image
    encoder feeding the oracle GeoWizard step."""
    from diffusion_e2e_ft_b200 import B200CLIPVisionModelWithProjection, CLIPImageProcessorConfig
    from oracle.clip_vision import geowizard_img_embed, random_vision_state_dict, tiny_vision_cfg
    from oracle.pipeline import DDIMOneStep
    unet_ref, vae_ref = MG.build_tiny("geowizard")
    unet, vae = engine_from_oracle(unet_ref, vae_ref, device)
    cfg = tiny_vision_cfg(projection_dim=unet_ref.config.cross_attention_dim)
    sd = random_vision_state_dict(cfg, seed=23)
    enc = B200CLIPVisionModelWithProjection(
        hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size, num_hidden_layers=cfg.num_hidden_layers,
        num_attention_heads=cfg.num_attention_heads, image_size=cfg.image_size, projection_dim=cfg.projection_dim).eval()
    enc.load_state_dict(sd)
    pipe = DepthNormalEstimationPipeline(unet, vae, DDIMScheduler(),
                                         image_encoder=enc.to(device), feature_extractor=CLIPImageProcessorConfig(cfg.image_size))
    rgb = (torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1)
    depth, normal = pipe.single_infer(rgb.to(device), 1, "indoor")
    emb_ref = geowizard_img_embed(sd, cfg, rgb)
    want_d, want_n = OP.geowizard_single_infer(unet_ref, vae_ref, DDIMOneStep(), rgb, emb_ref, domain="indoor")
    return dict(img_embed_rel_l2=rel_l2(pipe.encode_img_embed(rgb.to(device)), emb_ref),
                depth_rel_l2=rel_l2(depth, want_d), normal_angle_deg=mean_angle_deg(normal, want_n))
