"""GPU checks of the backward-pass operators (row a10) against torch.autograd on the same fp16-rounded inputs.
Each check returns (rel-L2 error, tolerance).  Registered into kernel_checks.CHECKS (`bwd_*`)."""
import math

import torch
import torch.nn.functional as F

from diffusion_e2e_ft_b200 import backward as bw
from diffusion_e2e_ft_b200 import ops

DEV = "cuda"
F16, F32 = torch.float16, torch.float32


def _rand(*shape, seed=0, scale=1.0, dtype=F16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def _worst(pairs):
    return max(rel_l2(a, b) for a, b in pairs)


# ----------------------------------------------------------------------------------------------- gather / sums
def check_gather_planar(in_f32=False, stride=1, up=1, off=(0, 0), NB=2, H=9, W=11, C=40, sliced=False):
    full = _rand(NB, H, W, C + 24, seed=1, dtype=F32 if in_f32 else F16)
    x = full[..., 8:8 + C] if sliced else full[..., :C].contiguous()
    Ho, Wo = (H * up - 1) // stride + 1, (W * up - 1) // stride + 1
    got = ops.gather_planar(x, out_hw=(Ho, Wo), stride=stride, up=up, off=off)
    torch.cuda.synchronize()
    xu = x.float().permute(0, 3, 1, 2)
    if up == 2:
        xu = F.interpolate(xu, scale_factor=2.0, mode="nearest")
    pad = 4
    xp = F.pad(xu, (pad, pad, pad, pad))
    ys = torch.arange(Ho, device=DEV) * stride + off[0] + pad
    xs = torch.arange(Wo, device=DEV) * stride + off[1] + pad
    ref = xp[:, :, ys][:, :, :, xs]                                   # [NB, C, Ho, Wo]
    ref = ref.permute(1, 0, 2, 3).reshape(C, -1).half().float()
    P = NB * Ho * Wo
    assert got.shape == (C, (P + 7) // 8 * 8)
    padz = got[:, P:].abs().max().item() if got.shape[1] > P else 0.0
    err = (got[:, :P].float() - ref).abs().max().item() + padz
    return err, 0.0


def check_col_sum(rows=1000, C=72, in_f32=False):
    x = _rand(rows, C + 8, seed=2, dtype=F32 if in_f32 else F16)[:, :C]
    got = ops.col_sum(x)
    torch.cuda.synchronize()
    return rel_l2(got, x.float().sum(0)), 2e-6


# ------------------------------------------------------------------------------------------------------ linear
def check_linear_bwd(M=300, N=320, K=640, seed=3):
    a = _rand(M, K, seed=seed)
    w = _rand(N, K, seed=seed + 1, scale=1 / math.sqrt(K))
    dy = _rand(M, N, seed=seed + 2)
    add = _rand(M, K, seed=seed + 3, dtype=F32)
    da, dw, db = bw.linear_bwd(a, w, dy, da_dtype=F32, da_add=add)
    torch.cuda.synchronize()
    af, wf = a.float().requires_grad_(True), w.float().requires_grad_(True)
    b = torch.zeros(N, device=DEV, requires_grad=True)
    (F.linear(af, wf, b) * dy.float()).sum().backward()
    return _worst([(da, af.grad + add), (dw, wf.grad), (db, b.grad)]), 3e-5


# -------------------------------------------------------------------------------------------------------- conv
def check_conv_wgrad(kind="s1", NB=2, H=12, W=10, Cin=64, Cout=128, seed=5):
    x = _rand(NB, H, W, Cin, seed=seed)
    w = _rand(Cout, Cin, 3, 3, seed=seed + 1, scale=1 / math.sqrt(9 * Cin)).float().requires_grad_(True)
    b = torch.zeros(Cout, device=DEV, requires_grad=True)
    xn = x.float().permute(0, 3, 1, 2)
    if kind == "s1":
        y = F.conv2d(xn, w, b, padding=1)
        args = dict(taps=ops.TAPS3)
    elif kind == "s2":
        y = F.conv2d(xn, w, b, stride=2, padding=1)
        args = dict(taps=ops.TAPS3, stride=2)
    elif kind == "s2_vae":
        y = F.conv2d(F.pad(xn, (0, 1, 0, 1)), w, b, stride=2)
        args = dict(taps=ops.TAPS3_PAD0, stride=2)
    else:
        y = F.conv2d(F.interpolate(xn, scale_factor=2.0, mode="nearest"), w, b, padding=1)
        args = dict(taps=ops.TAPS3, up=2)
    dy = _rand(*y.permute(0, 2, 3, 1).shape, seed=seed + 2)
    (y * dy.float().permute(0, 3, 1, 2)).sum().backward()
    dwp, db = bw.conv_wgrad(x, dy, **args)
    torch.cuda.synchronize()
    return _worst([(bw.unpack_conv_grad(dwp, Cin), w.grad), (db, b.grad)]), 3e-5


def check_conv_dgrad_add(kind="s2", NB=2, H=12, W=16, Cin=64, Cout=128, seed=7):
    k = 1 if kind == "1x1" else 3
    w = _rand(Cout, Cin, k, k, seed=seed, scale=1 / math.sqrt(k * k * Cin))
    x = _rand(NB, Cin, H, W, seed=seed + 1, dtype=F32).requires_grad_(True)
    if kind == "1x1":
        y = F.conv2d(x, w.float())
    elif kind == "s2":
        y = F.conv2d(x, w.float(), stride=2, padding=1)
    else:
        y = F.conv2d(F.pad(x, (0, 1, 0, 1)), w.float(), stride=2)
    dy = _rand(*y.permute(0, 2, 3, 1).shape, seed=seed + 2)
    add = _rand(NB, H, W, Cin, seed=seed + 3, dtype=F32)
    (y * dy.float().permute(0, 3, 1, 2)).sum().backward()
    got = bw.conv_dgrad(dy, w, Cin, kind=kind, add=add)
    torch.cuda.synchronize()
    return rel_l2(got, x.grad.permute(0, 2, 3, 1) + add), 3e-5


# ------------------------------------------------------------------------------------------------------- norms
def check_group_norm_bwd(C1=128, C2=0, in_f32=True, silu=True, add=True, out_f32=True, NB=2, H=9, W=7, seed=11):
    dt = F32 if in_f32 else F16
    odt = F32 if out_f32 else F16
    xs = [_rand(NB, H, W, C1, seed=seed, dtype=dt) * 1.5 + 0.3]
    if C2:
        xs.append(_rand(NB, H, W, C2, seed=seed + 1, dtype=dt) * 0.7 - 0.2)
    C = C1 + C2
    gamma = _rand(C, seed=seed + 2, dtype=F32) * 0.3 + 1.0
    beta = _rand(C, seed=seed + 3, dtype=F32) * 0.2
    dy = _rand(NB, H, W, C, seed=seed + 4)
    adds = [_rand(*x.shape, seed=seed + 5 + i, dtype=odt) for i, x in enumerate(xs)] if add else None
    mr = ops.group_norm_mean_rstd(xs[0], 1e-5, 32, xs[1] if C2 else None)
    dxs, dg, db = ops.group_norm_bwd(xs, dy, mr, gamma, beta, 32, silu, adds, odt)
    torch.cuda.synchronize()
    xr = [x.float().requires_grad_(True) for x in xs]
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = F.group_norm(torch.cat(xr, dim=3).permute(0, 3, 1, 2), 32, g, b, 1e-5)
    y = F.silu(y) if silu else y
    (y * dy.float().permute(0, 3, 1, 2)).sum().backward()
    pairs = [(dg, g.grad), (db, b.grad)]
    for i in range(len(xs)):
        pairs.append((dxs[i], xr[i].grad + (adds[i].float() if add else 0)))
    return _worst(pairs), (2e-5 if out_f32 else 1e-3)


def check_layer_norm_bwd(rows=777, C=320, in_f32=True, add=True, seed=13):
    x = _rand(rows, C, seed=seed, dtype=F32 if in_f32 else F16) * 2 + 0.5
    gamma = _rand(C, seed=seed + 1, dtype=F32) * 0.3 + 1.0
    beta = _rand(C, seed=seed + 2, dtype=F32) * 0.2
    dy = _rand(rows, C, seed=seed + 3)
    addt = _rand(rows, C, seed=seed + 4, dtype=F32) if add else None
    dx, dg, db = ops.layer_norm_bwd(x, dy, gamma, 1e-5, addt)
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    (F.layer_norm(xr, (C,), g, b, 1e-5) * dy.float()).sum().backward()
    return _worst([(dx, xr.grad + (addt if add else 0)), (dg, g.grad), (db, b.grad)]), 2e-5


def check_softmax_bwd(rows=300, cols=77, scale=0.125, seed=15):
    ld = (cols + 7) // 8 * 8
    s = torch.zeros(rows, ld, device=DEV)
    s[:, :cols] = _rand(rows, cols, seed=seed, dtype=F32) * 4
    dp = torch.zeros(rows, ld, device=DEV)
    dp[:, :cols] = _rand(rows, cols, seed=seed + 1, dtype=F32)
    p = ops.softmax_rows(s, scale, cols=cols)
    ds = ops.softmax_bwd_rows(p, dp, scale, cols=cols)
    torch.cuda.synchronize()
    sr = s[:, :cols].clone().requires_grad_(True)
    pr = torch.softmax(sr * scale, dim=-1)
    (pr * dp[:, :cols]).sum().backward()
    padz = ds[:, cols:].abs().max().item() if ld > cols else 0.0
    return rel_l2(ds[:, :cols], sr.grad) + padz, 2e-3      # P and dS are fp16


def check_act_bwd(act):
    x = _rand(1000, 96, seed=17) * 2
    dy = _rand(1000, 96, seed=18)
    got = ops.act_bwd(x, dy, act)
    torch.cuda.synchronize()
    xr = x.float().requires_grad_(True)
    ((F.silu(xr) if act == ops.ACT_SILU else F.gelu(xr)) * dy.float()).sum().backward()
    return rel_l2(got, xr.grad), 1e-3


def check_geglu_bwd(rows=500, inner=192):
    hg = _rand(rows, 2 * inner, seed=19) * 1.5
    dy = _rand(rows, inner, seed=20)
    got = ops.geglu_bwd(hg, dy)
    torch.cuda.synchronize()
    r = hg.float().requires_grad_(True)
    h, g = r.chunk(2, dim=-1)
    (h * F.gelu(g) * dy.float()).sum().backward()
    return rel_l2(got, r.grad), 1e-3


# --------------------------------------------------------------------------------------------------- attention
def check_attention_bwd(B=2, T=192, Tk=192, heads=2, fused_qkv=True, seed=23):
    C = heads * 64
    scale = 1 / 8.0
    if fused_qkv and T == Tk:
        qkv = _rand(B, T, 3 * C, seed=seed)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = _rand(B, T, C, seed=seed)
        kv = _rand(B, Tk, 2 * C, seed=seed + 1)
        k, v = kv[..., :C], kv[..., C:]
    do = _rand(B, T, C, seed=seed + 2)
    dq, dk, dv = bw.attention_bwd(q, k, v, do, heads, scale)
    torch.cuda.synchronize()

    def split(t):
        return t.float().unflatten(-1, (heads, 64)).transpose(1, 2)
    qr, kr, vr = (split(t).detach().requires_grad_(True) for t in (q, k, v))
    o = torch.softmax(qr @ kr.transpose(-1, -2) * scale, dim=-1) @ vr
    (o * split(do)).sum().backward()
    back = lambda t: t.transpose(1, 2).flatten(2)
    return _worst([(dq, back(qr.grad)), (dk, back(kr.grad)), (dv, back(vr.grad))]), 3e-3


# ------------------------------------------------------------------------------------------ losses / post-op
def _ssi_ref(p, y, m):
    mf = m.float()
    a00, a01, a11 = (mf * p * p).sum((1, 2, 3)), (mf * p).sum((1, 2, 3)), mf.sum((1, 2, 3))
    b0, b1 = (mf * p * y).sum((1, 2, 3)), (mf * y).sum((1, 2, 3))
    det = a00 * a11 - a01 * a01
    s = (a11 * b0 - a01 * b1) / det
    t = (-a01 * b0 + a00 * b1) / det
    r = s.view(-1, 1, 1, 1) * p + t.view(-1, 1, 1, 1) - y
    return r.abs()[m].mean()


def check_ssi_loss_bwd(B=3, H=40, W=56, seed=41):
    g = torch.Generator(device="cpu").manual_seed(seed)
    p = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).to(DEV)
    y = (torch.rand(B, 1, H, W, generator=g) * 9 + 0.5).to(DEV)
    m = (torch.rand(B, 1, H, W, generator=g) > 0.3).to(DEV)
    go = torch.tensor(512.0, device=DEV)
    got = ops.ssi_loss_bwd(p, y, m, go)
    torch.cuda.synchronize()
    pr = p.double().requires_grad_(True)
    (_ssi_ref(pr, y.double(), m) * 512.0).backward()
    return rel_l2(got, pr.grad), 1e-4


def check_angular_loss_bwd(B=2, H=40, W=56, seed=43):
    g = torch.Generator(device="cpu").manual_seed(seed)
    p = F.normalize(torch.randn(B, 3, H, W, generator=g), dim=1).to(DEV)
    y = F.normalize(torch.randn(B, 3, H, W, generator=g), dim=1).to(DEV)
    m = (torch.rand(B, 1, H, W, generator=g) > 0.3).to(DEV)
    go = torch.tensor(64.0, device=DEV)
    got = ops.angular_loss_bwd(p, y, m, go)
    torch.cuda.synchronize()
    pr = p.double().requires_grad_(True)
    (torch.acos((pr * y.double()).sum(1).clamp(-1, 1))[m[:, 0]].mean() * 64.0).backward()
    return rel_l2(got, pr.grad), 1e-4


def check_decode_post_bwd(normals, B=2, H=24, W=40, seed=45):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(B, 3, H, W, generator=g) * (0.4 if normals else 0.8)).to(DEV)
    dy = torch.randn(B, 3 if normals else 1, H, W, generator=g).to(DEV)
    got = ops.decode_post_bwd(x, dy, normals)
    torch.cuda.synchronize()
    xr = x.clone().requires_grad_(True)
    if normals:
        est = (xr / (xr.norm(dim=1, keepdim=True) + 1e-5)).clamp(-1, 1)
    else:
        est = xr.mean(1, keepdim=True).clamp(-1, 1)
    (est * dy).sum().backward()
    fwd = rel_l2(ops.decode_post(x, normals=normals, training=True), est)
    return rel_l2(got, xr.grad) + fwd, 1e-5


def check_upsample_nearest_bwd(NB=2, H=8, W=10, OH=15, OW=20, C=64, add=True, seed=47):
    dy = _rand(NB, OH, OW, C, seed=seed, dtype=F32)
    addt = _rand(NB, H, W, C, seed=seed + 1, dtype=F32) if add else None
    got = ops.upsample_nearest_bwd(dy, (H, W), addt)
    x = _rand(NB, H, W, C, seed=seed + 2)
    up = ops.upsample_nearest(x, (OH, OW))
    torch.cuda.synchronize()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref_up = F.interpolate(xr, size=(OH, OW), mode="nearest")
    (ref_up * dy.permute(0, 3, 1, 2)).sum().backward()
    fwd = (up.float() - ref_up.detach().permute(0, 2, 3, 1)).abs().max().item()
    return rel_l2(got, xr.grad.permute(0, 2, 3, 1) + (addt if add else 0)) + fwd, 1e-6


BWD_CHECKS = {
    "bwd_upsample_nearest_8x10_to_15x20": check_upsample_nearest_bwd,
    "bwd_upsample_nearest_6x19_to_11x38": lambda: check_upsample_nearest_bwd(1, 6, 19, 11, 38, 128, False),
    "bwd_ssi_loss": check_ssi_loss_bwd,
    "bwd_angular_loss": check_angular_loss_bwd,
    "bwd_decode_post_depth": lambda: check_decode_post_bwd(False),
    "bwd_decode_post_normals": lambda: check_decode_post_bwd(True),
    "bwd_gather_transpose": lambda: check_gather_planar(),
    "bwd_gather_f32_shift": lambda: check_gather_planar(in_f32=True, off=(-1, 1)),
    "bwd_gather_stride2": lambda: check_gather_planar(stride=2, off=(-1, -1)),
    "bwd_gather_stride2_pad0": lambda: check_gather_planar(stride=2, off=(2, 2), H=10, W=12),
    "bwd_gather_up2": lambda: check_gather_planar(up=2, off=(1, -1)),
    "bwd_gather_channel_slice": lambda: check_gather_planar(sliced=True, off=(0, 1)),
    "bwd_col_sum_f16": lambda: check_col_sum(),
    "bwd_col_sum_f32": lambda: check_col_sum(rows=5000, C=320, in_f32=True),
    "bwd_linear": lambda: check_linear_bwd(),
    "bwd_linear_big": lambda: check_linear_bwd(M=2304, N=1280, K=320, seed=31),
    "bwd_conv_wgrad_s1": lambda: check_conv_wgrad("s1"),
    "bwd_conv_wgrad_s2": lambda: check_conv_wgrad("s2"),
    "bwd_conv_wgrad_s2_vae": lambda: check_conv_wgrad("s2_vae"),
    "bwd_conv_wgrad_up": lambda: check_conv_wgrad("up"),
    "bwd_conv_dgrad_s2_add": lambda: check_conv_dgrad_add("s2"),
    "bwd_conv_dgrad_s2_vae_add": lambda: check_conv_dgrad_add("s2_vae"),
    "bwd_conv_dgrad_1x1_add": lambda: check_conv_dgrad_add("1x1"),
    "bwd_group_norm_f32": lambda: check_group_norm_bwd(),
    "bwd_group_norm_concat_f16": lambda: check_group_norm_bwd(C1=320, C2=64, in_f32=False),
    "bwd_group_norm_nosilu_f16out": lambda: check_group_norm_bwd(C1=512, silu=False, add=False, out_f32=False),
    "bwd_group_norm_wide": lambda: check_group_norm_bwd(C1=1280, C2=1280, in_f32=False, H=4, W=6),
    "bwd_layer_norm": lambda: check_layer_norm_bwd(),
    "bwd_layer_norm_1280_f16": lambda: check_layer_norm_bwd(rows=300, C=1280, in_f32=False, add=False),
    "bwd_softmax": lambda: check_softmax_bwd(),
    "bwd_softmax_wide": lambda: check_softmax_bwd(rows=64, cols=4800),
    "bwd_silu": lambda: check_act_bwd(ops.ACT_SILU),
    "bwd_gelu": lambda: check_act_bwd(ops.ACT_GELU),
    "bwd_geglu": lambda: check_geglu_bwd(),
    "bwd_attention_self": lambda: check_attention_bwd(),
    "bwd_attention_self_t300": lambda: check_attention_bwd(B=1, T=300, Tk=300, heads=2, seed=29),
    "bwd_attention_cross77_t4": lambda: check_attention_bwd(B=2, T=4, Tk=77, heads=4, fused_qkv=False, seed=33),
    "bwd_attention_cross77": lambda: check_attention_bwd(T=256, Tk=77, heads=5, fused_qkv=False),
}


if __name__ == "__main__":
    import sys
    import traceback
    names = [n for n in BWD_CHECKS if not sys.argv[1:] or any(a in n for a in sys.argv[1:])]
    bad = 0
    for n in names:
        try:
            err, tol = BWD_CHECKS[n]()
            ok = err == err and err <= tol
            print(f"{'ok  ' if ok else 'FAIL'} {n:34s} err {err:.3e} tol {tol:.1e}", flush=True)
            bad += not ok
        except Exception:
            bad += 1
            print(f"EXC  {n}\n{traceback.format_exc()}", flush=True)
    print(f"{len(names) - bad}/{len(names)} passed")
