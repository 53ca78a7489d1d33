"""CPU check of the HOST-side algebra of the backward operators (diffusion_e2e_ft_b200/backward.py): operand
transposes, tap lists, head views, padding — with every CUDA kernel replaced by a plain-torch restatement of its
documented contract (include/b200_e2eft.h).  The kernels themselves are checked on the GPU (`bwd_*` checks);
this file only makes sure the compositions ask them for the right thing.  Test infrastructure only."""
import pytest
import torch
import torch.nn.functional as F

import bwd_checks
from diffusion_e2e_ft_b200 import ops
from test_packing_cpu import tap_conv_reference

F16, F32 = torch.float16, torch.float32


def _emu_linear(a, w, bias=None, residual=None, out=None, out_dtype=F16, act=0, alpha=1.0, **kw):
    assert a.stride(-1) == 1 and w.stride(-1) == 1 and a.dtype == F16 and w.dtype == F16
    assert a.stride(-2) % 8 == 0 and w.stride(-2) % 8 == 0, "TMA: 16-byte row pitch"
    if a.dim() == 3:
        assert a.stride(0) % 8 == 0
    if w.dim() == 3:
        assert w.stride(0) % 8 == 0
    r = a.float() @ w.float().transpose(-1, -2) * alpha
    if bias is not None:
        r = r + bias
    if residual is not None:
        r = r + residual.float()
    if out is None:
        return r.to(out_dtype)
    out.copy_(r.to(out.dtype))
    return out


def _emu_gather(x, out_hw=None, stride=1, up=1, off=(0, 0)):
    NB, H, W, C = x.shape
    Ho, Wo = out_hw if out_hw is not None else (H, W)
    xu = x.float().permute(0, 3, 1, 2)
    if up == 2:
        xu = F.interpolate(xu, scale_factor=2.0, mode="nearest")
    pad = 4
    xp = F.pad(xu, (pad, pad + stride * Wo, pad, pad + stride * Ho))
    ys = torch.arange(Ho) * stride + off[0] + pad
    xs = torch.arange(Wo) * stride + off[1] + pad
    ref = xp[:, :, ys][:, :, :, xs].permute(1, 0, 2, 3).reshape(C, -1)
    P = NB * Ho * Wo
    out = torch.zeros(C, ops._ru8(P), dtype=F16)
    out[:, :P] = ref.half()
    return out


def _emu_conv2d(x, wp, cout, bias=None, taps=ops.TAPS3, stride=1, out_hw=None, residual=None, out=None,
                out_dtype=F16, out_mul=1, out_off=(0, 0), **kw):
    NB, H, W, _ = x.shape
    Ho, Wo = out_hw if out_hw is not None else (H, W)
    r = tap_conv_reference(x, wp, cout, taps, stride=stride, out_hw=(Ho, Wo))
    if out is None:
        out = torch.zeros(NB, Ho * out_mul, Wo * out_mul, cout, dtype=out_dtype)
    sl = (slice(None), slice(out_off[0], None, out_mul), slice(out_off[1], None, out_mul))
    if residual is not None:
        r = r + residual[sl].float()
    out[sl] = r.to(out.dtype)
    return out


def _emu_softmax_rows(s, scale, cols=None):
    cols = cols or s.shape[-1]
    p = torch.zeros(s.shape, dtype=F16)
    p[..., :cols] = torch.softmax(s[..., :cols] * scale, dim=-1).half()
    return p


def _emu_softmax_bwd_rows(p, dp, scale, cols=None):
    cols = cols or p.shape[-1]
    pf, d = p[..., :cols].float(), dp[..., :cols]
    ds = torch.zeros(p.shape, dtype=F16)
    ds[..., :cols] = (scale * pf * (d - (pf * d).sum(-1, keepdim=True))).half()
    return ds


def _emu_col_sum(x, out=None):
    r = x.float().sum(0)
    return r if out is None else out.add_(r)


@pytest.fixture
def emulated(monkeypatch):
    monkeypatch.setattr(bwd_checks, "DEV", "cpu")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for name, fn in dict(linear=_emu_linear, gather_planar=_emu_gather, conv2d=_emu_conv2d,
                         softmax_rows=_emu_softmax_rows, softmax_bwd_rows=_emu_softmax_bwd_rows,
                         col_sum=_emu_col_sum).items():
        monkeypatch.setattr(ops, name, fn)


COMPOSED = ["bwd_linear", "bwd_conv_wgrad_s1", "bwd_conv_wgrad_s2", "bwd_conv_wgrad_s2_vae", "bwd_conv_wgrad_up",
            "bwd_conv_dgrad_s2_add", "bwd_conv_dgrad_s2_vae_add", "bwd_conv_dgrad_1x1_add", "bwd_attention_self",
            "bwd_attention_cross77"]


@pytest.mark.parametrize("name", COMPOSED)
def test_backward_composition(emulated, name):
    err, tol = bwd_checks.BWD_CHECKS[name]()
    assert err <= max(tol, 3e-3), f"{name}: {err:.3e}"
