"""CPU tests of the host-side weight packing / operator algebra the kernels rely on, against PyTorch:
tap-list convolution semantics, the 4-phase upsample-conv decomposition, the fused 1x1 shortcut packing,
GEGLU tile interleave and the small-Cout fragment packing.  `tap_conv_reference` restates the C-ABI contract
of b200_conv2d_nhwc (include/b200_e2eft.h) in plain torch so the packing can be checked without a GPU."""
import torch
import torch.nn.functional as F

from diffusion_e2e_ft_b200 import ops
from diffusion_e2e_ft_b200.modules import Upsample2D


def tap_conv_reference(x_nhwc, wp, cout, taps, stride=1, out_hw=None, x2=None, out_mul=1, out_off=(0, 0), out=None):
    """out[n, ho*m+oy, wo*m+ox, :] = sum_t x[n, ho*s+dy_t, wo*s+dx_t, :] @ wp[:, t*Cin:(t+1)*Cin].T (+ x2 @ wp[:, T*Cin:].T),
    out-of-range pixels read as zero."""
    NB, H, W, Cin = x_nhwc.shape
    Ho, Wo = out_hw or (H, W)
    wpf = wp.float()
    res = torch.zeros(NB, Ho, Wo, cout)
    pad = 4
    xp = F.pad(x_nhwc.float().permute(0, 3, 1, 2), (pad, pad + stride * Wo, pad, pad + stride * Ho)).permute(0, 2, 3, 1)
    for t, (dy, dx) in enumerate(taps):
        ys = torch.arange(Ho) * stride + dy + pad
        xs = torch.arange(Wo) * stride + dx + pad
        patch = xp[:, ys][:, :, xs]                                   # [NB, Ho, Wo, Cin]
        res += patch @ wpf[:, t * Cin:(t + 1) * Cin].T
    if x2 is not None:
        res += x2.float() @ wpf[:, len(taps) * Cin:].T
    if out is None:
        out = torch.zeros(NB, Ho * out_mul, Wo * out_mul, cout)
    out[:, out_off[0]::out_mul, out_off[1]::out_mul] = res
    return out


def test_tap_conv_contract_equals_conv2d_for_all_padding_modes():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 9, 11, 8, generator=g)
    w = torch.randn(5, 8, 3, 3, generator=g)
    wp = ops.pack_conv(w).float()
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(tap_conv_reference(x, wp, 5, ops.TAPS3), ref, atol=2e-2)
    ref2 = F.conv2d(x.permute(0, 3, 1, 2), w, stride=2, padding=1).permute(0, 2, 3, 1)
    got2 = tap_conv_reference(x, wp, 5, ops.TAPS3, stride=2, out_hw=tuple(ref2.shape[1:3]))
    assert torch.allclose(got2, ref2, atol=2e-2)
    ref3 = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (0, 1, 0, 1)), w, stride=2).permute(0, 2, 3, 1)   # VAE downsample
    got3 = tap_conv_reference(x, wp, 5, ops.TAPS3_PAD0, stride=2, out_hw=tuple(ref3.shape[1:3]))
    assert torch.allclose(got3, ref3, atol=2e-2)


def test_fused_shortcut_packing():
    g = torch.Generator().manual_seed(1)
    x, x2 = torch.randn(1, 6, 7, 8, generator=g), torch.randn(1, 6, 7, 16, generator=g)
    w, ws = torch.randn(4, 8, 3, 3, generator=g), torch.randn(4, 16, 1, 1, generator=g)
    wp = ops.pack_conv(w, ws).float()
    ref = (F.conv2d(x.permute(0, 3, 1, 2), w, padding=1) + F.conv2d(x2.permute(0, 3, 1, 2), ws)).permute(0, 2, 3, 1)
    assert torch.allclose(tap_conv_reference(x, wp, 4, ops.TAPS3, x2=x2), ref, atol=5e-2)


def test_four_phase_upsample_conv_equals_interpolate_then_conv():
    g = torch.Generator().manual_seed(2)
    m = Upsample2D(8)
    with torch.no_grad():
        m.conv.weight.copy_(torch.randn(8, 8, 3, 3, generator=g) * 0.2)
    x = torch.randn(2, 5, 6, 8, generator=g)
    out = torch.zeros(2, 10, 12, 8)
    for (py, px), (taps, wp) in m._pack_phases().items():
        tap_conv_reference(x, wp.float(), 8, taps, out_mul=2, out_off=(py, px), out=out)
    up = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    ref = F.conv2d(up, m.conv.weight, padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(out, ref, atol=2e-2), (out - ref).abs().max()


def test_small_cout_fragment_packing_roundtrip():
    g = torch.Generator().manual_seed(3)
    w = torch.randn(3, 128, 3, 3, generator=g)
    wq = ops.pack_conv_small_cout(w).float()             # [chunk][tap][ks][n][k]
    assert wq.shape == (2, 9, 4, 8, 16)
    for (chunk, tap, ks, n, k) in [(0, 0, 0, 0, 0), (1, 4, 3, 2, 15), (0, 8, 1, 1, 7)]:
        c = chunk * 64 + ks * 16 + k
        assert abs(wq[chunk, tap, ks, n, k] - w[n, c, tap // 3, tap % 3].half().float()) < 1e-6
    assert wq[:, :, :, 3:].abs().max() == 0               # padded output channels are zero


# ---------------------------------------------------------------------------------- backward packing
def _grad_input(fn, x, dy):
    x = x.clone().requires_grad_(True)
    (fn(x) * dy).sum().backward()
    return x.grad


def test_conv_dgrad_stride1_is_a_tap_conv_with_flipped_weights():
    from diffusion_e2e_ft_b200.backward_packing import pack_conv_dgrad_s1
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(6, 8, 3, 3, generator=g) * 0.3).half().float()
    x = torch.randn(2, 8, 7, 9, generator=g)
    dy = torch.randn(2, 6, 7, 9, generator=g).half().float()
    want = _grad_input(lambda t: F.conv2d(t, w, padding=1), x, dy).permute(0, 2, 3, 1)
    wp, taps = pack_conv_dgrad_s1(w)
    got = tap_conv_reference(dy.permute(0, 2, 3, 1), wp, 8, taps)
    assert torch.allclose(got, want, atol=2e-2), (got - want).abs().max()


def test_conv_dgrad_stride2_four_phase():
    from diffusion_e2e_ft_b200.backward_packing import pack_conv_dgrad_s2
    g = torch.Generator().manual_seed(6)
    w = (torch.randn(5, 8, 3, 3, generator=g) * 0.3).half().float()
    for pad_lo, fwd in ((1, lambda t: F.conv2d(t, w, stride=2, padding=1)),
                        (0, lambda t: F.conv2d(F.pad(t, (0, 1, 0, 1)), w, stride=2))):
        x = torch.randn(2, 8, 10, 12, generator=g)
        y = fwd(x)
        dy = torch.randn(y.shape, generator=g).half().float()
        want = _grad_input(fwd, x, dy).permute(0, 2, 3, 1)
        got = torch.zeros_like(want)
        for (py, px), (wp, taps) in pack_conv_dgrad_s2(w, pad_lo).items():
            tap_conv_reference(dy.permute(0, 2, 3, 1), wp, 8, taps, out_hw=(5, 6), out_mul=2, out_off=(py, px), out=got)
        assert torch.allclose(got, want, atol=2e-2), (pad_lo, (got - want).abs().max())


def test_upsample_conv_dgrad_from_phase_weights():
    from diffusion_e2e_ft_b200.backward_packing import pack_upsample_conv_dgrad
    g = torch.Generator().manual_seed(7)
    m = Upsample2D(8)
    with torch.no_grad():
        m.conv.weight.copy_((torch.randn(8, 8, 3, 3, generator=g) * 0.2).half().float())
    x = torch.randn(1, 8, 5, 6, generator=g)
    fwd = lambda t: F.conv2d(F.interpolate(t, scale_factor=2.0, mode="nearest"), m.conv.weight, padding=1)
    dy = torch.randn(1, 8, 10, 12, generator=g).half().float()
    want = _grad_input(fwd, x, dy).permute(0, 2, 3, 1)
    got = torch.zeros_like(want)
    dy_nhwc = dy.permute(0, 2, 3, 1)
    for (py, px), (wp, taps) in pack_upsample_conv_dgrad(m._pack_phases()).items():
        got += tap_conv_reference(dy_nhwc[:, py::2, px::2].contiguous(), wp, 8, taps)
    assert torch.allclose(got, want, atol=3e-2), (got - want).abs().max()
