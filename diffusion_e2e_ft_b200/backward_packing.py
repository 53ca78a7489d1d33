"""Host-side weight packing for the BACKWARD convolutions (round-2 groundwork for SURVEY.md §8 row a10).

The data-gradient of every convolution on the hot path is again a tap-list convolution, i.e. it runs on the
existing `b200_conv2d_nhwc` kernel with re-packed weights — no new tensor-core kernel is needed for dgrad:

  * 3x3 / stride 1 / pad 1:   dX = tapconv(dY, flip(W)^T, taps (ky-1, kx-1))
  * 3x3 / stride 2 (pad 1 or the VAE's (0,1,0,1)): four output-parity phases, each a 1-, 2- or 4-tap conv of
    dY written with out_mul = 2 (the mirror image of the 4-phase upsample conv of the forward pass)
  * nearest-2x upsample + 3x3 conv (forward = four 2x2 phase convs): dX = sum over the four phases of the
    transposed phase conv applied to the stride-2 slice of dY.

These functions only build weights / tap lists; they are verified on CPU against torch.autograd in
tests/test_packing_cpu.py through the plain-torch restatement of the kernel contract.  They are NOT yet wired
into the modules: the weight-gradient GEMMs, attention / GroupNorm / LayerNorm backward kernels are missing.
"""
import torch

from .ops import F16


def pack_conv_dgrad_s1(w):
    """w [Cout, Cin, 3, 3] -> (packed [Cin, 9*Cout] fp16, taps) with dX = tapconv(dY, packed, Cin, taps)."""
    wt = w.detach().flip(2, 3).permute(1, 2, 3, 0)                     # [Cin, ky', kx', Cout], spatially flipped
    taps = [(ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    return wt.reshape(w.shape[1], -1).to(F16).contiguous(), taps


def pack_conv_dgrad_s2(w, pad_lo=1):
    """Stride-2 conv (input index = 2*o + k - pad_lo): per input parity (py, px) the taps that reach it.
    Returns {(py, px): (packed [Cin, T*Cout] fp16, taps over dY)}; phase result goes to dX[py::2, px::2]."""
    def axis(par):
        # input i = 2a + par receives dY[o] * w[k] with 2o + k - pad_lo = i  ->  k = par + pad_lo (mod 2)
        out = []
        for k in range(3):
            if (par + pad_lo - k) % 2 == 0:
                out.append(((par + pad_lo - k) // 2, k))                # (offset of o relative to a, kernel index)
        return out
    res = {}
    for py in (0, 1):
        for px in (0, 1):
            taps, mats = [], []
            for dy, ky in axis(py):
                for dx, kx in axis(px):
                    taps.append((dy, dx))
                    mats.append(w.detach()[:, :, ky, kx].t())          # [Cin, Cout]
            res[(py, px)] = (torch.stack(mats, dim=1).reshape(w.shape[1], -1).to(F16).contiguous(), taps)
    return res


def pack_upsample_conv_dgrad(phases):
    """`phases` = Upsample2D._pack_phases() of the forward: {(py,px): (taps, wp [Cout, 4*Cin])}.
    Returns {(py,px): (packed [Cin, 4*Cout] fp16, taps)}: dX += tapconv(dY[py::2, px::2], packed, Cin, taps)."""
    res = {}
    for key, (taps, wp) in phases.items():
        cout = wp.shape[0]
        cin = wp.shape[1] // len(taps)
        w4 = wp.float().reshape(cout, len(taps), cin)                   # [Cout, T, Cin]
        wt = w4.permute(2, 1, 0).reshape(cin, -1)                       # [Cin, T*Cout]
        res[key] = (wt.to(F16).contiguous(), [(-dy, -dx) for dy, dx in taps])
    return res
