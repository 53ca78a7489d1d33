"""CPU check of the HOST-side algebra of the backward operators (diffusion_e2e_ft_b200/backward.py): operand
transposes, tap lists, head views, padding — with every CUDA kernel replaced by a plain-torch restatement of its
documented contract (include/b200_e2eft.h).  The kernels themselves are checked on the GPU (`bwd_*` checks);
this file only makes sure the compositions ask them for the right thing.  Test infrastructure only."""
import pytest

import bwd_checks
import cpu_emulation


@pytest.fixture
def emulated(monkeypatch):
    cpu_emulation.install(monkeypatch)
    monkeypatch.setattr(bwd_checks, "DEV", "cpu")


COMPOSED = ["bwd_linear", "bwd_conv_wgrad_s1", "bwd_conv_wgrad_s2", "bwd_conv_wgrad_s2_vae", "bwd_conv_wgrad_up",
            "bwd_conv_dgrad_s2_add", "bwd_conv_dgrad_s2_vae_add", "bwd_conv_dgrad_1x1_add", "bwd_attention_self",
            "bwd_attention_cross77", "bwd_attention_self_t300", "bwd_attention_cross77_t4"]


@pytest.mark.parametrize("name", COMPOSED)
def test_backward_composition(emulated, name):
    err, tol = bwd_checks.BWD_CHECKS[name]()
    assert err <= max(tol, 3e-3), f"{name}: {err:.3e}"


@pytest.mark.parametrize("padded,split", [(True, 0), (False, 64), (True, 64)])
@pytest.mark.parametrize("name", ["bwd_conv_wgrad_s1", "bwd_conv_wgrad_s2", "bwd_conv_wgrad_up"])
def test_experimental_wgrad_modes(emulated, monkeypatch, name, padded, split):
    """round-2 weight-gradient variants (padded planar operands, split-K over the batch dimension): same result"""
    from diffusion_e2e_ft_b200 import backward as bw
    monkeypatch.setattr(bw, "WGRAD_PADDED", padded)
    monkeypatch.setattr(bw, "WGRAD_SPLIT_K", split)
    monkeypatch.setattr(bw, "WGRAD_MIN_KBLOCKS", 1)
    if split:
        assert bw._split_plan(240, 128, 64)[0] > 1
    err, tol = bwd_checks.BWD_CHECKS[name]()
    assert err <= 3e-3, f"{name}: {err:.3e}"
