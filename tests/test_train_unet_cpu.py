"""Host-side wiring of the differentiable UNet (autograd blocks: saved operands, gradient routing through skip
connections / the shared time embedding, parameter-gradient layouts) on CPU, with the CUDA kernels replaced by
their plain-torch contracts (tests/cpu_emulation.py).  The real kernels run the same check in test_engine_gpu.py."""
import pytest

import cpu_emulation
import engine_checks as EC
from diffusion_e2e_ft_b200 import ops


@pytest.fixture
def emulated(monkeypatch):
    cpu_emulation.install(monkeypatch)
    monkeypatch.setattr(ops, "FUSE_GN_STATS", False)


def test_unet_backward_wiring_matches_oracle_autograd(emulated):
    r = EC.run_unet_backward_tiny(device="cpu")
    assert not r["missing"], r["missing"]
    assert r["n_params"] == 686
    assert r["forward"] <= 3e-3, r
    assert r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r


@pytest.mark.parametrize("modality,tol", [("depth", 3e-2), ("normals", 6e-2)])
def test_training_micro_step_wiring(emulated, modality, tol):
    r = EC.run_training_step_tiny(device="cpu", modality=modality)
    assert not r["missing"], r["missing"]
    assert r["loss_rel"] <= 3e-3, r
    assert r["grad_global"] <= tol and r["grad_worst"] <= 3 * tol, r
