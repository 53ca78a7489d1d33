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


@pytest.mark.parametrize("hw", [(16, 16), (15, 20), (11, 38)])
def test_unet_backward_wiring_matches_oracle_autograd(emulated, hw):
    r = EC.run_unet_backward_tiny(device="cpu", hw=hw)
    assert not r["missing"], r["missing"]
    assert r["n_params"] == 686
    assert r["forward"] <= 3e-3, r
    assert r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r


def test_single_step_specialisations_wiring(emulated):
    """SURVEY.md §8 f1 host logic (folded A / VW tables, cached embedding, 4-channel conv_in) with emulated kernels."""
    r = EC.run_single_step_specialisations(device="cpu", hw=(8, 8))
    assert r["spec_vs_general"] <= 2e-3 and r["spec_vs_oracle"] <= 3e-3 and r["repeat_call"] == 0.0, r
    assert r["per_image_ctx_vs_oracle"] <= 3e-3, r


@pytest.mark.parametrize("modality,tol", [("depth", 3e-2), ("normals", 6e-2)])
def test_training_micro_step_wiring(emulated, modality, tol):
    r = EC.run_training_step_tiny(device="cpu", modality=modality)
    assert not r["missing"], r["missing"]
    assert r["loss_rel"] <= 3e-3, r
    assert r["grad_global"] <= tol and r["grad_worst"] <= 3 * tol, r


def test_geowizard_joint_training_step_wiring(emulated):
    """train_depth_normal.py:640-766 on the engine (kernels emulated): joint depth + normal loss and its backward."""
    r = EC.run_training_step_geowizard_tiny(device="cpu")
    assert not r["missing"], r["missing"]
    assert r["loss_rel"] <= 3e-3, r
    assert r["grad_global"] <= 6e-2 and r["grad_worst"] <= 0.2, r


def test_training_loop_wiring(emulated):
    """flat parameter / gradient buffers, loss scaling, packed-weight cache invalidation after the optimizer step"""
    from test_engine_gpu import _check_loop
    _check_loop(EC.run_training_loop_tiny(device="cpu"))


def test_two_rank_gloo_data_parallel_training_step():
    """DDP semantics of training/train.py (accelerate): each rank runs the micro-step on its own images, the flat
    gradient buffer is averaged over the ranks (one all-reduce), every rank applies the same clip + AdamW update.
    world_size 2 on gloo with the kernels emulated: the ranks' parameters must stay identical, and differ from a
    run without the exchange."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import sys, torch, torch.distributed as dist
torch.set_num_threads(4)
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests"); sys.path.insert(0, %r + "/tests/golden")
import cpu_emulation, engine_checks as E, make_golden as MG
from diffusion_e2e_ft_b200 import ops, DDIMScheduler
from diffusion_e2e_ft_b200.training import FlatTrainer, e2e_ft_loss
class MP:
    def setattr(self, o, n, v): setattr(o, n, v)
cpu_emulation.install(MP()); ops.FUSE_GN_STATS = False
r = int(sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29587", rank=r, world_size=2)
unet_ref, vae_ref = MG.build_tiny()
unet, vae = E.engine_from_oracle(unet_ref, vae_ref, "cpu")
unet.requires_grad_(True)
tr = FlatTrainer(unet, lr=1e-4, bucket_mb=0.25, accumulation_steps=2, overlap=True)   # many buckets: async all-reduce per bucket, launched from the backward hooks
assert len(tr._buckets) > 4
late = sum((p.numel() + 3) // 4 * 4 for n, p in unet.named_parameters() if any(k in n for k in tr.LATE_GRAD_KEYS))
assert any(b["hi"] == late for b in tr._buckets)        # late-gradient parameters lead the flat order, own bucket(s)
g = torch.Generator().manual_seed(100 + r)                     # different images on each rank
ctx = MG.inputs(5, 1, 77, 128, scale=0.5)
def batch():
    return (torch.rand(1, 3, 64, 64, generator=g) * 2 - 1, torch.rand(1, 1, 64, 64, generator=g) * 9.9 + 0.1,
            torch.rand(1, 1, 64, 64, generator=g) > 0.2)
micro = [batch(), batch()]                                     # one accumulation window = two micro-batches per rank
before = tr.flat_param.clone()
# the local (pre-exchange) gradients of both micro-steps, on a scratch pass without the exchange
local = []
for rgb, gt, mask in micro:
    loss, _ = e2e_ft_loss(unet, vae, DDIMScheduler(), rgb, gt, mask, ctx, "depth")
    tr.backward(loss, sync=False)
    local.append(tr.flat_grad.clone())
    tr.flat_grad.zero_()
tr._micro = 0
# the real window: micro-step 1 accumulates locally (no exchange), micro-step 2 all-reduces bucket by bucket while
# backward is still running, then clip + AdamW
stepped = []
for rgb, gt, mask in micro:
    loss, _ = e2e_ft_loss(unet, vae, DDIMScheduler(), rgb, gt, mask, ctx, "depth")
    if not stepped:
        try:
            tr.step()
            raise SystemExit("step() inside an accumulation window must raise")
        except RuntimeError:
            pass
    stepped.append(tr.micro_step(loss))
assert stepped == [False, True], stepped
both = [torch.zeros_like(tr.flat_param) for _ in range(2)]
dist.all_gather(both, tr.flat_param)
assert torch.equal(both[0], both[1]), "ranks diverged"
mine = local[0] + local[1]                                     # each already carries loss_scale / accumulation_steps
grads = [torch.zeros_like(mine) for _ in range(2)]
dist.all_gather(grads, mine)
assert not torch.allclose(grads[0], grads[1]), "ranks saw the same data"
# expected update: AdamW (emulated kernel contract) on the mean over ranks of the accumulated gradient
mean = (grads[0] + grads[1]) / 2
m, v = torch.zeros_like(mean), torch.zeros_like(mean)
cpu_emulation.adamw_step(before, mean, m, v, 1, lr=1e-4, grad_norm_sq_t=cpu_emulation.grad_norm_sq(mean),
                         max_grad_norm=1.0, grad_unscale=1.0 / tr.loss_scale())
assert torch.allclose(before, tr.flat_param, rtol=0, atol=2e-7), (before - tr.flat_param).abs().max()
assert tr.applied_steps() == 1 and tr.skipped_steps() == 0
# a non-finite gradient: the step is skipped (parameters and moments untouched) and the loss scale halves
keep, scale0 = tr.flat_param.clone(), tr.loss_scale()
tr.flat_grad[7] = float("inf"); tr._micro = 2; tr._synced = True
tr.step()
assert torch.equal(keep, tr.flat_param) and tr.skipped_steps() == 1 and tr.loss_scale() == scale0 / 2
# an all-zero gradient (every validity mask empty, train.py:503): skipped too, scale unchanged
tr._micro = 2; tr._synced = True
tr.step()
assert torch.equal(keep, tr.flat_param) and tr.skipped_steps() == 2 and tr.loss_scale() == scale0 / 2
print("OK", float(loss))
''' % (root, root, root)
    ps = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
          for r in range(2)]
    outs = [p.communicate(timeout=600) for p in ps]
    assert all(p.returncode == 0 for p in ps), outs
    assert all(o[0].startswith("OK") for o in outs), outs


def test_geowizard_joint_attention_backward_wiring(emulated):
    """GeoWizard-shaped UNet (class-embedding projection, 1 context token, joint self-attention over the depth /
    normal pair): the pair is differentiated as one 2L x 2L attention problem."""
    r = EC.run_unet_backward_tiny(device="cpu", kind="geowizard")
    assert not r["missing"], r["missing"]
    assert r["forward"] <= 3e-3, r
    assert r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r


def test_gradient_checkpointing_gives_identical_gradients(emulated):
    """unet.enable_gradient_checkpointing() (training/train.py:358-359): resnet / transformer blocks keep only their
    inputs and re-run their forward kernels inside backward — same kernels on the same inputs, so bit-identical."""
    import make_golden as MG
    import torch
    unet_ref, _ = MG.build_tiny()
    grads = []
    for ck in (False, True):
        unet, _ = EC.engine_from_oracle(unet_ref, None, "cpu")
        unet.requires_grad_(True)
        if ck:
            unet.enable_gradient_checkpointing()
        y = unet(MG.inputs(1, 2, 8, 16, 16), 999, MG.inputs(2, 2, 77, 128, scale=0.5)).sample
        (y * MG.inputs(7, 2, 4, 16, 16)).sum().backward()
        grads.append({n: p.grad.clone() for n, p in unet.named_parameters()})
    assert all(torch.equal(grads[0][n], grads[1][n]) for n in grads[0])
