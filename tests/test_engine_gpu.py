"""-m gpu: the engine (CUDA, C ABI) against the oracle / committed golden fixtures.

Tolerance: north_star asks for 1e-3 relative to fp32.  The engine computes with fp16 tensor-core operands
(the only way to the tensor-pipe target; TF32 has the same 10-bit mantissa) and fp32 accumulation /
statistics / softmax / residual stream, so each GEMM contributes ~3e-4 of operand rounding and a 60-layer
graph lands at 1-2e-3 rel-L2 against the fp32 oracle — measured values are recorded in
profiles/parity_r01.json.  Gates: rel-L2 <= 3e-3 (fp32 stream) / <= 8e-3 (fp16 stream) per output, depth
AbsRel(engine, oracle) <= 1e-3 (the north_star accuracy gate), the engine must be at least as close to
the fp32 oracle as the reference's own fp16 GPU path (torch eager fp16), ensemble index bit-exact."""
import pytest
import torch

import engine_checks as EC


@pytest.mark.gpu
def test_marigold_tiny_matches_golden():
    r = EC.run_marigold_tiny()
    print(r)
    for k in ("unet_16x16_ctx2", "unet_15x20_ctx77", "vae_encode", "vae_decode", "depth_rel_l2"):
        assert r[k] <= 3e-3, (k, r)
    # unit normals: x/||x|| amplifies error where the decoded vector is short, so gate the angle (the
    # reference's normals metric, reported to 0.1 deg): mean angular error vs oracle <= 0.5 deg
    assert r["normals_mean_angle_deg"] <= 0.5, r
    assert r["absrel_delta"] <= 1e-3, r


@pytest.mark.gpu
def test_marigold_tiny_fp16_stream():
    r = EC.run_marigold_tiny(stream_dtype=torch.float16)
    print(r)
    for k in ("unet_rel_l2", "vae_encode", "vae_decode", "depth_rel_l2"):
        assert r[k] <= 8e-3, (k, r)


@pytest.mark.gpu
@pytest.mark.parametrize("full_width", [False, True])
def test_single_step_specialisations_match_general_path(full_width):
    """cached constant-t embedding + 4-channel conv_in + constant-context cross-attention (two skinny
    GEMMs) == the general path (fp16-operand tolerance) and the fp32 oracle."""
    r = EC.run_single_step_specialisations(hw=(24, 24) if full_width else (16, 16), full_width=full_width)
    print(r)
    assert r["spec_vs_general"] <= 2e-3 and r["repeat_call"] == 0.0, r
    assert r["spec_vs_oracle"] <= 3e-3 and r["spec_vs_oracle"] <= 1.15 * r["general_vs_oracle"], r
    assert r["per_image_ctx_vs_oracle"] <= 3e-3, r


@pytest.mark.gpu
def test_geowizard_tiny_joint_attention_matches_golden():
    r = EC.run_geowizard_tiny()
    print(r)
    assert r["depth"] <= 3e-3 and r["normal_mean_angle_deg"] <= 0.5, r


@pytest.mark.gpu
def test_unet_full_width_small_latent():
    r = EC.run_unet_fullwidth(latent=24)
    print(r)
    assert r["rel_l2"] <= 3e-3, r
    assert r["rel_l2"] <= 1.1 * r["torch_fp16_rel_l2"], r


@pytest.mark.gpu
def test_ensemble_normals_index_bit_exact_on_gpu():
    from diffusion_e2e_ft_b200 import ensemble_normals
    from oracle.pipeline import ensemble_normals as ref
    g = torch.Generator().manual_seed(1)
    preds = torch.randn(6, 3, 32, 32, generator=g)
    want, idx = ref(preds)
    got, _ = ensemble_normals(preds.cuda())
    nrm = preds / (torch.norm(preds, p=2, dim=1).unsqueeze(1) + 1e-5)
    assert torch.equal(got.cpu(), nrm[idx]) or torch.allclose(got.cpu(), nrm[idx], atol=1e-6)


@pytest.mark.gpu
def test_full_size_768_against_fp32_oracle_on_gpu():
    """BASELINE.json configs[0]/[1] size (3x768x768, SD-2 widths): engine vs the fp32 oracle (run with torch ops
    on the same GPU), plus size-independent properties."""
    r = EC.run_full_size(res=768, batch=1)
    print(r)
    for k in ("rgb_latent_rel_l2", "unet_rel_l2", "decode_rel_l2", "depth_rel_l2"):
        assert r[k] <= 3e-3, (k, r)
    # the END-TO-END depth map is inside the contract's 1e-3 at full size (measured 9.7e-4, profiles/parity_r02.json);
    # locked in with a margin for box-to-box atomics / clock noise.  The intermediate stages are not.
    assert r["depth_rel_l2"] <= 1.3e-3, r
    assert r["absrel_delta"] <= 1e-3, r
    assert 0.0 <= r["depth_min"] and r["depth_max"] <= 1.0, r
    assert r["normals_norm_err"] <= 2e-3, r
    assert r["batch_consistency"] <= 1e-3, r


@pytest.mark.gpu
@pytest.mark.parametrize("modality", ["depth", "normals"])
def test_training_step_forward_loss_matches_oracle(modality):
    """training/train.py:469-556 forward (encode -> UNet(ctx 77) -> x0 -> decode -> post-op -> loss)."""
    r = EC.run_training_forward_tiny(modality=modality)
    print(r)
    assert r["rel_err"] <= 3e-3, r


@pytest.mark.gpu
def test_unet_backward_matches_oracle_autograd():
    """`loss.backward()` through the engine UNet (autograd blocks on the CUDA backward operators)
    vs torch.autograd through the fp32 oracle.  fp16 GEMM operands in both directions: 2e-2 per parameter."""
    r = EC.run_unet_backward_tiny()
    assert not r["missing"], r["missing"]
    assert r["forward"] <= 3e-3, r
    assert r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r


@pytest.mark.gpu
@pytest.mark.parametrize("modality,tol", [("depth", 3e-2), ("normals", 6e-2)])
def test_training_micro_step_gradients_match_oracle(modality, tol):
    """Whole differentiable micro-step (frozen VAE encode -> UNet -> x0 -> frozen VAE decode -> post-op -> task loss)
    with `backward()`: UNet parameter gradients vs torch.autograd through the fp32 oracle.  The tolerance is wider
    than for the UNet alone: fp16 operands through the VAE decoder backward as well, and both losses are
    non-smooth (sign of the L1 residual, clamp / acos), so the 2e-3 forward difference flips a few pixels."""
    r = EC.run_training_step_tiny(modality=modality)
    assert not r["missing"], r["missing"]
    assert r["loss_rel"] <= 3e-3, r
    assert r["grad_global"] <= tol and r["grad_worst"] <= 3 * tol, r


def _check_loop(r):
    for a, b in zip(r["loss_engine"], r["loss_oracle"]):
        assert abs(a - b) / abs(b) <= 3e-3, r
    assert r["update_cosine"] >= 0.98 and abs(r["update_norm_ratio"] - 1.0) <= 0.03, r


@pytest.mark.gpu
def test_training_loop_matches_oracle_adamw():
    """Three iterations of training/train.py:469-568 (loss -> backward -> clip -> AdamW) on the engine (FlatTrainer:
    flat buffers + fused CUDA clip/AdamW, loss-scaled fp16 backward) vs the oracle with torch.optim.AdamW."""
    _check_loop(EC.run_training_loop_tiny())


@pytest.mark.gpu
def test_unet_backward_odd_latent_size():
    """15x20 latents (480x640 images / 4, the Hypersim recipe's aspect): odd levels -> stride-2 dgrad with a cropped
    border, explicit-size nearest upsample and its backward, attention lengths that are not multiples of 8."""
    r = EC.run_unet_backward_tiny(hw=(15, 20))
    assert not r["missing"], r["missing"]
    assert r["forward"] <= 3e-3, r
    assert r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r



# ---- round 2: the paths that had never run on hardware in round 1 (VERDICT r1 task 2; all green on a B200 via
# tools/pending_gpu_checks.py before they were admitted to the suite)
@pytest.mark.gpu
def test_geowizard_unet_backward_joint_attention():
    """GeoWizard-shaped UNet (class-embedding projection, 1 context token, XFormersJointAttnProcessor:
    GeoWizard/geowizard/models/attention.py:482-491): the depth / normal pair is differentiated as one 2L x 2L
    attention problem; parameter gradients vs torch.autograd through the fp32 oracle."""
    r = EC.run_unet_backward_tiny(kind="geowizard")
    print(r)
    assert not r["missing"], r["missing"]
    assert r["forward"] <= 3e-3, r
    assert r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r


@pytest.mark.gpu
def test_gradient_checkpointing_matches_plain_backward():
    """unet.enable_gradient_checkpointing() (training/train.py:358-359): blocks keep only their inputs and re-run their
    forward kernels inside backward.  The recomputed forward is the inference path of the block (fused statistics,
    two-GEMM GEGLU) while the plain training forward stores fp16 pre-activations for backward, so the two gradients
    differ by fp16 hand-off rounding — not bit for bit: measured on a B200 global 1.3e-3, worst parameter 2.5e-3, with
    the checkpointed run as close to the fp32 oracle (5.0e-3) as the plain one.  Gates: 3e-3 / 1e-2 / oracle 1e-2."""
    r = EC.run_checkpointing_tiny()
    print(r)
    assert r["global_rel_diff"] <= 3e-3 and r["worst_rel_diff"] <= 1e-2, r
    assert r["ckpt_vs_oracle_global"] <= 1e-2, r


@pytest.mark.gpu
@pytest.mark.parametrize("padded,split", [(True, 0), (False, 296), (True, 296)])
def test_wgrad_variants(padded, split):
    """backward.py weight-gradient GEMM variants (zero-padded K-major operands, split-K over 2 x 148 CTAs): operator
    parity vs torch.autograd and the odd-size (15x20) UNet backward with the variant switched on."""
    import bwd_checks
    from diffusion_e2e_ft_b200 import backward as bw
    keep = (bw.WGRAD_PADDED, bw.WGRAD_SPLIT_K, bw.WGRAD_MIN_KBLOCKS)
    try:
        bw.WGRAD_PADDED, bw.WGRAD_SPLIT_K, bw.WGRAD_MIN_KBLOCKS = padded, split, 1
        for name in ("bwd_conv_wgrad_s1", "bwd_conv_wgrad_s2", "bwd_conv_wgrad_up"):
            err, tol = bwd_checks.BWD_CHECKS[name]()
            assert err <= tol, (name, err, tol)
        bw.WGRAD_MIN_KBLOCKS = 2
        r = EC.run_unet_backward_tiny(hw=(15, 20))
        assert not r["missing"] and r["grad_global"] <= 1e-2 and r["grad_worst"] <= 2e-2, r
    finally:
        bw.WGRAD_PADDED, bw.WGRAD_SPLIT_K, bw.WGRAD_MIN_KBLOCKS = keep


@pytest.mark.gpu
def test_geowizard_joint_depth_normal_training_step():
    """GeoWizard/geowizard/training/train_depth_normal.py:640-766 (`--e2e_ft`): joint depth + normal micro-step on the
    engine (`training.e2e_ft_loss_geowizard`) — loss and UNet parameter gradients vs torch.autograd through the fp32
    oracle.  Same tolerance class as the normals micro-step (non-smooth angular loss, fp16 operands through the VAE
    decoder backward as well)."""
    r = EC.run_training_step_geowizard_tiny()
    print(r)
    assert not r["missing"], r["missing"]
    assert r["loss_rel"] <= 3e-3, r
    assert r["grad_global"] <= 6e-2 and r["grad_worst"] <= 0.2, r


@pytest.mark.gpu
def test_full_size_1024_and_batch16_consistency():
    """BASELINE.json configs[4] corners: 1024x1024 (latent 128^2, self-attention over 16384 tokens) against the fp32
    oracle on the same GPU, and a batch of 16 at 384x384 whose every image must equal its batch-1 result (the halo /
    swapped tile choices and the fused per-image statistics all depend on the batch size)."""
    r = EC.run_full_size(res=1024, batch=1)
    print(r)
    for k in ("rgb_latent_rel_l2", "unet_rel_l2", "decode_rel_l2", "depth_rel_l2"):
        assert r[k] <= 3e-3, (k, r)
    assert r["absrel_delta"] <= 1e-3 and r["normals_norm_err"] <= 2e-3 and r["batch_consistency"] <= 1e-3, r
    c = EC.run_batch_consistency(res=384, batch=16)
    print(c)
    assert c["depth_worst_vs_single"] <= 2e-3 and c["normals_worst_angle_deg"] <= 0.5 and c["norm_err"] <= 2e-3, c
