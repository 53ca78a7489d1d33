"""CPU tests of the host side: C-ABI surface, state_dict compatibility with the oracle/diffusers layout,
scheduler, pipeline plumbing, loud failure without a GPU, 2-rank gloo batch sharding."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_every_declared_symbol():
    from diffusion_e2e_ft_b200 import lib
    L = lib.load()
    header = open(os.path.join(ROOT, "include", "b200_e2eft.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/b200_e2eft.h but not exported"
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    from diffusion_e2e_ft_b200.lib import ABI_VERSION
    assert L.b200_abi_version() == ABI_VERSION
    assert L.b200_geglu_block_n(2560) == 160 and L.b200_geglu_block_n(512) == 256


def test_c_abi_rejects_bad_arguments_without_launching():
    from diffusion_e2e_ft_b200 import lib
    L = lib.load()
    rc = L.b200_linear(None, 0, 0, None, 0, 0, 1, 1, 1, 1, None, 0, None, 0, 0, None, 0, 0, 0, 0, 1.0, None, 0, None, 0, 0, 0, 0, None)
    assert rc < 0 and b"null pointer" in L.b200_last_error_string()
    rc = L.b200_layer_norm(1, 0, 10, 12, 1, 1, 1e-5, 1, None)       # C not a multiple of 8
    assert rc < 0 and b"multiple of 8" in L.b200_last_error_string()


def test_engine_state_dict_matches_oracle_layout_and_loads():
    from diffusion_e2e_ft_b200 import B200UNet2DConditionModel, B200AutoencoderKL
    from oracle.unet import UNet2DConditionRef, tiny_config, seeded_init
    from oracle.vae import AutoencoderKLRef, tiny_vae_config
    cfg = tiny_config()
    ref = seeded_init(UNet2DConditionRef(cfg))
    eng = B200UNet2DConditionModel(block_out_channels=cfg.block_out_channels, attention_head_dim=cfg.attention_head_dim,
                                   cross_attention_dim=cfg.cross_attention_dim)
    missing, unexpected = eng.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    vr = seeded_init(AutoencoderKLRef(tiny_vae_config()))
    ve = B200AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    ve.load_state_dict(vr.state_dict(), strict=True)
    assert ve.config.scaling_factor == 0.18215 and ve.config["scaling_factor"] == 0.18215
    with torch.device("meta"):
        full = B200UNet2DConditionModel()
    assert sum(p.numel() for p in full.parameters()) == 865_922_244


def test_replace_unet_conv_in_semantics_on_engine_module():
    """training/util/unet_prep.py:6-21 restated verbatim must work on the drop-in module."""
    from torch.nn import Conv2d, Parameter
    from diffusion_e2e_ft_b200 import B200UNet2DConditionModel
    unet = B200UNet2DConditionModel(in_channels=4, block_out_channels=(64, 128, 256, 256),
                                    attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128)
    _weight = unet.conv_in.weight.clone().repeat((1, 2, 1, 1)) / 2
    _bias = unet.conv_in.bias.clone() / 2
    new = Conv2d(8, unet.conv_in.out_channels, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1))
    new.weight, new.bias = Parameter(_weight), Parameter(_bias)
    unet.conv_in = new
    unet.config['in_channels'] = 8
    assert unet.config.in_channels == 8 and unet.state_dict()["conv_in.weight"].shape[1] == 8


def test_no_cpu_fallback():
    from diffusion_e2e_ft_b200 import B200UNet2DConditionModel, ops
    unet = B200UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4),
                                    cross_attention_dim=128)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        unet(torch.zeros(1, 8, 8, 8), 999, torch.zeros(1, 2, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.layer_norm(torch.zeros(4, 64), torch.ones(64), torch.zeros(64))


def test_scheduler_matches_oracle_closed_form():
    from diffusion_e2e_ft_b200 import DDIMScheduler
    from oracle.pipeline import DDIMOneStep
    s, o = DDIMScheduler(), DDIMOneStep()
    s.set_timesteps(1)
    o.set_timesteps(1)
    assert s.timesteps.tolist() == o.timesteps.tolist() == [999]
    t, prev, a_t, a_prev = s.coefficients(0)
    assert t == 999 and prev == -1 and a_prev == 1.0
    assert abs(a_t - o.alphas_cumprod[999].item()) < 1e-9
    s.set_timesteps(10)
    assert s.timesteps.tolist() == [999, 899, 799, 699, 599, 499, 399, 299, 199, 99]


def test_ensembling_has_no_cpu_fallback():
    """The ensembling / post-processing ops are device kernels (csrc/postproc.cu; index bit-exactness is checked against
    the reference-run fixture in tests/test_reference_pins.py on the GPU): a CPU tensor must be refused, not routed
    through torch arithmetic."""
    import pytest
    from diffusion_e2e_ft_b200 import ensemble_depths, ensemble_normals
    preds = torch.randn(6, 3, 16, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ensemble_normals(preds)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ensemble_depths(preds[:, 0])


def test_two_rank_gloo_batch_sharding():
    """bench.py's multi-GPU path: images shard over ranks, no data-path collective; the only collective is
    the max-reduce of the timing.  Run the host logic with world_size 2 on gloo."""
    code = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
import bench
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29581", rank=int(sys.argv[1]), world_size=2)
lo, hi = bench.shard_range(16, dist.get_rank(), dist.get_world_size())
t = torch.tensor([float(hi - lo), 10.0 + dist.get_rank()])
ms = bench.max_over_ranks(10.0 + dist.get_rank(), "cpu")
cnt = torch.tensor([float(hi - lo)]); dist.all_reduce(cnt)
assert cnt.item() == 16 and ms == 11.0, (cnt, ms)
print("OK", lo, hi)
''' % ROOT
    ps = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
          for r in range(2)]
    outs = [p.communicate(timeout=180) for p in ps]
    assert all(p.returncode == 0 for p in ps), outs
    assert sorted(o[0].split()[1:] for o in outs) == [["0", "8"], ["8", "16"]]


def test_two_rank_gloo_gradient_allreduce():
    """training.allreduce_mean_: the only collective of the training path (DDP gradient average)."""
    code = r'''
import sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from diffusion_e2e_ft_b200.training import allreduce_mean_
r = int(sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29583", rank=r, world_size=2)
g = torch.full((1000,), float(r + 1))
allreduce_mean_(g)
assert torch.allclose(g, torch.full((1000,), 1.5)), g[:4]
print("OK")
''' % ROOT
    ps = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
          for r in range(2)]
    outs = [p.communicate(timeout=180) for p in ps]
    assert all(p.returncode == 0 for p in ps), outs


def test_save_pretrained_from_pretrained_round_trip(tmp_path):
    """training/train.py:322-339,610-630: the accelerate save / load hooks call `save_pretrained` / `from_pretrained`
    / `register_to_config(**config)` / `load_state_dict` on the UNet — diffusers directory layout (config.json +
    diffusion_pytorch_model.safetensors), diffusers parameter names, unknown config keys preserved."""
    import json
    from diffusion_e2e_ft_b200 import B200AutoencoderKL, B200UNet2DConditionModel
    unet = B200UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4),
                                    cross_attention_dim=128)
    d = tmp_path / "ckpt"
    unet.save_pretrained(str(d / "unet"))                                      # the save hook's call
    cfg = json.load(open(d / "unet" / "config.json"))
    assert cfg["_class_name"] == "UNet2DConditionModel" and cfg["block_out_channels"] == [64, 128, 256, 256]
    assert (d / "unet" / "diffusion_pytorch_model.safetensors").exists()
    cfg["dropout"] = 0.0                                                       # a diffusers key the engine does not model
    json.dump(cfg, open(d / "unet" / "config.json", "w"))
    load_model = B200UNet2DConditionModel.from_pretrained(str(d), subfolder="unet")    # the load hook's calls
    fresh = B200UNet2DConditionModel(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4),
                                     cross_attention_dim=128)
    fresh.register_to_config(**load_model.config)
    fresh.load_state_dict(load_model.state_dict())
    for (k, a), (_, b) in zip(unet.state_dict().items(), fresh.state_dict().items()):
        assert torch.equal(a, b), k
    assert load_model.config["_extra"]["dropout"] == 0.0 and load_model.config["cross_attention_dim"] == 128
    vae = B200AutoencoderKL(block_out_channels=(64, 64, 128, 128))
    vae.save_pretrained(str(d / "vae"), safe_serialization=False)
    v2 = B200AutoencoderKL.from_pretrained(str(d / "vae"), torch_dtype=torch.float16)
    assert v2.dtype == torch.float16 and v2.config["block_out_channels"] == (64, 64, 128, 128)
    assert torch.equal(v2.state_dict()["decoder.conv_in.weight"], vae.state_dict()["decoder.conv_in.weight"].half())


def test_product_package_never_imports_the_oracle():
    """Tier rule: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/ — the product path has
    no CPU fallback to route through.  Static check of every module of the package, plus bench.py's engine path."""
    import ast
    pkg = os.path.join(ROOT, "diffusion_e2e_ft_b200")
    for name in sorted(os.listdir(pkg)):
        if not name.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, name)).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), (name, mods)
    # bench.py: `oracle` may only be imported inside the CPU-leg functions
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    for fn in [n for n in tree.body if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").startswith("oracle"):
                assert any(k in fn.name for k in ("oracle", "cpu", "reference")), fn.name    # the CPU legs' helpers only
    for node in tree.body:                                    # and never at module level
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mods = [a.name for a in node.names] if isinstance(node, ast.Import) else [node.module or ""]
            assert not any(m.startswith("oracle") for m in mods)
