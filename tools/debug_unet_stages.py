"""Stage-by-stage divergence report: oracle (CPU fp32) vs engine (CUDA) on the tiny UNet."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import engine_checks as EC
import make_golden as MG
from diffusion_e2e_ft_b200 import modules as M

gain = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
from oracle.unet import UNet2DConditionRef, tiny_config, seeded_init
ref = seeded_init(UNet2DConditionRef(tiny_config()), seed=1234, attn_gain=gain).eval()
eng, _ = EC.engine_from_oracle(ref, None, "cuda:0")

ref_out = {}
for name, mod in ref.named_modules():
    cls = type(mod).__name__
    if cls in ("ResnetBlock2D", "Transformer2DModel", "Downsample2D", "Upsample2D", "BasicTransformerBlock"):
        mod.register_forward_hook(lambda m, i, o, name=name: ref_out.__setitem__(name, o.detach()))

eng_out = {}
def wrap(cls, to_nchw=True):
    orig = cls.run
    def run(self, *a, **k):
        o = orig(self, *a, **k)
        eng_out[self._dbg_name] = o.detach().float().cpu()
        return o
    cls.run = run
for name, mod in eng.named_modules():
    mod._dbg_name = name
for cls in (M.ResnetBlock2D, M.Transformer2DModel, M.Downsample2D, M.Upsample2D, M.BasicTransformerBlock):
    wrap(cls)

x = MG.inputs(1, 2, 8, 16, 16); ctx = MG.inputs(2, 2, 2, 128, scale=0.5)
with torch.no_grad():
    want = ref(x, 999, ctx).sample
    got = eng(x.cuda(), 999, ctx.cuda()).sample
print("final", EC.rel_l2(got, want))
for name in ref_out:
    if name not in eng_out: continue
    r, e = ref_out[name], eng_out[name]
    if e.dim() == 4 and r.dim() == 4: e = e.permute(0, 3, 1, 2)
    if e.dim() == 2 and r.dim() == 3: e = e.view(r.shape)
    print(f"{name:55s} {EC.rel_l2(e, r):.3e}  ref_std {r.std().item():.3f}")
