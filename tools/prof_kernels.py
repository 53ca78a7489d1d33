"""Tiny launcher for ncu captures of single kernels at production shapes.
    python tools/prof_kernels.py conv128|conv320|conv1280|attn|gn|linear"""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusion_e2e_ft_b200 import ops
what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
from diffusion_e2e_ft_b200 import lib as _l
if len(sys.argv) > 3:
    _l.load().b200_debug_set_flags(int(sys.argv[3]))
if len(sys.argv) > 4:
    _l.load().b200_debug_force_block_n(int(sys.argv[4]))
if len(sys.argv) > 5:
    _l.load().b200_debug_set_swap(int(sys.argv[5]))
if os.environ.get("B200_ATT_VERSION"):
    _l.load().b200_debug_set_attention_version(int(os.environ["B200_ATT_VERSION"]))
if os.environ.get("B200_HALO"):
    _l.load().b200_debug_set_halo(int(os.environ["B200_HALO"]))
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).half().to(dev)
if what.startswith("conv"):
    cfg = {"conv128": (4, 768, 768, 128, 128), "conv128n8": (8, 768, 768, 128, 128), "conv128n2": (2, 768, 768, 128, 128), "conv256": (8, 384, 384, 256, 256), "conv512": (8, 192, 192, 512, 512),
           "conv320": (8, 96, 96, 320, 320), "conv640": (8, 48, 48, 640, 640), "conv1280": (8, 24, 24, 1280, 1280),
           "conv1280s": (8, 12, 12, 1280, 1280), "conv2560": (8, 24, 24, 2560, 1280)}[what]
    NB, H, W, Cin, Cout = cfg
    x = r(NB, H, W, Cin); w = ops.pack_conv(r(Cout, Cin, 3, 3, sc=1 / math.sqrt(9 * Cin))); b = torch.zeros(Cout, device=dev)
    st = True if os.environ.get("B200_STATS") else None
    fn = lambda: ops.conv2d(x, w, Cout, bias=b, stats=st)
    flops = 2 * NB * H * W * Cout * 9 * Cin
elif what.startswith("res"):          # res16 / res32: 128->128 768^2 conv with residual, fp16 / fp32 stream
    NB, H, W, Cin, Cout = 4, 768, 768, 128, 128
    odt = torch.float16 if what == "res16" else torch.float32
    x = r(NB, H, W, Cin); w = ops.pack_conv(r(Cout, Cin, 3, 3, sc=1 / math.sqrt(9 * Cin))); b = torch.zeros(Cout, device=dev)
    resid = r(NB, H, W, Cout).to(odt)
    fn = lambda: ops.conv2d(x, w, Cout, bias=b, residual=resid, out_dtype=odt)
    flops = 2 * NB * H * W * Cout * 9 * Cin
elif what == "attn2304":             # level-1 self-attention: 10 heads, 2304 tokens
    qkv = r(8, 2304, 1920)
    fn = lambda: ops.attention_d64(qkv[..., :640], qkv[..., 640:1280], qkv[..., 1280:], 10, 0.125)
    flops = 4 * 8 * 10 * 2304 * 2304 * 64
elif what == "attn":
    qkv = r(8, 9216, 960)
    fn = lambda: ops.attention_d64(qkv[..., :320], qkv[..., 320:640], qkv[..., 640:], 5, 0.125)
    flops = 4 * 8 * 5 * 9216 * 9216 * 64
elif what == "gn":
    x = r(4, 768, 768, 128); gm = torch.ones(128, device=dev); bt = torch.zeros(128, device=dev)
    fn = lambda: ops.group_norm(x, gm, bt, 1e-6)
    flops = 0
elif what == "gn32":                  # fp32 stream in, fp16 out: the VAE's dominant GroupNorm shape (6 B / element)
    x = r(4, 768, 768, 128).float(); gm = torch.ones(128, device=dev); bt = torch.zeros(128, device=dev)
    fn = lambda: ops.group_norm(x, gm, bt, 1e-6)
    flops = 0
elif what == "lin320":                # K = 320 GEMM with fp32 residual / output (attention out-proj, proj_out)
    a = r(73728, 320); w = r(320, 320, sc=0.05); b = torch.zeros(320, device=dev)
    resid = r(73728, 320).float()
    fn = lambda: ops.linear(a, w, b, residual=resid, out_dtype=torch.float32)
    flops = 2 * 73728 * 320 * 320
elif what.startswith("ln"):            # ln320 / ln640 / ln1280: the UNet's LayerNorms, fp32 stream in, fp16 out (6 B / element)
    C = int(what[2:]); rows = {320: 73728, 640: 18432, 1280: 4608}[C] * (8 if C != 320 else 1)
    x = r(rows, C).float(); gm = torch.ones(C, device=dev); bt = torch.zeros(C, device=dev)
    fn = lambda: ops.layer_norm(x, gm, bt)
    flops = 0; nbytes = rows * C * 6
elif what == "softmax":               # the VAE attention's score matrix of 2 images: fp32 in, fp16 out
    sm = r(2 * 9216, 9216).float()
    fn = lambda: ops.softmax_rows(sm, 0.044)
    flops = 0; nbytes = 2 * 9216 * 9216 * 6
elif what == "geglu":
    pass
elif what == "linear":
    a = r(73728, 320); w = r(2560, 320, sc=0.05); b = torch.zeros(2560, device=dev)
    fn = lambda: ops.linear(a, w, b)
    flops = 2 * 73728 * 320 * 2560
if what == "geglu":
    a = r(73728, 320); w0 = r(2560, 320, sc=0.05); b0 = torch.zeros(2560, device=dev)
    wg, bg = ops.pack_geglu(w0, b0)
    fn = lambda: ops.linear(a, wg, bg, act=ops.ACT_GEGLU)
    flops = 2 * 73728 * 320 * 2560
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{what}: {ms:.3f} ms/iter  {flops / ms / 1e9:.1f} TFLOP/s" + (f"  {nbytes / ms / 1e6:.0f} GB/s" if "nbytes" in dir() else ""))
