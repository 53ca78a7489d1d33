"""Time the CLIP encoders on the engine kernels (random-init weights of the real architectures):
text encoder (23 x 1024) on the empty prompt (2 tokens) and on 77 tokens; image encoder (ViT-L/14) incl. the
bicubic-AA resize + normalisation, batch 1 / 4 / 8 from 768x768 inputs.  CUDA events, median of 10 after 3 warm-ups."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_e2e_ft_b200 import B200CLIPTextModel, B200CLIPVisionModelWithProjection, CLIPImageProcessorConfig  # noqa: E402


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = "cuda:0"
    out = {}
    with torch.device(dev):
        txt = B200CLIPTextModel().half().eval()
        vis = B200CLIPVisionModelWithProjection().half().eval()
    ids2 = torch.tensor([[49406, 49407]], device=dev)
    ids77 = torch.randint(1000, 40000, (1, 77), device=dev)
    out["text_2_tokens_ms"] = timeit(lambda: txt(ids2))
    out["text_77_tokens_ms"] = timeit(lambda: txt(ids77), n=5, warm=2)
    fe = CLIPImageProcessorConfig(224)
    for b in (1, 4, 8):
        rgb = torch.rand(b, 3, 768, 768, device=dev) * 2 - 1
        out[f"image_bs{b}_preprocess_ms"] = timeit(lambda: vis.preprocess(rgb, fe))
        x = vis.preprocess(rgb, fe).half()
        ms = timeit(lambda: vis(x))
        flops = b * (24 * (2 * 257 * 1024 * (3 * 1024 + 1024 + 2 * 4096) + 4 * 257 * 257 * 1024) + 2 * 256 * 588 * 1024)
        out[f"image_bs{b}_encoder_ms"] = ms
        out[f"image_bs{b}_tflops"] = flops / ms / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
