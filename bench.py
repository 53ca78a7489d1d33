#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native single-step denoising engine.

    python bench.py --gpus N --steps K --warmup W            (torchrun launches N>1, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1]: marigold-e2e-ft-depth inference, bs=8 per GPU, fp16 operands,
processing_res=768, 1 denoising step, zeros noise, synthetic 3x768x768 RGB, seeded random weights.
One "step" = one `MarigoldPipeline.single_infer` over a batch (VAE encode -> UNet -> x0 -> VAE decode
-> depth post-ops).  Metric: 768x768 depth images / second (whole job, all GPUs).

  value      device-timed throughput, inputs resident in HBM
  e2e        same through the public pipeline API from pinned HOST buffers (H2D + D2H inside the timing)
  roofline   implicit-GEMM conv kernel (the dominant kernel): algorithmic FLOPs / CUDA-event time of its
             launches inside the timed region, against the measured bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference: the ORACLE (oracle/, plain PyTorch fp32 restatement of the reference's
             diffusers path, which is not installable here) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# tensor-pipe FLOPs (2*MAC) per image, SURVEY.md §8(d): UNet / VAE enc / VAE dec
TFLOP_PER_IMAGE = {768: 10.501, 512: 4.429, 384: 2.444, 256: 1.10, 128: 0.28, 64: 0.07}
METRIC = "images_per_sec_768x768_depth"


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` images for `rank` (independent images: no collective)."""
    per = (total + world - 1) // world
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def max_over_ranks(ms, device):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    hbm_gbs=d["hbm_gbs"], source="measured")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
            except (ValueError, IndexError):
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------- oracle legs
def build_oracle(seed=1234, full=True):
    """Oracle modules with cheap synthetic weights (timing only): built on the meta device, then filled
    with U(-1/sqrt(fan_in), 1/sqrt(fan_in)) in place (default nn init of 950 M parameters takes ~45 s)."""
    import torch
    from oracle.unet import UNet2DConditionRef, UNetConfig, tiny_config
    from oracle.vae import AutoencoderKLRef, VAEConfig, tiny_vae_config
    g = torch.Generator().manual_seed(seed)
    with torch.device("meta"):
        unet = UNet2DConditionRef(UNetConfig() if full else tiny_config())
        vae = AutoencoderKLRef(VAEConfig() if full else tiny_vae_config())
    for m in (unet, vae):
        m.to_empty(device="cpu")
        with torch.no_grad():
            for name, p in m.named_parameters():
                if p.dim() >= 2:
                    p.uniform_(-1.0, 1.0, generator=g).mul_(1.0 / (p[0].numel() ** 0.5))
                elif "norm" in name and name.endswith("weight"):
                    p.fill_(1.0)
                else:
                    p.zero_()
        m.eval()
    return unet, vae


def oracle_step(unet, vae, res, batch=1):
    import torch
    from oracle import pipeline as OP
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(batch, 3, res, res, generator=g) * 2 - 1
    ete = torch.randn(1, 2, unet.config.cross_attention_dim, generator=g) * 0.5
    t0 = time.perf_counter()
    with torch.no_grad():
        OP.marigold_single_infer(unet, vae, OP.DDIMOneStep(), rgb, ete)
    return time.perf_counter() - t0


def pick_cpu_threads(unet, vae):
    """Host thread count for the oracle: all cores is often NOT fastest on a 128-way shared host
    (oversubscribed oneDNN thread pools), so calibrate on a tiny problem and keep the best."""
    import torch
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (avail, 64, 32, 16, 8) if 1 <= c <= avail}, reverse=True)
    best, best_t = cands[-1], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        oracle_step(unet, vae, 64)
        t = oracle_step(unet, vae, 64)
        if t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def pick_cpu_res(unet, vae, budget_s_per_step):
    """Largest resolution whose oracle step fits the per-step budget, from a small calibration run."""
    t = oracle_step(unet, vae, 128)
    tf_s = TFLOP_PER_IMAGE[128] / t                           # conservative: small problems run slower
    for res in (768, 512, 384, 256):
        if TFLOP_PER_IMAGE[res] / tf_s <= budget_s_per_step:
            return res
    return 128


def run_reference(args):
    """The reference's own (CPU, fp32, PyTorch) path on the host cores: the oracle restatement, since
    diffusers==0.30.2 cannot be installed offline (DESIGN.md §Reference arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    unet, vae = build_oracle()
    cores = pick_cpu_threads(unet, vae)
    res = pick_cpu_res(unet, vae, 120.0 / max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        oracle_step(unet, vae, res)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(unet, vae, res)
    dt = time.perf_counter() - t0
    eq_images = args.steps * TFLOP_PER_IMAGE[res] / TFLOP_PER_IMAGE[768]
    value = eq_images / dt
    sample = (f"{args.steps} x oracle single_infer of 1 image at {res}x{res} fp32 on {cores} host threads; "
              f"converted to 768x768-equivalent images by the tensor-FLOP ratio {TFLOP_PER_IMAGE[res]}/{TFLOP_PER_IMAGE[768]}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "marigold-e2e-ft-depth single_infer, 1 step, zeros noise, 768x768 (CPU oracle)",
                   "global_batch": 1},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_leg(budget_s=20.0):
    unet, vae = build_oracle()
    cores = pick_cpu_threads(unet, vae)
    res = pick_cpu_res(unet, vae, budget_s)
    t = oracle_step(unet, vae, res)
    value = (TFLOP_PER_IMAGE[res] / TFLOP_PER_IMAGE[768]) / t
    return {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"1 oracle single_infer (fp32 PyTorch CPU restatement of the diffusers path) of 1 image at "
                      f"{res}x{res} in {t:.1f} s, scaled to 768x768-equivalent images by tensor-FLOP ratio"}


# --------------------------------------------------------------------------------------------- engine
def build_engine(device, stream_dtype, module_dtype, seed=1234):
    import torch
    from diffusion_e2e_ft_b200 import (B200AutoencoderKL, B200UNet2DConditionModel, DDIMScheduler, MarigoldPipeline)
    torch.manual_seed(seed)
    with torch.device(device):
        unet = B200UNet2DConditionModel(stream_dtype=stream_dtype)
        vae = B200AutoencoderKL(stream_dtype=stream_dtype)
    unet.to(module_dtype).eval().requires_grad_(False)
    vae.to(module_dtype).eval().requires_grad_(False)
    ete = (torch.randn(1, 2, 1024, device=device) * 0.5).to(module_dtype)
    return MarigoldPipeline(unet, vae, DDIMScheduler(), empty_text_embed=ete)


def run_engine(args):
    import torch
    import torch.distributed as dist
    from diffusion_e2e_ft_b200 import ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bs, res = args.batch, args.res
    sdt = torch.float32 if args.stream == "fp32" else torch.float16
    pipe = build_engine(dev, sdt, torch.float16)
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    host_rgb = (torch.rand(bs, 3, res, res, generator=g) * 2 - 1).half().pin_memory()
    host_out = torch.empty(bs, 1, res, res, dtype=torch.float16).pin_memory()
    dev_rgb = host_rgb.to(dev)

    def step_resident():
        return pipe.single_infer(dev_rgb, 1, False, noise="zeros")

    def step_e2e():
        x = host_rgb.to(dev, non_blocking=True)
        y = pipe.single_infer(x, 1, False, noise="zeros")
        host_out.copy_(y, non_blocking=True)
        return y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, time_conv=False):
        ops.STATS.reset()
        ops.STATS.time_kind = ("gemm" if args.dump_shapes else "conv") if time_conv else None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ops.STATS.time_kind = None
        return max_over_ranks(e0.elapsed_time(e1), dev)

    # ---- timed region A: the step as a user runs it (CUDA-graph replay inside the pipeline)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = ops.STATS.launches
    total_flops = sum(ops.STATS.flops.values())
    flops_by_kind = dict(ops.STATS.flops)

    # ---- timed region B (roofline leg): the same step launched eagerly so every implicit-GEMM conv
    # launch can be bracketed with CUDA events on the launching stream
    pipe.use_cuda_graph = False
    step_resident()
    rsteps = min(args.steps, 3)
    eager_ms = timed(step_resident, rsteps, time_conv=True)
    stats = ops.STATS
    conv_ms = sum(e[0].elapsed_time(e[1]) for e in stats.events)
    conv_flops = sum(e[2] for e in stats.events)
    if args.dump_shapes and rank == 0:
        agg = {}
        for e in stats.lin_events:
            a = agg.setdefault(str(e[3]), [0, 0.0, 0])
            a[0] += 1; a[1] += e[0].elapsed_time(e[1]); a[2] += e[2]
        rows = sorted(((k, n, ms_, fl / (ms_ / 1e3) / 1e12) for k, (n, ms_, fl) in agg.items()), key=lambda r: -r[2])
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'linear_shapes.txt'), 'w') as f:
            f.write('(B,M,N,K,act,residual,out) launches total_ms TFLOP/s\n')
            for k, n, ms_, tf in rows:
                f.write(f'{k:55s} {n:4d} {ms_:9.3f} {tf:8.1f}\n')
        agg = {}
        for e in stats.events:
            a = agg.setdefault(str(e[3]), [0, 0.0, 0])
            a[0] += 1; a[1] += e[0].elapsed_time(e[1]); a[2] += e[2]
        rows = sorted(((k, n, ms_, fl / (ms_ / 1e3) / 1e12) for k, (n, ms_, fl) in agg.items()), key=lambda r: -r[2])
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'conv_shapes.txt'), 'w') as f:
            f.write('(NB,H,W,Cin,C2,Cout,taps,stride,out) launches total_ms TFLOP/s\n')
            for k, n, ms_, tf in rows:
                f.write(f'{k:55s} {n:4d} {ms_:9.3f} {tf:8.1f}\n')
    n_conv = len(stats.events)
    # per-op breakdown of one eager step (CUDA events around every op)
    ops.STATS.time_all = True
    ops.STATS.op_events = []
    step_resident()
    torch.cuda.synchronize()
    ops.STATS.time_all = False
    breakdown = {}
    for name, a, b2 in ops.STATS.op_events:
        breakdown[name] = breakdown.get(name, 0.0) + a.elapsed_time(b2)
    ops.STATS.op_events = []
    pipe.use_cuda_graph = True

    # ---- UNet-only forward (part of the headline metric triple), graph-replayed
    lat = torch.randn(bs, 8, res // 8, res // 8, device=dev, dtype=torch.float16)
    ctx = pipe.empty_text_embed.repeat(bs, 1, 1)
    with torch.no_grad():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                pipe.unet(lat, 999, encoder_hidden_states=ctx)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.STATS.reset()
        ug = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ug):
            pipe.unet(lat, 999, encoder_hidden_states=ctx)
        unet_flops = sum(ops.STATS.flops.values())
        ug.replay()
        unet_ms = timed(ug.replay, args.steps) / args.steps

    for _ in range(2):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    peaks = measured_peaks()
    images = bs * world * args.steps
    value = images / (ms / 1e3)
    e2e_value = images / (e2e_ms / 1e3)
    out = None
    if rank == 0:
        conv_tf = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        ms_eager = eager_ms / rsteps
        peak = peaks["tflops_sustained"]
        traffic, traffic_of = None, None
        tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_of = (f"one ncu --set full capture of {tj.get('kernel')} on {tj.get('shape')}: "
                          f"{tj.get('algorithmic_flops', 0) / 1e12:.3f} TFLOP, algorithmic bytes >= "
                          f"{tj.get('algorithmic_bytes_min', 0) / 1e9:.3f} GB ({tj.get('source')})")
        out = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": f"marigold-e2e-ft-depth single_infer bs={bs}/GPU {res}x{res}, 1 step, zeros noise "
                                   f"(BASELINE.json configs[1])",
                       "global_batch": bs * world, "parallelism": f"dp{world} (independent images, no collective)",
                       "stream_dtype": args.stream, "operands": "fp16 x fp16 -> fp32 accumulate",
                       "l2": "per-step working set (GBs of activations + 1.9 GB weights) >> 126 MB L2; no flush needed"},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": host_rgb.numel() * 2,
                    "d2h_bytes_per_step": host_out.numel() * 2, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "unet_fwd_ms": unet_ms,
            "unet_tensor_frac": unet_flops / (unet_ms / 1e3) / 1e12 / peak,
            "step_tensor_tflops": total_flops / (ms / 1e3) / 1e12,
            "step_tensor_frac": total_flops / (ms / 1e3) / 1e12 / peak,
            "flops_per_step": flops_by_kind,
            "breakdown_ms_eager_step": {k: round(v, 2) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1])},
            "roofline": {"kernel": "gemm_conv_kernel (implicit-GEMM conv3x3, tcgen05+TMA)", "bound": "tensor",
                         "achieved": conv_tf, "peak": peak, "unit": "TFLOP/s", "frac": conv_tf / peak,
                         "traffic": traffic, "traffic_of": traffic_of, "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                         "launches_timed": n_conv, "avg_launch_ms": conv_ms / max(1, n_conv),
                         "share_of_step": (conv_ms / rsteps) / (ms / args.steps),
                         "timed_in": f"{rsteps} eagerly launched steps of the same workload ({ms_eager:.1f} ms/step eager)"},
            "clocks": clocks,
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_baseline_leg()
            except Exception as e:  # noqa: BLE001
                out["cpu_baseline"] = {"error": repr(e)[:200]}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--res", type=int, default=768)
    ap.add_argument("--stream", default="fp32", choices=["fp32", "fp16"], help="residual-stream dtype in the engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-shapes", action="store_true", help="write per-shape conv timings to gpurun_out/")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_engine(args)


if __name__ == "__main__":
    main()
