#!/usr/bin/env python
"""bench.py -- panoptic masks/sec of the LDMSeg sampling hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --steps K --warmup W    (the reference-equivalent CPU path)

A "step" is one pass of the hot path over one per-GPU batch of synthetic 512x512 RGB:
AutoencoderKL encode -> 50-step DDIM loop (UNet forward + scheduler step) -> seg-VAE decode to
panoptic ids.  N=1 runs BASELINE.json configs[1] (SD-1.5 UNet, 64x64 latent, 50 steps, batch 1);
N>1 keeps the per-GPU batch fixed (weak scaling) and all-gathers the decoded ids over NCCL at the
end of every step (the path's only collective).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "latent-diffusion-segmentation_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

# algorithmic FLOPs (SURVEY.md §8d / BASELINE.md §2; 1 MAC = 2 FLOP; conv + linear + QK^T + PV only)
GF_UNET_FWD = 771.4e9
GF_VAE_ENC = 1116.7e9
GF_SEG_DEC = 49.5e9
STEPS_DDIM = 50
FLOP_PER_MASK = STEPS_DDIM * GF_UNET_FWD + GF_VAE_ENC + GF_SEG_DEC

SCHED_KW = dict(prediction_type="epsilon", beta_schedule="scaled_linear", num_train_timesteps=1000,
                beta_start=0.00085, beta_end=0.012, steps_offset=1, clip_sample=False, set_alpha_to_one=False,
                thresholding=False, weight="none", max_snr=5.0)
SEG_KW = dict(in_channels=7, int_channels=256, out_channels=128, block_out_channels=[32, 64, 128, 256],
              latent_channels=4, num_latents=2, num_upscalers=2, upscale_channels=256, norm_num_groups=32,
              scaling_factor=0.18215, parametrization="gaussian")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(burst=d.get("bf16_tflops", 1590.0), sustained=d.get("bf16_tflops_sustained", 1400.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def build_models(device, seed=0):
    import torch
    from ldmseg.models import UNet, GeneralVAESeg, GeneralVAEImage
    from ldmseg.schedulers import DDIMNoiseScheduler
    torch.manual_seed(seed)
    with torch.device(device):
        vae_image = GeneralVAEImage()
        vae_image.set_scaling_factor(0.18215)
        vae_semseg = GeneralVAESeg(**SEG_KW)
        unet = UNet()
        unet.remove_cross_attention()
        # released configuration (tools/scripts/eval.sh:16-17): self-conditioning, 12-ch conv_in.  The image /
        # cond slices are random-init here (the reference zero-inits them before training) so that the RGB
        # branch is exercised numerically -- benchmarking only, stated in `data`.
        unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="random", cond_channels=4,
                            init_mode_cond="random")
    sched = DDIMNoiseScheduler(**SCHED_KW)
    return unet.eval(), vae_image.eval(), vae_semseg.eval(), sched


def run_ours(args):
    import torch
    import torch.distributed as dist
    from ldmseg import _native as nat
    from ldmseg.engine.dist import gather_ids, shard_range
    from ldmseg.engine.sampler import B200Sampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nat.load()
    S, L = args.size, args.size // 8
    steps_ddim = args.ddim_steps
    unet, vae_image, vae_semseg, sched = build_models(dev)
    sampler = B200Sampler(unet, sched, vae_image, vae_semseg, self_condition=True)
    peaks = measured_peaks()
    prof = os.environ.get("LDMSEG_PROFILE") == "1"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def measure(B, sample_kw=None, with_clocks=False, with_roofline=True):
        """One configuration at per-GPU batch B (weak scaling: global batch = B x world).  Inputs follow SURVEY.md
        §8e: RGB and the initial noise are drawn ONCE on the CPU generators for the GLOBAL batch and sliced per
        rank, so the result does not depend on the number of GPUs; decoded ids are all-gathered (engine/dist.py)."""
        sample_kw = dict(sample_kw or {})
        gb = B * world
        lo, hi = shard_range(gb, rank, world)
        host_rgb = torch.rand(gb, 3, S, S, generator=torch.Generator().manual_seed(1234))[lo:hi].contiguous().pin_memory()
        noise = torch.randn(gb, 4, L, L, generator=torch.Generator().manual_seed(42))[lo:hi].contiguous().pin_memory()
        dev_rgb = host_rgb.to(dev)
        if sample_kw.pop("inpaint", False):
            # configs[3]: 50 % random-pixel mask (MaskingGenerator 'random_local', data/util/mask_generator.py:87-91)
            # over ground-truth latents (seeded randn x 0.18215)
            import numpy as np
            mk = torch.from_numpy((np.random.RandomState(7).rand(gb, 1, L, L) < 0.5).astype("float32"))[lo:hi]
            kn = (torch.randn(gb, 4, L, L, generator=torch.Generator().manual_seed(99)) * 0.18215)[lo:hi]
            sample_kw.update(mask=mk.to(dev), known_latents=kn.to(dev))

        def step_device():
            ids, prob = sampler.generate(dev_rgb, steps_ddim, seed=42, noise=noise, **sample_kw)
            return gather_ids(ids, gb), prob

        def step_e2e():
            x = host_rgb.to(dev, non_blocking=True)
            ids, prob = sampler.generate(x, steps_ddim, seed=42, noise=noise, **sample_kw)
            return gather_ids(ids, gb).cpu(), prob.cpu()

        for _ in range(max(args.warmup, 1)):
            step_device()
        clocks = ClockSampler(local) if (with_clocks and rank == 0) else None
        if clocks:
            clocks.start()
        # LDMSEG_PROFILE=1: bracket the timed region with cudaProfilerStart/Stop so that
        # `ncu --profile-from-start off --metrics gpu__time_duration.sum -c N python bench.py ...` lists exactly the
        # launches of the timed region (numbers printed by such a run are not bench values)
        if prof and with_clocks:
            torch.cuda.profiler.start()
        ms = timed(step_device, args.steps)
        if prof and with_clocks:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        clock_info = clocks.stop() if clocks else None
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        masks = gb * args.steps
        value, e2e_value = masks / (ms / 1e3), masks / (ms_e2e / 1e3)
        # launch accounting: kernel nodes executed per mask-batch (graph replays included)
        st = [v for k, v in sampler._state.items() if k[0] == B][-1]
        unet_launch = st["plan"].n_launch + 3
        enc_launch = vae_image._get_engine().plan(B, S).n_launch + 2
        dec_launch = vae_semseg._get_engine().dec_plan(B, L).n_launch + 2
        launches_per_step = enc_launch + st["n"] * unet_launch + dec_launch + 7
        flop_per_mask = st["n"] * GF_UNET_FWD * (L / 64) ** 2 + GF_VAE_ENC * (S / 512) ** 2 + GF_SEG_DEC * (L / 64) ** 2
        if L != 64:   # attention grows quadratically: use the survey's totals for the 128x128 latent
            flop_per_mask = st["n"] * 4555.2e9 + 4879.0e9 + 197.9e9
        per_gpu = value / world
        rec = {
            "value": round(value, 4), "unit": "masks/s", "ms_per_step": round(ms / args.steps, 3),
            "per_gpu_batch": B, "global_batch": gb,
            "e2e": {"value": round(e2e_value, 4), "unit": "masks/s", "ms_per_step": round(ms_e2e / args.steps, 3),
                    "h2d_bytes_per_step": int(gb * 3 * S * S * 4 + gb * 4 * L * L * 4),
                    "d2h_bytes_per_step": int(gb * S * S * (1 + 4)), "bytes_are": "whole job (all ranks)"},
            "gpu_launches": int(launches_per_step * args.steps), "gpu_launches_are": "per rank, whole timed region",
            "roofline_step": {"bound": "tensor", "achieved": round(per_gpu * flop_per_mask / 1e12, 2),
                              "peak": peaks["sustained"], "unit": "TFLOP/s", "peak_kind": f"sustained, {peaks['source']}",
                              "frac": round(per_gpu * flop_per_mask / 1e12 / peaks["sustained"], 4),
                              "flop_per_mask": flop_per_mask},
        }
        if clock_info is not None:
            rec["clocks"] = clock_info
        if with_roofline and rank == 0:
            rec["roofline"] = kernel_roofline(st["plan"], peaks)
            rec["roofline_norm"] = norm_roofline(st["plan"], peaks)
        return rec

    extra = {}
    if args.config == "inpaint":
        extra = dict(inpaint=True)
    elif args.config == "ddpm":
        extra = dict(ddpm=True, ddpm_noise="device")
    main = measure(args.batch, extra, with_clocks=True)
    config3 = None
    if args.config == "ddim" and args.batch != 8 and not args.no_config3 and args.size == 512:
        # BASELINE configs[2]: batch 64 over 8 GPUs = 8 per GPU.  Measured in the same run so that the driver's
        # scaling records carry the batch-sharded configuration next to the batch-1 headline.
        config3 = measure(8, None, with_clocks=False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    B = args.batch
    what = {"ddim": f"{steps_ddim}-step DDIM", "inpaint": f"{steps_ddim}-step DDIM, mask inpainting 50 % sparsity (extension)",
            "ddpm": f"{steps_ddim}-step DDPM (ancestral noise drawn on the device; extension)"}[args.config]
    line = {
        "metric": "panoptic masks/sec (50-step DDIM, 512px, 64x64 latent)",
        "value": main["value"], "unit": "masks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": f"synthetic (torch.rand RGB {S}x{S}, seeded random-init weights; conv_in image/cond slices random instead of zero)",
        "config": {"workload": f"SD-1.5 UNet (12-ch conv_in, cross-attn removed), {S}x{S} RGB, {L}x{L} latent, {what}, batch {B} per GPU",
                   "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": f"batch-sharded x{world}: global RGB / noise draws sliced per rank, all-gather of ids per step",
                   "l2": "bf16 weights streamed per UNet forward (1.63 GB) exceed the 126 MB L2; no explicit flush"},
        "e2e": main["e2e"], "gpu_launches": main["gpu_launches"], "gpu_launches_are": main["gpu_launches_are"],
        "clocks": main.get("clocks"), "roofline": main.get("roofline"), "roofline_norm": main.get("roofline_norm"),
        "roofline_step": main["roofline_step"],
    }
    if config3 is not None:
        config3["config"] = {"workload": "BASELINE configs[2] share: same path, batch 8 per GPU (batch 64 on 8 GPUs)",
                             "per_gpu_batch": 8, "global_batch": 8 * world}
        line["config3"] = config3
    if not args.no_cpu_baseline and world == 1:
        models = _oracle_models()
        line["cpu_baseline"] = cpu_baseline_sample(models)
        if not args.no_library_baseline:
            # free our engines first: the library leg needs its own 3.3 GB of fp32 weights + activations
            line["gpu_library_baseline"] = gpu_library_baseline(models, dev, main, config3)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _graph_ms(fn, reps=10):
    import torch
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def kernel_roofline(plan, peaks):
    """Dominant kernel = the tcgen05 implicit GEMM (84 % of the UNet forward's FLOPs).  Its time is measured
    where it runs in the timed region -- inside the captured CUDA graph of one UNet forward (warm L2, PDL
    overlap): CUDA-event time of the graph minus the time of the same graph with the igemm launches left
    out, on the launching stream.  FLOPs are counted from the launch parameters."""
    import torch
    from ldmseg import _native as nat
    real = nat.igemm
    params = []

    def record(p, simple=False):
        params.append(p)
        real(p, simple)

    try:
        nat.igemm = record
        plan.run()
        torch.cuda.synchronize()
    finally:
        nat.igemm = real
    # ALGORITHMIC FLOPs (the reference's operator, 2 FLOP per multiply-add): an Upsample2D launch
    # counts as the 9-tap convolution over the up-sampled pixels that the reference runs; the folded kernel executes
    # 4/9 of that (four 2x2 phase GEMMs over the input pixels) -- reported beside it as `executed_gflop_per_forward`
    flops, executed = 0.0, 0.0
    for p in params:
        m = p.nb * p.h * p.w
        k = sum(p.seg_taps[i] * p.src_c[p.seg_src[i]] for i in range(p.nseg))
        if p.upsample2:
            flops += 2.0 * (4 * m) * p.n * 9 * p.src_c[p.seg_src[0]]
            executed += 2.0 * (4 * m) * p.n * k
        else:
            flops += 2.0 * m * p.n * k
            executed += 2.0 * m * p.n * k
    is_ig = [t.startswith("igemm:") for t in plan.tags]

    def runner(skip_igemm):
        def f():
            old = nat.set_pdl(plan.pdl)
            try:
                for op, ig in zip(plan.ops, is_ig):
                    if not (skip_igemm and ig):
                        op()
            finally:
                nat.set_pdl(old)
        return f

    def only_igemm():
        old = nat.set_pdl(plan.pdl)
        try:
            for op, ig in zip(plan.ops, is_ig):
                if ig:
                    op()
        finally:
            nat.set_pdl(old)

    full_ms = _graph_ms(runner(False))
    rest_ms = _graph_ms(runner(True))
    ig_ms = max(full_ms - rest_ms, 1e-6)
    # cross-check: the 148 igemm launches alone, back to back in plan order, in one graph (operands = what the full
    # run left in the buffers); no subtraction, but no neighbours to overlap with either
    only_ms = _graph_ms(only_igemm)
    n = sum(is_ig)
    achieved = flops / (ig_ms / 1e3) / 1e12
    # dram bytes per launch of this kernel from the committed ncu --set full capture (tools/make_traffic_json.py)
    traffic, traffic_src = None, None
    # (one file per captured batch: the batch-8 capture holds the 320-wide pair tiles)
    tpath = os.path.join(ROOT, "profiles", "igemm_dram_traffic.json")
    tpath8 = os.path.join(ROOT, "profiles", "igemm_dram_traffic_b8.json")
    if getattr(plan, "nb", 1) >= 8 and os.path.exists(tpath8):
        tpath = tpath8
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("traffic"), f"{tj.get('source')}: {tj.get('unit')}, {tj.get('launches_captured')} launches"
    return {"bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit GEMM: conv3x3/conv1x1/linear)",
            "achieved": round(achieved, 2), "peak": peaks["burst"], "unit": "TFLOP/s",
            "peak_kind": f"burst, {peaks['source']}", "frac": round(achieved / peaks["burst"], 4),
            "traffic": traffic, "traffic_source": traffic_src, "launches": n, "algorithmic_gflop_per_launch": round(flops / 1e9 / max(n, 1), 2),
            "algorithmic_gflop_per_forward": round(flops / 1e9, 1),
            "executed_gflop_per_forward": round(executed / 1e9, 1), "avg_launch_us": round(ig_ms * 1e3 / max(n, 1), 2),
            "share_of_unet_forward": round(ig_ms / full_ms, 3), "unet_forward_ms_graph": round(full_ms, 3),
            "frac_of_sustained_peak": round(achieved / peaks["sustained"], 4),
            "igemm_only_graph": {"ms": round(only_ms, 3), "achieved": round(flops / (only_ms / 1e3) / 1e12, 2),
                                 "frac": round(flops / (only_ms / 1e3) / 1e12 / peaks["burst"], 4)},
            "method": "in-graph: CUDA-event time of the UNet-forward graph minus the same graph without its "
                      "igemm launches (batch as benchmarked)"}


def norm_roofline(plan, peaks):
    """Second roofline entry: the GroupNorm(+SiLU) apply pass (61 launches per forward), HBM-bound.  Algorithmic
    bytes: read + write the activation once, 2 x rows x C x 2 B (bf16) per launch; time = the
    forward graph minus the same graph without its GroupNorm launches."""
    is_gn = [t.startswith("gn:") or t.startswith("gn3:") for t in plan.tags]
    from ldmseg import _native as nat
    nbytes = 0.0
    for t in plan.tags:
        if t.startswith("gn:") or t.startswith("gn3:"):
            f = t.split(":")
            nbytes += 2.0 * int(f[1]) * int(f[-1][1:]) * 2      # rows x channels x (read + write) x bf16

    def runner(skip):
        def f():
            old = nat.set_pdl(plan.pdl)
            try:
                for op, g in zip(plan.ops, is_gn):
                    if not (skip and g):
                        op()
            finally:
                nat.set_pdl(old)
        return f

    full_ms = _graph_ms(runner(False))
    rest_ms = _graph_ms(runner(True))
    gn_ms = max(full_ms - rest_ms, 1e-6)
    n = sum(is_gn)
    ach = nbytes / (gn_ms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": "gn_apply_cs_kernel (GroupNorm + SiLU apply, statistics from the producer epilogue)",
            "achieved": round(ach, 1), "peak": peaks["hbm"], "unit": "GB/s", "frac": round(ach / peaks["hbm"], 4),
            "traffic": None, "launches": n, "algorithmic_mb_per_launch": round(nbytes / 1e6 / max(n, 1), 3),
            "avg_launch_us": round(gn_ms * 1e3 / max(n, 1), 2), "share_of_unet_forward": round(gn_ms / full_ms, 3),
            "method": "in-graph: forward graph minus the same graph without its GroupNorm launches"}


def gpu_library_baseline(models, dev, main, config3):
    """The bar the survey set for the B200 box (SURVEY.md §2.1 / BASELINE.md §4): the SAME architecture (the oracle's
    restated diffusers modules -- the reference's model code without our kernels) on stock PyTorch library kernels
    (cuDNN / cuBLAS / SDPA) on this GPU: fp32 with TF32 off (the reference's precision, tools/main_ldm.py:168,
    base.yaml:95) and bf16 channels-last, eager and with the UNet forward captured in a CUDA graph; batch 1 and 8.
    Reports ms per UNet forward and masks/s for the whole chain (encode + 50 x (UNet + scheduler.step) + decode)."""
    import torch
    unet, vae_image, seg, sched, orc = models
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {"what": "restated diffusers model on stock PyTorch kernels (cuDNN/cuBLAS/SDPA), same GPU, same run",
           "torch": torch.__version__, "variants": {}}
    g = torch.Generator().manual_seed(1234)
    ts = None

    def ev_ms(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    try:
        for dtype, name, fmt in ((torch.float32, "fp32_no_tf32", torch.contiguous_format),
                                 (torch.bfloat16, "bf16_channels_last", torch.channels_last)):
            u = unet.to(device=dev, dtype=dtype).to(memory_format=fmt)
            vi = vae_image.to(device=dev, dtype=dtype).to(memory_format=fmt)
            sg = seg.to(device=dev, dtype=torch.float32)
            sched.set_timesteps_inference(STEPS_DDIM)
            ts = sched.timesteps
            for B in (1, 8):
                x = torch.randn(B, 12, 64, 64, generator=g).to(device=dev, dtype=dtype).contiguous(memory_format=fmt)
                t_dev = torch.tensor(999, device=dev)
                with torch.no_grad():
                    eager = ev_ms(lambda: u(x, t_dev).sample, 5)
                    static_out = None
                    gr = torch.cuda.CUDAGraph()
                    try:
                        side = torch.cuda.Stream()
                        side.wait_stream(torch.cuda.current_stream())
                        with torch.cuda.stream(side):
                            for _ in range(2):
                                u(x, t_dev)
                        torch.cuda.current_stream().wait_stream(side)
                        with torch.cuda.graph(gr):
                            static_out = u(x, t_dev).sample
                        graph = ev_ms(gr.replay, 10)
                    except Exception as e:  # noqa: BLE001 - a library limitation is a result, not a failure of the bench
                        graph, gr = None, None
                        out.setdefault("notes", []).append(f"{name} b{B}: graph capture failed: {type(e).__name__}")
                    # whole chain, UNet graphed when possible
                    rgb = torch.rand(B, 3, 512, 512, generator=g).to(dev)
                    lat0 = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(42)).to(dev)

                    def chain():
                        rl = vi.encode((2 * rgb - 1).to(dtype)).latent_dist.mode().float() * 0.18215
                        lat, cond = lat0.clone(), torch.zeros_like(lat0)
                        for i, t in enumerate(ts):
                            inp = torch.cat([lat, rl, cond], 1).to(dtype)
                            if gr is not None:
                                x.copy_(inp)
                                t_dev.fill_(int(t))
                                gr.replay()
                                eps = static_out.float()
                            else:
                                eps = u(inp, t.to(dev)).sample.float()
                            o = sched.step(eps, t, lat)
                            cond = o.pred_original_sample
                            lat = o.pred_original_sample if i == len(ts) - 1 else o.prev_sample
                        return orc.decode_latents(lat, sg).argmax(1)

                    chain_ms = ev_ms(chain, 1 if dtype == torch.float32 else 2, warm=1)
                ours = main if B == main["per_gpu_batch"] else (config3 if config3 and B == config3["per_gpu_batch"] else None)
                rec = {"unet_forward_ms_eager": round(eager, 3),
                       "unet_forward_ms_graph": None if graph is None else round(graph, 3),
                       "masks_per_s": round(B / (chain_ms / 1e3), 3), "ms_per_mask_batch": round(chain_ms, 2)}
                if ours is not None:
                    rec["ours_masks_per_s"] = ours["value"]
                    rec["ours_over_library"] = round(ours["value"] / max(rec["masks_per_s"], 1e-9), 3)
                    if ours.get("roofline"):
                        rec["ours_unet_forward_ms_graph"] = ours["roofline"]["unet_forward_ms_graph"]
                out["variants"][f"{name}_b{B}"] = rec
                del gr, static_out
                torch.cuda.empty_cache()
    finally:
        unet.to("cpu", torch.float32)
        vae_image.to("cpu", torch.float32)
        seg.to("cpu")
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def _oracle_models(seed=0):
    import torch
    from oracle import diffusers_restated as dr
    from oracle import ldmseg_restated as orc
    torch.manual_seed(seed)
    unet = orc.build_ldmseg_unet(seed=seed, cond_channels=4, image_init="zero")
    vae_image = dr.AutoencoderKL().eval()
    seg = orc.GeneralVAESeg(**{k: v for k, v in SEG_KW.items() if k != "parametrization"}).eval()
    sched = orc.DDIMNoiseScheduler(**SCHED_KW)
    return unet, vae_image, seg, sched, orc


def cpu_baseline_sample(models=None):
    """Bounded sample of configs[1] on the host cores with the oracle port (fp32 PyTorch CPU): the VAE encode,
    2 of the 50 UNet+scheduler steps, and the decode; masks/s extrapolated to 50 steps."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae_image, seg, sched, orc = models or _oracle_models()
    g = torch.Generator().manual_seed(1234)
    rgb = torch.rand(1, 3, 512, 512, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        rgb_lat = orc.encode_inputs(rgb, vae_image, 0.18215)
        t_enc = time.perf_counter() - t0
        sched.set_timesteps_inference(STEPS_DDIM)
        lat = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(42))
        cond = torch.zeros_like(rgb_lat)
        ts = sched.timesteps
        unet(torch.cat([lat, rgb_lat, cond], 1), ts[0])  # warm-up
        t0 = time.perf_counter()
        for i in range(2):
            eps = unet(torch.cat([lat, rgb_lat, cond], 1), ts[i]).sample
            out = sched.step(eps, ts[i], lat)
            cond, lat = out.pred_original_sample, out.prev_sample
        t_step = (time.perf_counter() - t0) / 2
        t0 = time.perf_counter()
        logits = orc.decode_latents(lat, seg)
        logits.argmax(1)
        t_dec = time.perf_counter() - t0
    per_mask = t_enc + STEPS_DDIM * t_step + t_dec
    return {"value": round(1.0 / per_mask, 5), "unit": "masks/s", "cores": cores, "kind": "port",
            "sample": f"oracle fp32 on CPU: 1 VAE encode ({t_enc:.2f}s) + 2 of 50 UNet+DDIM steps ({t_step:.2f}s each) "
                      f"+ 1 decode ({t_dec:.2f}s), batch 1, extrapolated to 50 steps",
            "seconds_per_mask_extrapolated": round(per_mask, 2)}


def run_reference(args):
    """The reference's own CPU implementation of the path = the oracle port (the reference is pure Python on
    top of diffusers, which is absent).  Each step = ONE UNet forward + scheduler step at batch 1
    (1/50 of a mask); masks/s is extrapolated with the encode / decode measured once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae_image, seg, sched, orc = _oracle_models()
    g = torch.Generator().manual_seed(1234)
    rgb = torch.rand(1, 3, 512, 512, generator=g)
    with torch.no_grad():
        t0 = time.perf_counter()
        rgb_lat = orc.encode_inputs(rgb, vae_image, 0.18215)
        t_enc = time.perf_counter() - t0
        sched.set_timesteps_inference(STEPS_DDIM)
        lat = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(42))
        cond = torch.zeros_like(rgb_lat)
        ts = sched.timesteps

        def one(i):
            nonlocal lat, cond
            eps = unet(torch.cat([lat, rgb_lat, cond], 1), ts[i % len(ts)]).sample
            out = sched.step(eps, ts[i % len(ts)], lat)
            cond, lat = out.pred_original_sample, out.prev_sample

        for i in range(args.warmup):
            one(i)
        t0 = time.perf_counter()
        for i in range(args.steps):
            one(args.warmup + i)
        t_step = (time.perf_counter() - t0) / max(args.steps, 1)
        t0 = time.perf_counter()
        orc.decode_latents(torch.nan_to_num(lat), seg).argmax(1)
        t_dec = time.perf_counter() - t0
    per_mask = t_enc + STEPS_DDIM * t_step + t_dec
    value = 1.0 / per_mask
    sample = (f"oracle fp32 on {cores} CPU threads, batch 1: each timed step = 1 UNet forward + DDIM step "
              f"({t_step:.2f}s); encode {t_enc:.2f}s and decode {t_dec:.2f}s measured once; masks/s extrapolated to 50 steps")
    line = {
        "impl": "reference", "metric": "panoptic masks/sec (50-step DDIM, 512px, 64x64 latent)",
        "value": round(value, 5), "unit": "masks/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t_step * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (torch.rand RGB 512x512, seeded random-init weights)",
        "config": {"workload": "SD-1.5 UNet (12-ch conv_in, cross-attn removed), 512x512 RGB, 64x64 latent, 50-step DDIM, batch 1 (CPU)"},
        "cpu_baseline": {"value": round(value, 5), "unit": "masks/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 5), "unit": "masks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="per-GPU batch (configs[1] = 1; configs[2] shards 8 per GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the stock-PyTorch GPU leg")
    ap.add_argument("--no-config3", action="store_true", help="skip the batch-8-per-GPU record (BASELINE configs[2])")
    ap.add_argument("--config", default="ddim", choices=["ddim", "inpaint", "ddpm"],
                    help="ddim = configs[1]/[2]; inpaint = configs[3] (mask inpainting); ddpm = configs[4] (use --size 1024 "
                         "--ddim-steps 100 --batch 2)")
    ap.add_argument("--size", type=int, default=512, help="RGB size (1024 for configs[4])")
    ap.add_argument("--ddim-steps", type=int, default=STEPS_DDIM)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
