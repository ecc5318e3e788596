"""Times the DROP-IN path: the reference's own loop shape -- `unet(inputs, t, encoder_hidden_states=None).sample` +
`scheduler.step(...)` twice per iteration, as TrainerDiffusion.sample (trainers_ldm_cond.py:1127-1159) drives them from
Python -- against the fused B200Sampler, both at batch 1, 50 steps, 64x64 latent.  `unet(...)` replays one captured
graph of the forward per call (LDMSEG_CUDA_GRAPH=0: ~230 ctypes launches per call instead)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))

import torch  # noqa: E402


def main():
    from bench import build_models
    from ldmseg.engine.sampler import B200Sampler
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    unet, vae_image, vae_semseg, sched = build_models(dev)
    rgb = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(1)).to(dev) * 0.7

    def api_loop():
        sched.set_timesteps_inference(50)
        sched.move_timesteps_to(dev)
        lat = torch.randn((1, 4, 64, 64), generator=torch.Generator().manual_seed(42)).to(dev)
        cond = torch.zeros_like(rgb)
        for i, t in enumerate(sched.timesteps):
            eps = unet(torch.cat([lat, rgb, cond], 1), t, encoder_hidden_states=None).sample
            cond = sched.step(eps, t, lat).pred_original_sample
            o = sched.step(eps, t, lat)
            lat = o.pred_original_sample if i == len(sched.timesteps) - 1 else o.prev_sample
        return lat

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    out = {"what": "50-step loop at batch 1, 64x64 latent, wall-clock ms per loop (host + device)"}
    out["api_loop_graphed_unet_ms"] = round(timed(api_loop), 2)
    unet._get_engine().use_graph = False
    out["api_loop_eager_launches_ms"] = round(timed(api_loop), 2)
    unet._get_engine().use_graph = True
    s = B200Sampler(unet, sched, self_condition=True)
    out["fused_sampler_ms"] = round(timed(lambda: s.sample(rgb, 50, seed=42)), 2)
    a, b = api_loop(), s.sample(rgb, 50, seed=42)
    out["api_vs_fused_rel_l2"] = float(((a - b).norm() / b.norm()).item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
