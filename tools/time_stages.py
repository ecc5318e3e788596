"""Where a mask-batch's time goes outside the UNet loop: CUDA-event timing of the stages of B200Sampler.generate."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "latent-diffusion-segmentation_b200"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    from bench import build_models
    from ldmseg.engine.sampler import B200Sampler
    from ldmseg.schedulers import DDIMNoiseScheduler
    from oracle.make_golden import SCHED_KW
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    unet, vae_image, vae_semseg, _ = build_models(dev)
    sch = DDIMNoiseScheduler(**SCHED_KW)
    sampler = B200Sampler(unet, sch, vae_image, vae_semseg)
    B = args.batch
    rgb = torch.rand(B, 3, 512, 512, generator=torch.Generator().manual_seed(1234)).to(dev)
    noise = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(42)).pin_memory()
    for _ in range(2):
        sampler.generate(rgb, 50, seed=42, noise=noise)
    torch.cuda.synchronize()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    for rep in range(2):
        e0 = ev()
        lat = sampler.encode_rgb(rgb)
        e1 = ev()
        z = sampler.sample(lat, 50, seed=42, noise=noise)
        e2 = ev()
        ids, prob = vae_semseg._get_engine().decode_ids(z, scale=1.0 / vae_semseg.scaling_factor)
        e3 = ev()
        torch.cuda.synchronize()
        print(f"batch {B}: encode {e0.elapsed_time(e1):.2f} ms, sample {e1.elapsed_time(e2):.2f} ms, "
              f"decode_ids {e2.elapsed_time(e3):.2f} ms, total {e0.elapsed_time(e3):.2f} ms")


if __name__ == "__main__":
    main()
