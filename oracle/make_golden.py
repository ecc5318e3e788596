"""ORACLE -- test infrastructure only.  Generates tests/golden/*.npz from THE REFERENCE ITSELF.

Runs only in the build container, where /root/reference is mounted.  The reference's own files are
imported unmodified under `sys.modules` stubs for the third-party packages that are absent here:

  detectron2.utils.visualizer, easydict   -> empty stubs (only touched by logging/visualisation)
  diffusers, diffusers.models.unet_2d_blocks, diffusers.training_utils
                                          -> oracle.diffusers_restated (CPU restatement, App. A)

so that
  * ldmseg/schedulers/ddim_scheduler.py and ldmseg/models/vae.py (GeneralVAESeg, LayerNorm2d,
    DiagonalGaussianDistribution) run as the reference wrote them -> true golden vectors;
  * ldmseg/models/unet.py (UNet.forward / modify_encoder / remove_cross_attention) runs unmodified on
    top of the restated diffusers base class -> golden vectors for the LDMSeg glue.

Usage:  python -m oracle.make_golden            (writes tests/golden/, a few MB)
The fixtures are committed; nothing at test/bench time reads /root/reference.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

# conv_in must keep 320 output channels: the reference asserts it (ldmseg/models/unet.py:216)
TINY = dict(block_out_channels=(320, 64, 128, 128), attention_head_dim=4, cross_attention_dim=64)


def import_reference():
    sys.path.insert(0, ROOT)
    from oracle import diffusers_restated as dr

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    dummy = type("Dummy", (), {})
    stub("detectron2")
    stub("detectron2.utils")
    stub("detectron2.utils.visualizer", Visualizer=dummy, _PanopticPrediction=dummy, ColorMode=dummy,
         _OFF_WHITE=(1, 1, 1), _create_text_labels=lambda *a, **k: None)
    stub("easydict", EasyDict=dict)
    stub("diffusers", UNet2DConditionModel=dr.UNet2DConditionModel, AutoencoderKL=dr.AutoencoderKL)
    stub("diffusers.models")
    stub("diffusers.models.unet_2d_blocks", UNetMidBlock2D=dr.UNetMidBlock2D)
    stub("diffusers.training_utils", EMAModel=dr.EMAModel)
    sys.path.insert(0, REF)
    import ldmseg  # noqa: F401  (the reference package)
    from ldmseg.models.unet import UNet
    from ldmseg.models.vae import GeneralVAESeg, GeneralVAEImage
    from ldmseg.schedulers.ddim_scheduler import DDIMNoiseScheduler
    assert ldmseg.__file__.startswith(REF), ldmseg.__file__
    return UNet, GeneralVAESeg, GeneralVAEImage, DDIMNoiseScheduler


SCHED_KW = dict(prediction_type="epsilon", beta_schedule="scaled_linear", num_train_timesteps=1000,
                beta_start=0.00085, beta_end=0.012, steps_offset=1, clip_sample=False,
                set_alpha_to_one=False, thresholding=False, dynamic_thresholding_ratio=0.995,
                clip_sample_range=1.0, sample_max_value=1.0, weight="none", max_snr=5.0)
VAE_KW = dict(in_channels=7, int_channels=256, out_channels=128, block_out_channels=[32, 64, 128, 256],
              latent_channels=4, num_latents=2, num_upscalers=2, upscale_channels=256, norm_num_groups=32,
              scaling_factor=0.2, parametrization="gaussian", act_fn="none", clamp_output=False,
              freeze_codebook=False, num_mid_blocks=0, fuse_rgb=False, resize_input=False,
              skip_encoder=False)


def golden_scheduler(DDIM):
    out = {}
    g = torch.Generator().manual_seed(0)
    eps = torch.randn(2, 4, 16, 16, generator=g)
    x = torch.randn(2, 4, 16, 16, generator=g)
    out["eps"], out["x"] = eps.numpy(), x.numpy()
    for ptype in ("epsilon", "sample", "v_prediction"):
        kw = dict(SCHED_KW)
        kw["prediction_type"] = ptype
        s = DDIM(**kw)
        if ptype == "epsilon":
            out["alphas_cumprod"] = s.alphas_cumprod.numpy()
            out["final_alpha_cumprod"] = s.final_alpha_cumprod.numpy()
        for n in (10, 50, 100):
            s.set_timesteps_inference(n)
            out[f"timesteps_{n}"] = s.timesteps.numpy()
            for which, idx in (("first", 0), ("mid", n // 2), ("last", n - 1)):
                r = s.step(eps, s.timesteps[idx], x)
                out[f"step_{ptype}_{n}_{which}_prev"] = r.prev_sample.numpy()
                out[f"step_{ptype}_{n}_{which}_x0"] = r.pred_original_sample.numpy()
    # clip_sample + use_clipped_model_output branch
    kw = dict(SCHED_KW)
    kw["clip_sample"] = True
    s = DDIM(**kw)
    s.set_timesteps_inference(50)
    r = s.step(eps, s.timesteps[3], x, use_clipped_model_output=True)
    out["step_clip_prev"], out["step_clip_x0"] = r.prev_sample.numpy(), r.pred_original_sample.numpy()
    # add_noise / remove_noise
    s = DDIM(**SCHED_KW)
    t = torch.tensor([999, 19])
    noisy = s.add_noise(x, eps.clone(), t)
    out["add_noise"] = noisy.numpy()
    out["remove_noise"] = s.remove_noise(noisy, eps, t).numpy()
    # other beta schedules
    for sched in ("linear", "squaredcos_cap_v2", "sigmoid"):
        kw = dict(SCHED_KW)
        kw["beta_schedule"] = sched
        out[f"alphas_cumprod_{sched}"] = DDIM(**kw).alphas_cumprod.numpy()
    np.savez_compressed(os.path.join(OUT, "scheduler.npz"), **out)
    print("scheduler.npz:", len(out), "arrays")


def golden_segvae(VAESeg):
    torch.manual_seed(0)
    vae = VAESeg(**VAE_KW).eval()
    print("GeneralVAESeg params", sum(p.numel() for p in vae.parameters()))
    print("encoder keys", sorted({k.split('.')[1] for k in vae.state_dict() if k.startswith('encoder')}, key=int))
    print("decoder keys", sorted({k.split('.')[1] for k in vae.state_dict() if k.startswith('decoder')}, key=int))
    g = torch.Generator().manual_seed(1)
    z = torch.randn(1, 4, 16, 16, generator=g)
    bits = (torch.rand(1, 7, 64, 64, generator=g) > 0.5).float()
    with torch.no_grad():
        dec = vae.decode(z / 0.2)
        dec_nointerp = vae.decode(z / 0.2, interpolate=False)
        post = vae.encode(bits).latent_dist
    out = dict(z=z.numpy(), bits=bits.numpy(), decode_head=dec[:, :8].numpy(),
               decode_chan_mean=dec.mean(dim=(0, 2, 3)).numpy(), decode_chan_std=dec.std(dim=(0, 2, 3)).numpy(),
               decode_argmax=dec.argmax(1).numpy().astype(np.uint8),
               decode_nointerp_head=dec_nointerp[:, :8].numpy(), enc_mean=post.mode().numpy(),
               enc_logvar=post.logvar.numpy(),
               state_keys=np.array(sorted(vae.state_dict().keys())))
    np.savez_compressed(os.path.join(OUT, "segvae.npz"), **out)
    print("segvae.npz written; decode", tuple(dec.shape))


def golden_unet_glue(UNet):
    """The reference's UNet subclass (forward/modify_encoder/remove_cross_attention) on the restated
    base, tiny width (same topology) so the fixture and its test are fast."""
    torch.manual_seed(0)
    unet = UNet.from_pretrained(None, subfolder="unet", **TINY)
    unet.remove_cross_attention()
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4,
                        init_mode_cond="zero")
    unet.eval()
    # the default image/cond init is zero; randomise those slices (seeded) so the RGB/self-cond inputs matter
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        unet.conv_in.weight[:, 4:].copy_(torch.randn(unet.conv_in.weight[:, 4:].shape, generator=g) * 0.05)
    x = torch.randn(2, 12, 16, 16, generator=g)
    out = {"x": x.numpy(), "conv_in_tail": unet.conv_in.weight[:, 4:].detach().numpy()}
    with torch.no_grad():
        for t in (999, 19):
            y = unet(x, torch.tensor(t), encoder_hidden_states=None).sample
            out[f"y_t{t}"] = y.numpy()
        yt = unet(x, torch.tensor(500), encoder_hidden_states=None, return_dict=False)
        assert isinstance(yt, tuple)
    out["n_params"] = np.array(sum(p.numel() for p in unet.parameters()))
    out["state_keys"] = np.array(sorted(unet.state_dict().keys()))
    np.savez_compressed(os.path.join(OUT, "unet_glue_tiny.npz"), **out)
    print("unet_glue_tiny.npz written; params", int(out["n_params"]))


def golden_sample_loop(UNet, VAESeg, DDIM):
    """10-step sample() semantics (BASELINE config #1 in miniature): the reference's scheduler + UNet
    glue driven by a line-for-line transcription of TrainerDiffusion.sample's loop body
    (trainers_ldm_cond.py:1088-1159; the trainer class itself needs COCO/detectron2 to construct)."""
    torch.manual_seed(0)
    unet = UNet.from_pretrained(None, subfolder="unet", **TINY)
    unet.remove_cross_attention()
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4,
                        init_mode_cond="zero")
    unet.eval()
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        unet.conv_in.weight[:, 4:].copy_(torch.randn(unet.conv_in.weight[:, 4:].shape, generator=g) * 0.05)
    rgb_latents = torch.randn(1, 4, 16, 16, generator=g) * 0.18215 * 4
    sched = DDIM(**SCHED_KW)
    sched.set_timesteps_inference(10)
    rng = torch.Generator().manual_seed(42)
    latents = torch.randn((1, 4, 16, 16), generator=rng)
    latents = latents * sched.init_noise_sigma
    condition = torch.zeros_like(rgb_latents)
    with torch.no_grad():
        for i, t in enumerate(sched.timesteps):
            inputs = torch.cat([latents, rgb_latents, condition], dim=1)
            noise_pred = unet(inputs, t, encoder_hidden_states=None).sample
            condition = sched.step(noise_pred, t, latents).pred_original_sample
            if i == len(sched.timesteps) - 1:
                latents = sched.step(noise_pred, t, latents).pred_original_sample
            else:
                latents = sched.step(noise_pred, t, latents).prev_sample
    np.savez_compressed(os.path.join(OUT, "sample_loop_tiny.npz"), rgb_latents=rgb_latents.numpy(),
                        final_latents=latents.numpy())
    print("sample_loop_tiny.npz written")


def main():
    os.makedirs(OUT, exist_ok=True)
    UNet, VAESeg, VAEImage, DDIM = import_reference()
    golden_scheduler(DDIM)
    golden_segvae(VAESeg)
    golden_unet_glue(UNet)
    golden_sample_loop(UNet, VAESeg, DDIM)


if __name__ == "__main__":
    main()
