"""CPU tests: the oracle restatement (oracle/) against golden vectors produced by the reference's own
files (oracle/make_golden.py imports /root/reference in the build container)."""
import numpy as np
import pytest
import torch

from oracle import diffusers_restated as dr
from oracle import ldmseg_restated as orc
from oracle.make_golden import SCHED_KW, VAE_KW, TINY


@pytest.fixture(scope="module")
def sched_golden(golden_dir):
    return np.load(f"{golden_dir}/scheduler.npz")


def test_param_counts_match_published():
    with torch.device("meta"):
        unet = dr.UNet2DConditionModel()
        vae = dr.AutoencoderKL()
    assert sum(p.numel() for p in unet.parameters()) == 859_520_964      # stock SD-v1 UNet
    assert sum(p.numel() for p in vae.encoder.parameters()) == 34_163_592  # AutoencoderKL encoder
    with torch.device("meta"):
        u = orc.UNet()
        u.remove_cross_attention()
        u.modify_encoder(in_channels=8, cond_channels=4)
    assert sum(p.numel() for p in u.parameters()) == 815_556_484         # LDMSeg surgery, 12-ch conv_in


def test_scheduler_tables_and_grids(sched_golden):
    s = orc.DDIMNoiseScheduler(**SCHED_KW)
    np.testing.assert_array_equal(s.alphas_cumprod.numpy(), sched_golden["alphas_cumprod"])
    np.testing.assert_array_equal(s.final_alpha_cumprod.numpy(), sched_golden["final_alpha_cumprod"])
    for n in (10, 50, 100):
        s.set_timesteps_inference(n)
        np.testing.assert_array_equal(s.timesteps.numpy(), sched_golden[f"timesteps_{n}"])
    assert s.timesteps[0] == 999
    for name in ("linear", "squaredcos_cap_v2", "sigmoid"):
        kw = dict(SCHED_KW, beta_schedule=name)
        np.testing.assert_array_equal(orc.DDIMNoiseScheduler(**kw).alphas_cumprod.numpy(),
                                      sched_golden[f"alphas_cumprod_{name}"])


@pytest.mark.parametrize("ptype", ["epsilon", "sample", "v_prediction"])
def test_scheduler_step_bit_exact(sched_golden, ptype):
    eps, x = torch.from_numpy(sched_golden["eps"]), torch.from_numpy(sched_golden["x"])
    s = orc.DDIMNoiseScheduler(**dict(SCHED_KW, prediction_type=ptype))
    for n in (10, 50, 100):
        s.set_timesteps_inference(n)
        for which, idx in (("first", 0), ("mid", n // 2), ("last", n - 1)):
            r = s.step(eps, s.timesteps[idx], x)
            np.testing.assert_array_equal(r.prev_sample.numpy(), sched_golden[f"step_{ptype}_{n}_{which}_prev"])
            np.testing.assert_array_equal(r.pred_original_sample.numpy(),
                                          sched_golden[f"step_{ptype}_{n}_{which}_x0"])


def test_scheduler_clip_and_noise(sched_golden):
    eps, x = torch.from_numpy(sched_golden["eps"]), torch.from_numpy(sched_golden["x"])
    s = orc.DDIMNoiseScheduler(**dict(SCHED_KW, clip_sample=True))
    s.set_timesteps_inference(50)
    r = s.step(eps, s.timesteps[3], x, use_clipped_model_output=True)
    np.testing.assert_array_equal(r.prev_sample.numpy(), sched_golden["step_clip_prev"])
    np.testing.assert_array_equal(r.pred_original_sample.numpy(), sched_golden["step_clip_x0"])
    s = orc.DDIMNoiseScheduler(**SCHED_KW)
    t = torch.tensor([999, 19])
    noisy = s.add_noise(x, eps, t)
    np.testing.assert_array_equal(noisy.numpy(), sched_golden["add_noise"])
    np.testing.assert_array_equal(s.remove_noise(noisy, eps, t).numpy(), sched_golden["remove_noise"])


def _seg_kwargs():
    return {k: v for k, v in VAE_KW.items()
            if k in ("in_channels", "int_channels", "out_channels", "block_out_channels", "latent_channels",
                     "norm_num_groups", "scaling_factor", "num_latents", "num_upscalers", "upscale_channels")}


def test_segvae_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/segvae.npz")
    torch.manual_seed(0)
    vae = orc.GeneralVAESeg(**_seg_kwargs()).eval()
    assert sorted(vae.state_dict().keys()) == list(g["state_keys"])
    assert sum(p.numel() for p in vae.parameters()) == 2_023_208
    z, bits = torch.from_numpy(g["z"]), torch.from_numpy(g["bits"])
    with torch.no_grad():
        dec = vae.decode(z / 0.2)
        dec_ni = vae.decode(z / 0.2, interpolate=False)
        post = vae.encode(bits).latent_dist
    np.testing.assert_allclose(dec[:, :8].numpy(), g["decode_head"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(dec.mean(dim=(0, 2, 3)).numpy(), g["decode_chan_mean"], atol=1e-5)
    np.testing.assert_array_equal(dec.argmax(1).numpy().astype(np.uint8), g["decode_argmax"])
    np.testing.assert_allclose(dec_ni[:, :8].numpy(), g["decode_nointerp_head"], atol=1e-5)
    np.testing.assert_allclose(post.mode().numpy(), g["enc_mean"], atol=1e-5)
    np.testing.assert_allclose(post.logvar.numpy(), g["enc_logvar"], atol=1e-5)


def _tiny_unet():
    torch.manual_seed(0)
    unet = orc.UNet(**TINY)
    unet.remove_cross_attention()
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4,
                        init_mode_cond="zero")
    return unet.eval()


def test_unet_glue_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/unet_glue_tiny.npz")
    unet = _tiny_unet()
    assert sorted(unet.state_dict().keys()) == list(g["state_keys"])
    assert sum(p.numel() for p in unet.parameters()) == int(g["n_params"])
    with torch.no_grad():
        unet.conv_in.weight[:, 4:].copy_(torch.from_numpy(g["conv_in_tail"]))
        x = torch.from_numpy(g["x"])
        for t in (999, 19):
            y = unet(x, torch.tensor(t)).sample
            np.testing.assert_allclose(y.numpy(), g[f"y_t{t}"], rtol=0, atol=2e-5)
        assert isinstance(unet(x, torch.tensor(500), return_dict=False), tuple)


def test_sample_loop_matches_reference(golden_dir):
    g = np.load(f"{golden_dir}/sample_loop_tiny.npz")
    gu = np.load(f"{golden_dir}/unet_glue_tiny.npz")
    unet = _tiny_unet()
    with torch.no_grad():
        unet.conv_in.weight[:, 4:].copy_(torch.from_numpy(gu["conv_in_tail"]))
    sched = orc.DDIMNoiseScheduler(**SCHED_KW)
    out = orc.sample(unet, sched, torch.from_numpy(g["rgb_latents"]), num_inference_steps=10, seed=42)
    np.testing.assert_allclose(out.numpy(), g["final_latents"], rtol=0, atol=5e-4)


def test_ddpm_and_inpaint_extensions_are_consistent():
    """The two labelled extensions reduce to plain DDIM when switched off."""
    unet = _tiny_unet()
    sched = orc.DDIMNoiseScheduler(**SCHED_KW)
    rgb = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    base = orc.sample(unet, sched, rgb, 10, seed=1)
    zero_mask = torch.zeros(1, 1, 16, 16)
    same = orc.sample(unet, sched, rgb, 10, seed=1, mask=zero_mask, known_latents=torch.zeros(1, 4, 16, 16))
    torch.testing.assert_close(base, same)
    full = orc.sample(unet, sched, rgb, 10, seed=1, mask=torch.ones(1, 1, 16, 16), known_latents=rgb)
    torch.testing.assert_close(full, rgb)  # fully known region is returned unchanged
    noisy = orc.sample(unet, sched, rgb, 10, seed=1, ddpm=True)
    assert torch.isfinite(noisy).all() and not torch.allclose(noisy, base)


def test_panoptic_postprocess_rules():
    """The oracle transcription of compute_pq's per-image tail (trainers_ldm_cond.py:1261-1313) on structured logits:
    area rule, ignore label, overlap rule, and agreement of the split (float stage / integer stage) formulation that
    the CUDA kernels are checked against."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    win = torch.randint(0, 6, (1, 8, 8), generator=g)
    win[:, 0, 0] = 10                     # one small cell -> below count_th
    win[:, 7, 6:8] = 11                   # wins two cells ...
    base = torch.full((1, 128, 8, 8), -8.0)
    base.scatter_(1, win[:, None], 8.0)
    halo = (win != 11) & (torch.rand(1, 8, 8, generator=g) < 0.6)
    base[:, 11][halo] = 1.5               # ... but is weakly positive over many more -> fails overlap_th
    logits = F.interpolate(base, size=(128, 128), mode="nearest")
    (seg, ids), = orc.panoptic_postprocess(logits, [(96, 120)], 0.5, 300, 0.5, 0, True)
    kept = set(ids)
    assert 1 not in kept                                     # class 0 = ignore label
    assert 11 not in kept and 12 not in kept                 # count_th / overlap_th
    assert kept and kept <= {2, 3, 4, 5, 6} and set(np.unique(seg)) == kept | {0}
    # same result from (pred, area histogram, sigmoid-mask histogram) -> integer stage
    r = F.interpolate(logits.float(), size=(96, 120), mode="bilinear", align_corners=False)[0]
    pred = r.argmax(0)
    pred[F.softmax(r, 0).max(0)[0] < 0.5] = -1
    pred = pred.numpy()
    area = np.bincount(pred[pred >= 0].ravel(), minlength=128)
    orig = (torch.sigmoid(r) >= 0.5).sum((1, 2)).numpy()
    seg2, ids2 = orc.panoptic_filter(pred, area, orig, None, 300, 0.5, 0)
    assert ids2 == ids and (seg2 == seg).all()


def test_sample_guidance_and_ddpm_seed():
    """sample(): the doubled batch + guidance combine (trainers_ldm_cond.py:1098-1146) reduces to the plain loop at
    guidance 1 with identical halves; the DDPM extension draws its noise from seed + 1."""
    torch.manual_seed(0)
    u = orc.UNet(**dict(TINY, cross_attention_dim=32))
    u.modify_encoder(in_channels=8, cond_channels=0)
    u = u.eval()
    rgb = torch.randn(1, 4, 8, 8) * 0.5
    enc = torch.randn(1, 5, 32)
    enc2 = torch.cat([enc, enc])
    s = orc.DDIMNoiseScheduler(**SCHED_KW)

    class Fixed(torch.nn.Module):            # the plain loop with the same (conditional) states on a single batch
        def forward(self, x, t, encoder_hidden_states=None):
            return u(x, t, encoder_hidden_states=enc)
    a = orc.sample(u, s, rgb, 3, seed=5, self_condition=False, encoder_hidden_states=enc2, guidance_scale=1.0)
    b = orc.sample(Fixed(), s, rgb, 3, seed=5, self_condition=False)
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
    c = orc.sample(u, s, rgb, 3, seed=5, self_condition=False, encoder_hidden_states=enc2, guidance_scale=7.5)
    assert c.shape == a.shape and torch.isfinite(c).all()
    d1 = orc.sample(Fixed(), s, rgb, 4, seed=5, self_condition=False, ddpm=True)
    d2 = orc.sample(Fixed(), s, rgb, 4, seed=6, noise=torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(5)),
                    self_condition=False, ddpm=True)
    assert not torch.allclose(d1, d2)            # same initial noise, different ancestral stream
