"""ORACLE -- test infrastructure only.  Never imported by the product path.

CPU fp32 restatement (plain PyTorch / numpy) of the in-repo half of LDMSeg's sampling hot path.
Every function cites the reference lines it follows (paths under /root/reference).

Pinned against the reference itself: tests/golden/*.npz were produced by importing the
reference's own `ldmseg/schedulers/ddim_scheduler.py`, `ldmseg/models/vae.py` and
`ldmseg/models/unet.py` (the latter on top of oracle.diffusers_restated, because diffusers is
absent) in the build container -- see oracle/make_golden.py -- and tests/test_oracle.py checks this
restatement against them.  The diffusers base classes underneath remain "parity unpinned"
(oracle/diffusers_restated.py header).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import diffusers_restated as dr


class OutputDict(OrderedDict):
    """ldmseg/utils/utils.py:26-31: item assignment mirrors to attributes."""

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)


# --------------------------------------------------------------------------------------------
class DDIMNoiseScheduler:
    """ldmseg/schedulers/ddim_scheduler.py:26-291 (deterministic DDIM, eta = 0)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 thresholding=False, dynamic_thresholding_ratio=0.995, clip_sample_range=1.0,
                 sample_max_value=1.0, weight="none", max_snr=5.0, device=None, verbose=True):
        T = num_train_timesteps
        if beta_schedule == "linear":                                   # :51-52
            betas = torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":                          # :53-57
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
        elif beta_schedule == "squaredcos_cap_v2":                      # :58-60, 138-153
            import math

            def abar(s):
                return math.cos((s + 0.008) / 1.008 * math.pi / 2) ** 2
            betas = torch.tensor([min(1 - abar((i + 1) / T) / abar(i / T), 0.999) for i in range(T)],
                                 dtype=torch.float32)
        elif beta_schedule == "sigmoid":                                # :61-64
            betas = torch.sigmoid(torch.linspace(-6, 6, T)) * (beta_end - beta_start) + beta_start
        else:
            raise NotImplementedError(beta_schedule)                    # :65-66
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)         # :68-69
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]  # :75
        snr = self.alphas_cumprod / (1 - self.alphas_cumprod)           # :97-117
        assert weight in ["inverse_log_snr", "max_clamp_snr", "linear", "fixed", "none"]
        if weight == "inverse_log_snr":
            w = torch.log(1.0 / snr).clamp(min=1)
            w = w / w[-1]
        elif weight == "max_clamp_snr":
            w = snr.clamp(max=max_snr) / snr
        elif weight == "fixed":
            w = snr.clone()
            w[: len(w) // 4] = 0.1
        elif weight == "linear":
            w = torch.arange(1, T + 1) / T
        else:
            w = torch.ones_like(snr)
        self.weights = w
        self.num_train_timesteps = T
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, T)[::-1].copy().astype(np.int64))  # :84
        self.clip_sample = clip_sample
        self.clip_sample_range = clip_sample_range
        self.prediction_type = prediction_type
        self.thresholding = thresholding
        self.steps_offset = steps_offset
        self.init_noise_sigma = 1.0

    def set_timesteps_inference(self, num_inference_steps, device=None, tmin=0):  # :119-131
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        self.steps_offset = ratio - 1                                   # overwrites the config (Q3)
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        ts = torch.from_numpy(ts).to(device) + self.steps_offset
        self.timesteps = ts[ts >= tmin]

    def move_timesteps_to(self, device):                                # :133-136
        self.timesteps = self.timesteps.to(device)

    def add_noise(self, original_samples, noise, timesteps, scale=1.0):  # :155-187
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        a = ac[timesteps.to(original_samples.device)].flatten()
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return (a ** 0.5).view(shape) * scale * original_samples + ((1 - a) ** 0.5).view(shape) * noise

    def remove_noise(self, noisy_samples, noise, timesteps, scale=1.0):  # :189-216
        ac = self.alphas_cumprod.to(device=noisy_samples.device, dtype=noisy_samples.dtype)
        a = ac[timesteps.to(noisy_samples.device)].flatten()
        shape = (-1,) + (1,) * (noisy_samples.dim() - 1)
        return (noisy_samples - ((1 - a) ** 0.5).view(shape) * noise) / ((a ** 0.5).view(shape) * scale)

    def step(self, model_output, timestep, sample, use_clipped_model_output=False):  # :218-269
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps             # :231
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod  # :234-235
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":                                         # :239-241
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        elif self.prediction_type == "sample":                                        # :242-244
            x0 = model_output
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        elif self.prediction_type == "v_prediction":                                  # :245-247
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise NotImplementedError
        if self.thresholding:                                                         # :252-253
            raise NotImplementedError
        if self.clip_sample:                                                          # :254-257
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        if use_clipped_model_output:                                                  # :259-261
            eps = (sample - a_t ** 0.5 * x0) / b_t ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev) ** 0.5 * eps                         # :264-267
        return OutputDict(prev_sample=prev, pred_original_sample=x0)

    # ---- extension (absent in the reference, SURVEY §8a row 11 ii): ancestral / DDPM step = DDIM eta=1
    def step_ddpm(self, model_output, timestep, sample, noise):
        t = int(timestep)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        var = (1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)
        sigma = var.clamp(min=0) ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev - sigma ** 2).clamp(min=0) ** 0.5 * model_output + sigma * noise
        return OutputDict(prev_sample=prev, pred_original_sample=x0)

    def __len__(self):
        return self.num_train_timesteps


# --------------------------------------------------------------------------------------------
class LayerNorm2d(nn.Module):
    """ldmseg/models/vae.py:309-322: per-pixel LayerNorm over channels, biased variance."""

    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x):
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        x = (x - mu) / torch.sqrt(var + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class GeneralVAESeg(nn.Module):
    """ldmseg/models/vae.py:42-271, default 'gaussian' parametrization, num_mid_blocks = 0."""

    def __init__(self, in_channels=7, int_channels=256, out_channels=128, block_out_channels=(32, 64, 128, 256),
                 latent_channels=4, norm_num_groups=32, scaling_factor=0.2, num_latents=2, num_upscalers=2,
                 upscale_channels=256, **_):
        super().__init__()
        boc = tuple(block_out_channels)
        self.downsample_factor = 2 ** (len(boc) - 1)                    # :71
        self.interpolation_factor = self.downsample_factor // (2 ** num_upscalers)  # :72
        enc = [nn.Conv2d(in_channels, boc[0], 3, padding=1), nn.SiLU()]  # :190-194
        for i in range(len(boc) - 1):                                    # :198-207
            enc += [nn.Conv2d(boc[i], boc[i], 3, padding=1),
                    nn.Conv2d(boc[i], boc[i + 1], 3, padding=1, stride=2), nn.SiLU()]
        enc += [nn.Conv2d(boc[-1], int_channels, 3, padding=1), nn.Identity(),        # :212-231
                nn.GroupNorm(num_channels=int_channels, num_groups=norm_num_groups, eps=1e-6), nn.SiLU(),
                nn.Conv2d(int_channels, latent_channels * num_latents, 3, padding=1)]  # :233-237
        self.encoder = nn.Sequential(*enc)
        dec = [nn.Conv2d(latent_channels, int_channels, 3, padding=1), nn.Identity()]  # :133,147
        for i in range(num_upscalers):                                   # :151-159
            dec += [nn.ConvTranspose2d(int_channels if i == 0 else upscale_channels, upscale_channels, 2, stride=2),
                    LayerNorm2d(upscale_channels), nn.SiLU()]
        dec += [nn.GroupNorm(norm_num_groups, upscale_channels), nn.SiLU(),            # :160-166
                nn.Conv2d(upscale_channels, out_channels, 3, padding=1)]
        self.decoder = nn.Sequential(*dec)
        self.scaling_factor = scaling_factor

    def encode(self, semseg):                                            # :252-265
        return OutputDict(latent_dist=dr.DiagonalGaussianDistribution(self.encoder(semseg)))

    def decode(self, z, interpolate=True):                               # :267-271
        x = self.decoder(z)
        if interpolate:
            x = F.interpolate(x, scale_factor=self.interpolation_factor, mode="bilinear", align_corners=False)
        return x


# --------------------------------------------------------------------------------------------
class UNet(dr.UNet2DConditionModel):
    """ldmseg/models/unet.py:24-436, default path only (no dual encoder / separate conv)."""

    def remove_cross_attention(self):                                    # :83-105
        blocks = [b for b in self.down_blocks if getattr(b, "has_cross_attention", False)]
        blocks += [self.mid_block]
        blocks += [b for b in self.up_blocks if getattr(b, "has_cross_attention", False)]
        for b in blocks:
            for attn in b.attentions:
                for tb in attn.transformer_blocks:
                    tb.attn2 = None
                    tb.norm2 = None

    def modify_encoder(self, in_channels=4, init_mode_seg="copy", init_mode_image="copy", cond_channels=0,
                       init_mode_cond="zero", **_):                      # :124-233 (in_channels == 8 branch)
        assert in_channels in [4, 8]
        if in_channels != 8:
            return
        old = self.conv_in
        new = nn.Conv2d(in_channels + cond_channels, old.out_channels, old.kernel_size, old.stride,
                        old.padding, bias=old.bias is not None)          # :182-183 (fresh default init)
        with torch.no_grad():
            if init_mode_seg == "copy":                                  # :185-186
                new.weight[:, :4].copy_(old.weight)
            elif init_mode_seg == "zero":
                new.weight[:, :4].zero_()
            if init_mode_image == "copy":                                # :199-200
                new.weight[:, 4:8].copy_(old.weight)
            elif init_mode_image == "zero":                              # :205-206
                new.weight[:, 4:8].zero_()
            new.bias.copy_(old.bias)                                     # :213
            if cond_channels > 0 and init_mode_cond == "zero":           # :222-224
                new.weight[:, 8:].zero_()
        self.new_conv = new   # the reference keeps this alias registered: its state-dict (and its
        self.conv_in = self.new_conv  # checkpoints) carry both new_conv.* and conv_in.* keys   # :182,233

    def forward(self, sample, timestep, encoder_hidden_states=None, return_dict=True, **_):  # :281-436
        timesteps = timestep.expand(sample.shape[0])                     # :302-303
        emb = self.time_embedding(self.time_proj(timesteps).to(dtype=self.dtype))  # :305-307
        sample = self.conv_in(sample)                                    # :357
        res = (sample,)                                                  # :360
        for blk in self.down_blocks:                                     # :361-373
            if getattr(blk, "has_cross_attention", False):
                sample, r = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample, r = blk(hidden_states=sample, temb=emb)
            res += r
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)  # :388-395
        for blk in self.up_blocks:                                       # :401-425
            n = len(blk.resnets)
            r, res = res[-n:], res[:-n]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=r,
                             encoder_hidden_states=encoder_hidden_states)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=r)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))  # :428-431
        if not return_dict:
            return (sample,)
        return OutputDict(sample=sample)


def build_ldmseg_unet(seed: int = 0, cond_channels: int = 4, image_init: str = "zero", **cfg) -> UNet:
    """tools/main_ldm.py:146-160 with random-init weights (no checkpoints offline): from-config UNet,
    'remove' descriptor (cross-attention stripped), conv_in widened to 8 + cond_channels."""
    torch.manual_seed(seed)
    unet = UNet(**cfg)
    unet.remove_cross_attention()                                        # descriptors.py:94-96
    unet.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image=image_init,
                        cond_channels=cond_channels, init_mode_cond="zero")
    return unet.eval()


# --------------------------------------------------------------------------------------------
@torch.no_grad()
def encode_inputs(images, vae_image, scaling_factor=0.18215):
    """ldmseg/trainers/trainers_ldm_cond.py:334-394, mode() branch, no resize."""
    images = 2.0 * images - 1.0                                          # :369
    latents = vae_image.encode(images).latent_dist.mode()                # :375
    return latents * scaling_factor                                      # :392


@torch.no_grad()
def sample(unet, scheduler, rgb_latents, num_inference_steps=50, seed=42, self_condition=True,
           noise=None, return_all=False, mask=None, known_latents=None, ddpm=False,
           encoder_hidden_states=None, guidance_scale=7.5):
    """ldmseg/trainers/trainers_ldm_cond.py:1045-1170.  `encoder_hidden_states` [2B, T, D] (uncond | cond rows,
    as built at :1104 / :1116) switches on the doubled batch + guidance combine of :1098-1123, 1143-1146.
    Extensions (labelled, absent in the reference): `mask`/`known_latents` inpainting blend,
    `ddpm` ancestral noise (generator seeded seed + 1)."""
    scheduler.set_timesteps_inference(num_inference_steps)               # :1078-1080
    b, _, L, _ = rgb_latents.shape
    gen = torch.Generator().manual_seed(seed) if seed is not None else None  # :1088
    latents = noise if noise is not None else torch.randn((b, 4, L, L), generator=gen)  # :1090
    multiplier = 2 if encoder_hidden_states is not None else 1          # :1098-1119
    latents = latents * scheduler.init_noise_sigma                       # :1121
    rgb_latents = torch.cat([rgb_latents] * multiplier)                  # :1123
    condition = torch.zeros_like(rgb_latents)                            # :1126
    fixed_noise = latents.clone()
    extra = torch.Generator().manual_seed((seed if seed is not None else 0) + 1)
    alls = []
    ts = scheduler.timesteps
    for i, t in enumerate(ts):                                           # :1127
        lmi = torch.cat([latents] * multiplier)                          # :1128
        if self_condition:
            inp = torch.cat([lmi, rgb_latents, condition], dim=1)        # :1133
        else:
            inp = torch.cat([lmi, rgb_latents], dim=1)                   # :1135
        eps = unet(inp.float(), t, encoder_hidden_states=encoder_hidden_states).sample    # :1141
        if multiplier > 1:                                               # :1143-1146
            eu, et = eps.chunk(2)
            eps = eu + guidance_scale * (et - eu)
        if ddpm and i != len(ts) - 1:
            z = torch.randn(latents.shape, generator=extra)
            out = scheduler.step_ddpm(eps, t, latents, z)
        else:
            out = scheduler.step(eps, t, latents)                        # :1150,1156,1159 (same args, Q2)
        if self_condition:
            condition = out.pred_original_sample                         # :1150
        last = i == len(ts) - 1
        latents = out.pred_original_sample if last else out.prev_sample  # :1154-1159 (Q1)
        if mask is not None:  # extension: paste the known region, noised to the next level
            if last:
                known = known_latents
            else:
                known = scheduler.add_noise(known_latents, fixed_noise, ts[i + 1].expand(b))
            latents = mask * known + (1 - mask) * latents
        if return_all:
            alls.append(latents)
    return torch.cat(alls, 0) if return_all else latents


@torch.no_grad()
def decode_latents(latents, vae_semseg, return_logits=True):
    """ldmseg/trainers/trainers_ldm_cond.py:396-442."""
    logits = vae_semseg.decode(latents * (1.0 / vae_semseg.scaling_factor)).float()  # :421-423
    if return_logits:
        return logits
    return torch.argmax(logits, dim=1)                                   # :428


# --------------------------------------------------------------------------------------------
def panoptic_postprocess(masks_logits, sizes, mask_th=0.5, count_th=512, overlap_th=0.5, ignore_label=0,
                         threshold_output=True, padding_masks=None):
    """ldmseg/trainers/trainers_ldm_cond.py:1261-1313, per image: (crop_padding :1264) -> bilinear resize of the
    logits to the original (h, w) (:1267-1272) -> argmax (:1275) -> softmax-max < mask_th -> -1 (:1276-1284) ->
    per-segment filtering by area `count_th` and by the overlap of the argmax region with the thresholded sigmoid
    mask (:1293-1313).  masks_logits f32 [B, C, H, W] (already at the RGB size, :1252-1257), sizes [(h, w)].
    Returns [(panoptic_seg int64 [h, w] with 0 = void, [segment ids])] -- `panoptic_pred + 1`, `segments_info` ids."""
    out = []
    for i, logit in enumerate(masks_logits):
        if padding_masks is not None:                                    # crop_padding, :1172-1178
            co = padding_masks[i].nonzero()
            y0, y1, x0, x1 = co[:, 0].min(), co[:, 0].max(), co[:, 1].min(), co[:, 1].max()
            logit = logit[:, y0:y1 + 1, x0:x1 + 1]
        h, w = sizes[i]
        r = F.interpolate(logit[None].float(), size=(h, w), mode="bilinear", align_corners=False)[0]   # :1267-1272
        pred = torch.argmax(r, dim=0)                                    # :1275
        if threshold_output:
            probs = F.softmax(r, dim=0).max(dim=0)[0]                    # :1277,1283
            pred[probs < mask_th] = -1                                   # :1284
        pred = pred.numpy()
        sig = torch.sigmoid(r).numpy()                                   # :1288-1289
        ids = []
        for label, count in zip(*np.unique(pred, return_counts=True)):   # :1293
            if count < count_th or label in {-1, ignore_label}:          # :1296-1298
                pred[pred == label] = -1
                continue
            original = sig[label] >= mask_th                             # :1301
            if (pred == label).sum() / original.sum() < overlap_th:      # :1302-1304
                pred[pred == label] = -1
                continue
            ids.append(int(label) + 1)                                   # :1306-1312
        out.append((pred + 1, ids))                                      # :1313
    return out


def panoptic_filter(pred, area, orig_area, mask_th_unused=None, count_th=512, overlap_th=0.5, ignore_label=0):
    """The integer half of the routine above on its own (bit-exact target of the CUDA filter kernel): `pred` int
    [h, w] with -1 = below threshold, `area[l]` = #pixels with pred == l, `orig_area[l]` = #pixels whose sigmoid
    for class l reaches mask_th.  Returns (pred + 1 with dropped segments zeroed, kept ids)."""
    pred = pred.copy()
    ids = []
    for label in range(len(area)):
        count = int(area[label])
        if count == 0:
            continue
        if count < count_th or label == ignore_label:
            pred[pred == label] = -1
            continue
        if count / float(orig_area[label]) < overlap_th:
            pred[pred == label] = -1
            continue
        ids.append(label + 1)
    return pred + 1, ids


@torch.no_grad()
def self_condition_estimate(unet, scheduler, latents, rgb_latents, noise, timesteps, encoder_hidden_states=None):
    """The training step's no-grad forward (ldmseg/trainers/trainers_ldm_cond.py:813-831): add_noise at per-sample
    timesteps, UNet on cat[noisy, rgb, zeros], remove_noise -> the self-conditioning estimate of x0."""
    noisy = scheduler.add_noise(latents, noise, timesteps)               # :820
    cond = torch.zeros_like(noisy)                                       # :826
    inp = torch.cat([noisy, rgb_latents, cond], dim=1)                   # :828
    pred = unet(inp, timesteps, encoder_hidden_states=encoder_hidden_states).sample   # :830
    return noisy, pred, scheduler.remove_noise(noisy, pred, timesteps)   # :831
