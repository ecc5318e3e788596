"""`ldmseg.models.GeneralVAEImage` / `GeneralVAESeg` -- drop-ins for
/root/reference/ldmseg/models/vae.py:36-39 and :42-307.

Same names, constructor arguments, attributes (`scaling_factor`, `downsample_factor`,
`interpolation_factor`), methods (`encode(x).latent_dist.{mode,sample}`, `decode(z, interpolate)`,
`set_scaling_factor`, `load_pretrained`) and state-dict keys.  encode / decode execute on the
sm_100a kernels (ldmseg.engine.vae_engine); there is no PyTorch / CPU fallback.

Scope: the default 'gaussian' parametrization with num_mid_blocks = 0; the
discrete / gumbel / bottleneck AE variants are stage-1 training options and raise
NotImplementedError.
"""
from __future__ import annotations

import json
import os
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

from ldmseg.utils import OutputDict
from . import _blocks as B


class RangeDict(OutputDict):
    min: torch.Tensor
    max: torch.Tensor


class VAEOutput(OutputDict):
    sample: torch.Tensor
    posterior: torch.Tensor


class EncoderOutput(OutputDict):
    latent_dist: torch.Tensor


class DiagonalGaussianDistribution(object):
    """Posterior over latents given the encoder moments (vae.py:370-424).  `mode()` is the mean;
    `sample()` draws with torch's generator (host-side choice of noise, element-wise math only)."""

    def __init__(self, parameters: torch.Tensor, clamp_output: bool = False, act_fn: str = "none"):
        self.parameters = parameters
        if clamp_output:
            parameters = torch.clamp(parameters, -5.0, 5.0)
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.mean = self.to_range(self.mean, act_fn)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        self.clamp_output = clamp_output
        self.act_fn = act_fn

    def to_range(self, x, act_fn):
        if act_fn == "sigmoid":
            return 2 * torch.sigmoid(x) - 1
        if act_fn == "tanh":
            return torch.tanh(x)
        if act_fn == "clip":
            return torch.clamp(x, -1, 1)
        if act_fn == "none":
            return x
        raise NotImplementedError

    def mode(self):
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.FloatTensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def kl(self):
        return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])

    def get_range(self):
        return RangeDict(min=self.mean.min(), max=self.mean.max())


# ================================================================================================
class GeneralVAEImage(nn.Module):
    """AutoencoderKL (SD-v1 `vae/config.json`) -- encoder half + quant_conv.  tools/main_ldm.py:138
    replaces the decoder by nn.Identity, so it is not built (decoder.* keys of a checkpoint are
    ignored on load)."""

    def __init__(self, in_channels: int = 3, out_channels: int = 3, latent_channels: int = 4,
                 block_out_channels: Tuple[int] = (128, 256, 512, 512), layers_per_block: int = 2,
                 norm_num_groups: int = 32, scaling_factor: float = 0.18215, **_):
        super().__init__()
        self.encoder = B.Encoder(in_channels, latent_channels, tuple(block_out_channels), layers_per_block,
                                 norm_num_groups)
        self.decoder = nn.Identity()
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.scaling_factor = scaling_factor
        self.config = B.make_config(in_channels=in_channels, latent_channels=latent_channels,
                                    block_out_channels=tuple(block_out_channels),
                                    layers_per_block=layers_per_block, norm_num_groups=norm_num_groups,
                                    scaling_factor=scaling_factor)
        self._engine = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, subfolder: Optional[str] = None,
                        cache_dir: Optional[str] = None, **kwargs):
        root = pretrained_model_name_or_path
        d = os.path.join(root, subfolder) if (root and subfolder) else root
        cfg, state = {}, None
        if d and os.path.isdir(d):
            cj = os.path.join(d, "config.json")
            if os.path.exists(cj):
                with open(cj) as f:
                    cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
            st = os.path.join(d, "diffusion_pytorch_model.safetensors")
            bn = os.path.join(d, "diffusion_pytorch_model.bin")
            if os.path.exists(st):
                from safetensors.torch import load_file
                state = load_file(st)
            elif os.path.exists(bn):
                state = torch.load(bn, map_location="cpu")
        else:
            print(f"[ldmseg_b200] '{root}' is not a local directory: building the SD-v1 AutoencoderKL encoder "
                  "from its config with random-init weights (no network access)")
        cfg.update(kwargs)
        model = cls(**cfg)
        if state is not None:
            state = {k: v for k, v in state.items() if not k.startswith(("decoder.", "post_quant_conv."))}
            model.load_state_dict(state, strict=True)
        return model

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def set_scaling_factor(self, scaling_factor):
        self.scaling_factor = scaling_factor

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _get_engine(self):
        if self._engine is None:
            from ldmseg.engine.vae_engine import ImageEncoderEngine
            self._engine = ImageEncoderEngine(self)
        return self._engine

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        if not x.is_cuda:
            raise RuntimeError("ldmseg_b200.GeneralVAEImage.encode needs CUDA tensors (no CPU fallback)")
        moments = self._get_engine().encode(x).to(x.dtype)
        return EncoderOutput(latent_dist=DiagonalGaussianDistribution(moments))

    def forward(self, *a, **k):
        raise NotImplementedError("the AutoencoderKL decoder is replaced by nn.Identity in LDMSeg (tools/main_ldm.py:138)")


# ================================================================================================
class LayerNorm2d(nn.Module):
    """Parameter container of the per-pixel channel LayerNorm (vae.py:309-322); executed fused with the
    ConvTranspose pixel shuffle + SiLU by ldmseg_convt_shuffle_ln."""

    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps


class GeneralVAESeg(nn.Module):
    def __init__(
        self,
        in_channels: int = 3,
        int_channels: int = 256,
        out_channels: int = 128,
        block_out_channels: Tuple[int] = (32, 64, 128, 256),
        latent_channels: int = 4,
        norm_num_groups: int = 32,
        scaling_factor: float = 0.18215,
        pretrained_path: Optional[str] = None,
        encoder: Optional[nn.Module] = None,
        num_mid_blocks: int = 0,
        num_latents: int = 2,
        num_upscalers: int = 1,
        upscale_channels: int = 256,
        parametrization: str = "gaussian",
        fuse_rgb: bool = False,
        resize_input: bool = False,
        act_fn: str = "none",
        clamp_output: bool = False,
        freeze_codebook: bool = False,
        skip_encoder: bool = False,
    ) -> None:
        super().__init__()
        assert parametrization in ["gaussian", "discrete_gumbel_softmax", "discrete_codebook", "auto"]
        if parametrization != "gaussian":
            raise NotImplementedError(f"parametrization '{parametrization}' is a stage-1 training variant (not built)")
        if num_mid_blocks > 0 or resize_input or skip_encoder or encoder is not None:
            raise NotImplementedError("only the default shallow encoder/decoder (num_mid_blocks=0) is built")
        assert num_latents in [1, 2, 32]
        block_out_channels = tuple(block_out_channels)
        self.enable_mid_block = False
        self.num_mid_blocks = num_mid_blocks
        self.downsample_factor = 2 ** (len(block_out_channels) - 1)
        self.interpolation_factor = self.downsample_factor // (2 ** num_upscalers)
        if fuse_rgb:
            in_channels += 3
        # encoder (vae.py:174-244): Sequential indices must match the reference's state-dict
        enc = [nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1), nn.SiLU()]
        for i in range(len(block_out_channels) - 1):
            ci, co = block_out_channels[i], block_out_channels[i + 1]
            enc += [nn.Conv2d(ci, ci, 3, padding=1), nn.Conv2d(ci, co, 3, padding=1, stride=2), nn.SiLU()]
        enc += [nn.Conv2d(block_out_channels[-1], int_channels, 3, padding=1), nn.Identity(),
                nn.GroupNorm(num_channels=int_channels, num_groups=norm_num_groups, eps=1e-6), nn.SiLU(),
                nn.Conv2d(int_channels, latent_channels * num_latents, 3, padding=1)]
        self.encoder = nn.Sequential(*enc)
        # decoder (vae.py:123-172)
        dec = [nn.Conv2d(latent_channels, int_channels, 3, padding=1), nn.Identity()]
        for i in range(num_upscalers):
            dec += [nn.ConvTranspose2d(int_channels if i == 0 else upscale_channels, upscale_channels, 2, stride=2),
                    LayerNorm2d(upscale_channels), nn.SiLU()]
        dec += [nn.GroupNorm(norm_num_groups, upscale_channels), nn.SiLU(),
                nn.Conv2d(upscale_channels, out_channels, 3, padding=1)]
        self.decoder = nn.Sequential(*dec)
        self.scaling_factor = scaling_factor
        self.gradient_checkpoint = False
        self.parametrization = parametrization
        self.num_latents = num_latents
        self.act_fn = act_fn
        self.clamp_output = clamp_output
        self._engine = None
        if pretrained_path is not None:
            self.load_pretrained(pretrained_path)

    def enable_gradient_checkpointing(self):
        raise NotImplementedError("Gradient checkpointing not implemented for a shallow VAE")

    def load_pretrained(self, pretrained_path):
        data = torch.load(pretrained_path, map_location="cpu")
        data["vae"] = {k.replace("module.", ""): v for k, v in data["vae"].items()}
        msg = self.load_state_dict(data["vae"], strict=True)
        print(f"Loaded pretrained VAE from {pretrained_path} with message {msg}")

    def freeze_layers(self):
        raise NotImplementedError

    def freeze_encoder(self):
        self.encoder.requires_grad_(False)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def _get_engine(self):
        if self._engine is None:
            from ldmseg.engine.vae_engine import SegVAEEngine
            self._engine = SegVAEEngine(self)
        return self._engine

    @torch.no_grad()
    def encode(self, semseg):
        if not semseg.is_cuda:
            raise RuntimeError("ldmseg_b200.GeneralVAESeg.encode needs CUDA tensors (no CPU fallback)")
        moments = self._get_engine().encode(semseg)
        return EncoderOutput(latent_dist=DiagonalGaussianDistribution(
            moments, clamp_output=self.clamp_output, act_fn=self.act_fn))

    @torch.no_grad()
    def decode(self, z, interpolate=True):
        if not z.is_cuda:
            raise RuntimeError("ldmseg_b200.GeneralVAESeg.decode needs CUDA tensors (no CPU fallback)")
        if interpolate and self.interpolation_factor != 2:
            raise NotImplementedError("only the x2 bilinear tail of the released config (num_upscalers=2) is built")
        return self._get_engine().decode(z, interpolate=interpolate)

    @torch.no_grad()
    def decode_ids(self, z):
        """Extension (fast path): argmax class ids (uint8) and max softmax probability at full resolution
        without materialising the logits (what trainers_ldm_cond.py:428-433 computes from them)."""
        return self._get_engine().decode_ids(z)

    @torch.no_grad()
    def decode_panoptic(self, z, sizes, crops=None, mask_th: float = 0.5, count_th: int = 512,
                        overlap_th: float = 0.5, ignore_label: int = 0, threshold_output: bool = True):
        """Extension: decode + the per-image post-processing of `TrainerDiffusion.compute_pq`
        (trainers_ldm_cond.py:1243-1313: resize to the original size, argmax, mask_th, count_th / overlap_th
        segment filtering) on the device.  Returns [(panoptic_seg uint8 [h, w] CPU tensor, segment ids)] -- the
        reference's `processed_results[i]["panoptic_seg"]` with `segments_info` reduced to its ids."""
        if not z.is_cuda:
            raise RuntimeError("ldmseg_b200.GeneralVAESeg.decode_panoptic needs CUDA tensors (no CPU fallback)")
        ids, keep = self._get_engine().decode_panoptic(z, sizes, crops, 1.0, mask_th, count_th, overlap_th,
                                                       ignore_label, threshold_output)
        ids, keep = ids.cpu(), keep.cpu()            # uint8 ids + a 128-entry table per image: the only D2H traffic
        out = []
        for i, (h, w) in enumerate(sizes):
            out.append((ids[i, : h * w].view(h, w), [int(c) + 1 for c in keep[i].nonzero().flatten()]))
        return out

    def forward(self, sample, sample_posterior: bool = True, return_dict: bool = True, generator=None,
                rgb_sample=None, valid_mask=None):
        x = sample
        if rgb_sample is not None:
            x = torch.cat([x, rgb_sample], dim=1)
        posterior = self.encode(x).latent_dist
        z = posterior.sample(generator=generator) if sample_posterior else posterior.mode()
        if valid_mask is not None:
            z = z * valid_mask[:, None]
        dec = self.decode(z, interpolate=False)
        if not return_dict:
            return (dec,)
        return VAEOutput(sample=dec, posterior=posterior)
