"""`ldmseg.models.UNet` -- drop-in for /root/reference/ldmseg/models/unet.py:24-436.

Same class name, constructor route (`from_pretrained`), surgery helpers (`remove_cross_attention`,
`modify_encoder`, `freeze_layers`, ...) and `forward` signature / return type, same diffusers-0.16.1
state-dict keys.  The arithmetic of `forward` runs on ldmseg.engine.UNetEngine (hand-written
sm_100a kernels through the C ABI); there is no PyTorch / CPU fallback.

Scope: the default sampling configuration: cross-attention removed,
conv_in widened to 8 (+cond) channels, no dual encoder / separate conv / dropout.  Options of the
reference outside that configuration keep their signature and raise NotImplementedError.
"""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from ldmseg.utils import OutputDict
from . import _blocks as B

SD_V1_UNET_CONFIG = dict(
    in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2, attention_head_dim=8, cross_attention_dim=768, norm_num_groups=32, norm_eps=1e-5,
    flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
)


class UNetOutput(OutputDict):
    sample: torch.FloatTensor


class UNet(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD_V1_UNET_CONFIG)
        cfg.update({k: v for k, v in overrides.items() if k in cfg})
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        self.config = B.make_config(**cfg)
        boc, heads, xdim = cfg["block_out_channels"], cfg["attention_head_dim"], cfg["cross_attention_dim"]
        groups, eps, lpb = cfg["norm_num_groups"], cfg["norm_eps"], cfg["layers_per_block"]
        temb = boc[0] * 4
        self.time_proj = B.Timesteps(boc[0], cfg["flip_sin_to_cos"], cfg["freq_shift"])
        self.time_embedding = B.TimestepEmbedding(boc[0], temb)
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.encoder_hid_proj = None
        self.down_blocks = nn.ModuleList()
        oc = boc[0]
        for i, kind in enumerate(cfg["down_block_types"]):
            ic, oc = oc, boc[i]
            final = i == len(boc) - 1
            if kind == "CrossAttnDownBlock2D":
                self.down_blocks.append(B.CrossAttnDownBlock2D(ic, oc, temb, heads, xdim, not final, lpb, eps, groups))
            elif kind == "DownBlock2D":
                self.down_blocks.append(B.DownBlock2D(ic, oc, temb, not final, lpb, eps, groups))
            else:
                raise NotImplementedError(kind)
        self.mid_block = B.UNetMidBlock2DCrossAttn(boc[-1], temb, heads, xdim, eps, groups)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        oc = rev[0]
        for i, kind in enumerate(cfg["up_block_types"]):
            prev, oc = oc, rev[i]
            ic = rev[min(i + 1, len(boc) - 1)]
            final = i == len(boc) - 1
            if kind == "CrossAttnUpBlock2D":
                self.up_blocks.append(B.CrossAttnUpBlock2D(ic, oc, prev, temb, heads, xdim, not final, lpb + 1, eps, groups))
            elif kind == "UpBlock2D":
                self.up_blocks.append(B.UpBlock2D(ic, oc, prev, temb, not final, lpb + 1, eps, groups))
            else:
                raise NotImplementedError(kind)
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)
        self._engine = None
        self._use_graph = os.environ.get("LDMSEG_CUDA_GRAPH", "1") != "0"

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, subfolder: Optional[str] = None,
                        cache_dir: Optional[str] = None, **kwargs):
        """Local directories in the diffusers layout (<path>/<subfolder>/config.json +
        diffusion_pytorch_model.{safetensors,bin}) are loaded; anything else (hub ids cannot be
        fetched: no network) yields the SD-v1 architecture with random-init weights."""
        root = pretrained_model_name_or_path
        d = os.path.join(root, subfolder) if (root and subfolder) else root
        cfg = {}
        state = None
        if d and os.path.isdir(d):
            cj = os.path.join(d, "config.json")
            if os.path.exists(cj):
                with open(cj) as f:
                    cfg = json.load(f)
            st = os.path.join(d, "diffusion_pytorch_model.safetensors")
            bn = os.path.join(d, "diffusion_pytorch_model.bin")
            if os.path.exists(st):
                from safetensors.torch import load_file
                state = load_file(st)
            elif os.path.exists(bn):
                state = torch.load(bn, map_location="cpu")
        else:
            print(f"[ldmseg_b200] '{root}' is not a local directory: building the SD-v1 UNet from its "
                  "config with random-init weights (no network access)")
        cfg.update(kwargs)
        model = cls(**cfg)
        if state is not None:
            model.load_state_dict(state, strict=True)
        return model

    # ------------------------------------------------------------------ nn.Module plumbing
    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    def invalidate_engine(self) -> None:
        """Drop packed weights / plans / graphs (call after changing parameters in place)."""
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def enable_gradient_checkpointing(self):
        pass  # inference-only engine: nothing to checkpoint

    # ------------------------------------------------------------------ reference helpers
    def define_dropout(self, dropout: float = 0.0, mode: str = "standard") -> None:
        if dropout <= 0.0:
            return
        raise NotImplementedError("input dropout is a training-time option (out of scope of the sampling path)")

    def define_learnable_embedding(self, in_channels, out_channels):
        assert self.encoder_hid_proj is None
        self.object_queries = nn.Embedding(in_channels, out_channels)
        self._engine = None

    def define_separate_encoder(self, add_adaptor: bool = False, init_mode_adaptor: str = "random"):
        raise NotImplementedError("the dual-encoder variant is off by default and not built (SURVEY.md §2 row 1)")

    def define_upscaler(self, num_classes: int = 128, norm_num_groups: int = 32, dim: int = 256) -> None:
        raise NotImplementedError("define_upscaler is never used by tools/main_ldm.py (out of scope)")

    def remove_cross_attention(self):
        blocks = [b for b in self.down_blocks if getattr(b, "has_cross_attention", False)]
        if hasattr(self, "mid_block"):
            blocks.append(self.mid_block)
        blocks += [b for b in self.up_blocks if getattr(b, "has_cross_attention", False)]
        for blk in blocks:
            for attn_block in blk.attentions:
                for tb in attn_block.transformer_blocks:
                    tb.attn2 = None
                    tb.norm2 = None
        self._engine = None

    def get_lr_func(self, name: str, lr_decay_rate: float = 1.0) -> float:
        if name.startswith("module."):
            name = name[len("module."):]
        if name.startswith("conv_in.") or name.startswith("down_blocks."):
            return lr_decay_rate
        return 1.0

    def modify_encoder_hidden_state_proj(self, in_channels: int, out_channels: int) -> None:
        self.encoder_hid_proj = nn.Linear(in_channels, out_channels)
        self._engine = None

    def modify_encoder(self, in_channels: int = 4, init_mode_seg: str = "copy", init_mode_image: str = "copy",
                       cond_channels: int = 0, init_mode_cond: str = "zero", separate_conv: bool = False,
                       separate_encoder: bool = False, add_adaptor: bool = False,
                       init_mode_adaptor: str = "random") -> None:
        assert in_channels in [4, 8], "in_channels must be 4 or 8"
        assert separate_conv + separate_encoder <= 1, "separate_conv and separate_encoder cannot both be True"
        if separate_conv or separate_encoder:
            raise NotImplementedError("separate_conv / separate_encoder variants are not built (off by default)")
        if in_channels != 8:
            return
        old = self.conv_in
        new = nn.Conv2d(in_channels + cond_channels, old.out_channels, kernel_size=old.kernel_size,
                        stride=old.stride, padding=old.padding, bias=old.bias is not None)
        new = new.to(device=old.weight.device, dtype=old.weight.dtype)

        def fill(sl, mode, what):
            if mode in ("copy", "div"):  # the reference's "div" does not divide (Q9): same as copy
                new.weight.data[:, sl].copy_(old.weight.data)
            elif mode == "mean":
                new.weight.data[:, sl].copy_(torch.mean(old.weight.data, dim=1, keepdim=True).repeat(1, 4, 1, 1))
            elif mode == "zero":
                new.weight.data[:, sl].zero_()
            elif mode == "random":
                pass
            else:
                raise NotImplementedError(f"init_mode {what} {mode} not implemented")

        fill(slice(0, 4), init_mode_seg, "seg")
        fill(slice(4, 8), init_mode_image, "seg")
        new.bias.data.copy_(old.bias.data)
        assert new.weight.data.shape == torch.Size([old.out_channels, 8 + cond_channels, 3, 3])
        if cond_channels > 0:
            if init_mode_cond == "zero":
                new.weight.data[:, 8:].zero_()
            elif init_mode_cond == "random":
                pass
            else:
                raise NotImplementedError(f"init_mode cond {init_mode_cond} not implemented")
        # the reference keeps `new_conv` registered next to `conv_in` (unet.py:182,233), so its state-dict
        # and checkpoints carry both key sets; mirror that for strict loading
        self.new_conv = new
        self.conv_in = self.new_conv
        self._engine = None

    def freeze_layers(self, layers: Tuple[str] = ("norm", "time_embedding")) -> None:
        for layer in layers:
            if layer == "norm":
                for m in self.modules():
                    if isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
                        m.requires_grad_(False)
            elif layer == "time_embedding":
                self.time_embedding.requires_grad_(False)
            elif layer in ("conv_in", "down_blocks"):
                pass  # only meaningful with the dual encoder
            else:
                raise NotImplementedError(f"layer {layer} not implemented")

    @torch.no_grad()
    def self_condition_estimate(self, scheduler, latents, rgb_latents, noise, timesteps,
                                encoder_hidden_states=None):
        """Extension: the no-grad forward of the training step (trainers_ldm_cond.py:813-831) on the sampling
        kernels -- add_noise at per-sample timesteps, UNet on cat[noisy, rgb, zeros], remove_noise.
        Returns (noisy_latents, pred, condition)."""
        if not latents.is_cuda:
            raise RuntimeError("ldmseg_b200.UNet.self_condition_estimate needs CUDA tensors (no CPU fallback)")
        return self._get_engine().self_condition_forward(scheduler, latents, rgb_latents, noise, timesteps,
                                                         encoder_hidden_states)

    # ------------------------------------------------------------------ forward
    def _get_engine(self):
        if self._engine is None:
            from ldmseg.engine.unet_engine import UNetEngine
            self._engine = UNetEngine(self)
        return self._engine

    @torch.no_grad()
    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor = None, class_labels: Optional[torch.Tensor] = None,
                timestep_cond: Optional[torch.Tensor] = None, attention_mask: Optional[torch.Tensor] = None,
                cross_attention_kwargs: Optional[Dict[str, Any]] = None,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = True,
                timestep_img: Optional[Union[torch.Tensor, float, int]] = None) -> Union[UNetOutput, Tuple]:
        if not sample.is_cuda:
            raise RuntimeError("ldmseg_b200.UNet.forward needs CUDA tensors: the sampling hot path has no CPU "
                               "fallback (the CPU oracle lives under oracle/ and is test infrastructure)")
        if down_block_additional_residuals is not None or mid_block_additional_residual is not None:
            raise NotImplementedError("additional residuals (dual encoder / ControlNet) are not built")
        if not torch.is_tensor(timestep):
            timestep = torch.tensor(timestep, device=sample.device)
        timesteps = timestep.to(sample.device).expand(sample.shape[0])  # accepts either device (Q5)
        out = self._get_engine().forward(sample.float(), timesteps, encoder_hidden_states)
        out = out.to(sample.dtype) if sample.dtype != torch.float32 else out
        if not return_dict:
            return (out,)
        return UNetOutput(sample=out)
