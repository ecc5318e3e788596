"""Parameter containers with the diffusers-0.16.1 module / state-dict names.

These modules hold the fp32 parameters (so `state_dict()` / `load_state_dict()` are key-compatible
with reference checkpoints) and describe the architecture to the CUDA engine.
They carry NO arithmetic: the forward pass of the model is executed by ldmseg.engine on the
hand-written sm_100a kernels, never by PyTorch ops.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch.nn as nn


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers are not callable
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of the B200 engine; call the owning model "
            "(UNet / GeneralVAEImage / GeneralVAESeg) instead -- there is no PyTorch fallback path.")


class ResnetBlock2D(_Container):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int], groups: int,
                 eps: float):
        super().__init__()
        self.in_channels, self.out_channels, self.eps, self.groups = in_channels, out_channels, eps, groups
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class Attention(_Container):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv, inner, bias=False)
        self.to_v = nn.Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])


class GEGLU(_Container):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(_Container):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])


class BasicTransformerBlock(_Container):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int]):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim) if cross_attention_dim is not None else None
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head) if cross_attention_dim is not None else None
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)


class Transformer2DModel(_Container):
    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: Optional[int],
                 groups: int):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.channels = heads, dim_head, in_channels
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)


class Downsample2D(_Container):
    def __init__(self, channels: int, padding: int):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)


class Upsample2D(_Container):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)


class CrossAttnDownBlock2D(_Container):
    has_cross_attention = True

    def __init__(self, cin, cout, temb, heads, xdim, add_downsample, num_layers, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps)
                                      for i in range(num_layers)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, xdim, groups)
                                         for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 1)]) if add_downsample else None


class DownBlock2D(_Container):
    def __init__(self, cin, cout, temb, add_downsample, num_layers, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups, eps)
                                      for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 1)]) if add_downsample else None


class UNetMidBlock2DCrossAttn(_Container):
    has_cross_attention = True

    def __init__(self, channels, temb, heads, xdim, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, temb, groups, eps) for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, channels // heads, channels, xdim, groups)])


def _up_resnets(cin, cout, prev, temb, num_layers, eps, groups):
    mods = []
    for i in range(num_layers):
        skip = cin if i == num_layers - 1 else cout
        rin = prev if i == 0 else cout
        mods.append(ResnetBlock2D(rin + skip, cout, temb, groups, eps))
    return nn.ModuleList(mods)


class UpBlock2D(_Container):
    def __init__(self, cin, cout, prev, temb, add_upsample, num_layers, eps, groups):
        super().__init__()
        self.resnets = _up_resnets(cin, cout, prev, temb, num_layers, eps, groups)
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None


class CrossAttnUpBlock2D(_Container):
    has_cross_attention = True

    def __init__(self, cin, cout, prev, temb, heads, xdim, add_upsample, num_layers, eps, groups):
        super().__init__()
        self.resnets = _up_resnets(cin, cout, prev, temb, num_layers, eps, groups)
        self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, xdim, groups)
                                         for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None


class TimestepEmbedding(_Container):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)


class Timesteps(_Container):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift


# ---- AutoencoderKL encoder ---------------------------------------------------------------------
class AttentionBlock(_Container):
    def __init__(self, channels: int, groups: int, eps: float):
        super().__init__()
        self.channels = channels
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)


class UNetMidBlock2D(_Container):
    def __init__(self, in_channels: int, resnet_eps: float = 1e-6, resnet_groups: int = 32,
                 temb_channels=None, add_attention: bool = True, **_):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, resnet_groups, resnet_eps)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList([AttentionBlock(in_channels, resnet_groups, resnet_eps)
                                         if add_attention else None])


class DownEncoderBlock2D(_Container):
    def __init__(self, cin, cout, add_downsample, num_layers, eps, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, groups, eps)
                                      for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, 0)]) if add_downsample else None


class Encoder(_Container):
    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block, groups):
        super().__init__()
        boc = tuple(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        oc = boc[0]
        for i, c in enumerate(boc):
            ic, oc = oc, c
            self.down_blocks.append(DownEncoderBlock2D(ic, oc, i != len(boc) - 1, layers_per_block, 1e-6, groups))
        self.mid_block = UNetMidBlock2D(boc[-1], 1e-6, groups, None)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)


def make_config(**kw) -> SimpleNamespace:
    return SimpleNamespace(**kw)
