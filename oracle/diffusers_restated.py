"""ORACLE -- test infrastructure only.  Never imported by the product path.

CPU fp32 restatement (plain PyTorch) of the parts of HuggingFace `diffusers==0.16.1`
(/root/reference/data/environment.yml:50) that LDMSeg's sampling hot path executes.  diffusers is a
third-party dependency that is NOT vendored in /root/reference and NOT installable here (no
network), so its published architecture is restated from the SD-v1 `unet/config.json` /
`vae/config.json` and the 0.16.1 module structure, following SURVEY.md Appendix A.  Anchors:

  * exact parameter counts: stock UNet2DConditionModel 859 520 964; after LDMSeg surgery
    (cross-attention removed, 12-ch conv_in) 815 556 484; AutoencoderKL encoder+quant_conv
    34 163 592 (asserted in tests/test_oracle.py);
  * diffusers state-dict key names (so reference checkpoints `data['unet']` load strictly);
  * the reference's own call sites: ldmseg/models/unet.py:13-14,24,281-436 (UNet2DConditionModel
    attributes used by UNet.forward), ldmseg/models/vae.py:15-16,36 (AutoencoderKL),
    tools/main_ldm.py:137-160.

PARITY STATUS: the reference has no tests and no golden vectors for this path, and
diffusers itself cannot be run here, so the numerical details of this file are "parity unpinned"
against real diffusers; they are pinned only structurally (counts, key names) and against the
reference's own importable files (scheduler, seg VAE) via tests/golden.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# embeddings (App. A.2 items 1-2)
class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        ang = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------------------------
# resnet / sampling blocks
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int] = 1280,
                 groups: int = 32, eps: float = 1e-5, output_scale_factor: float = 1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = (nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0)
                              if in_channels != out_channels else None)
        self.output_scale_factor = output_scale_factor

    def forward(self, x, temb=None):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        if temb is not None and self.time_emb_proj is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


# --------------------------------------------------------------------------------------------
# attention (App. A.2: Transformer2DModel / BasicTransformerBlock / Attention / GEGLU)
class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None):
        b, n, _ = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.to_q(hidden_states)
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        d = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * d)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int]):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        if cross_attention_dim is not None:
            self.norm2 = nn.LayerNorm(dim)
            self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        else:
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **_):
        hidden_states = self.attn1(self.norm1(hidden_states)) + hidden_states
        if self.attn2 is not None:  # LDMSeg sets attn2 = norm2 = None (unet.py:83-105)
            hidden_states = self.attn2(self.norm2(hidden_states),
                                       encoder_hidden_states=encoder_hidden_states) + hidden_states
        return self.ff(self.norm3(hidden_states)) + hidden_states


class Transformer2DModel(nn.Module):
    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: Optional[int],
                 groups: int = 32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, hidden_states, encoder_hidden_states=None, **kw):
        b, _, h, w = hidden_states.shape
        residual = hidden_states
        x = self.proj_in(self.norm(hidden_states))
        inner = x.shape[1]
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, inner)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states)
        x = x.reshape(b, h, w, inner).permute(0, 3, 1, 2).contiguous()
        return SimpleNamespace(sample=self.proj_out(x) + residual)


# --------------------------------------------------------------------------------------------
# UNet blocks
class CrossAttnDownBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, heads, cross_attention_dim,
                 add_downsample, num_layers=2, eps=1e-5, groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, groups, eps)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim, groups)
            for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, padding=1)]) if add_downsample else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None):
        outs = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class DownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, add_downsample, num_layers=2, eps=1e-5,
                 groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, groups, eps)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, padding=1)]) if add_downsample else None

    def forward(self, hidden_states, temb=None):
        outs = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class UNetMidBlock2DCrossAttn(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, heads, cross_attention_dim, eps=1e-5, groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps),
                                      ResnetBlock2D(in_channels, in_channels, temb_channels, groups, eps)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(heads, in_channels // heads, in_channels, cross_attention_dim, groups)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


def _up_resnets(in_channels, out_channels, prev_output_channel, temb_channels, num_layers, eps, groups):
    mods = []
    for i in range(num_layers):
        skip = in_channels if i == num_layers - 1 else out_channels
        rin = prev_output_channel if i == 0 else out_channels
        mods.append(ResnetBlock2D(rin + skip, out_channels, temb_channels, groups, eps))
    return nn.ModuleList(mods)


class UpBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, add_upsample,
                 num_layers=3, eps=1e-5, groups=32):
        super().__init__()
        self.resnets = _up_resnets(in_channels, out_channels, prev_output_channel, temb_channels,
                                   num_layers, eps, groups)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None):
        for resnet in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = resnet(torch.cat([hidden_states, res], dim=1), temb)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


class CrossAttnUpBlock2D(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, heads,
                 cross_attention_dim, add_upsample, num_layers=3, eps=1e-5, groups=32):
        super().__init__()
        self.resnets = _up_resnets(in_channels, out_channels, prev_output_channel, temb_channels,
                                   num_layers, eps, groups)
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim, groups)
            for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = resnet(torch.cat([hidden_states, res], dim=1), temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states).sample
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states, upsample_size)
        return hidden_states


SD_V1_UNET_CONFIG = dict(
    in_channels=4, out_channels=4, sample_size=64, block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2, attention_head_dim=8, cross_attention_dim=768, norm_num_groups=32, norm_eps=1e-5,
    flip_sin_to_cos=True, freq_shift=0,
    down_block_types=("CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
    up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"),
)


class UNet2DConditionModel(nn.Module):
    """Base class in the shape the reference's `UNet` subclass expects (unet.py:24,281-436):
    attributes time_proj, time_embedding, conv_in, down_blocks, mid_block, up_blocks, conv_norm_out,
    conv_act, conv_out, encoder_hid_proj, config, dtype."""

    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD_V1_UNET_CONFIG)
        cfg.update(overrides)
        self.config = SimpleNamespace(**cfg)
        boc = tuple(cfg["block_out_channels"])
        heads = cfg["attention_head_dim"]
        xdim = cfg["cross_attention_dim"]
        groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
        temb = boc[0] * 4
        self.time_proj = Timesteps(boc[0], cfg["flip_sin_to_cos"], cfg["freq_shift"])
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.encoder_hid_proj = None
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, kind in enumerate(cfg["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            final = i == len(boc) - 1
            if kind == "CrossAttnDownBlock2D":
                blk = CrossAttnDownBlock2D(in_ch, out_ch, temb, heads, xdim, not final,
                                           cfg["layers_per_block"], eps, groups)
            else:
                blk = DownBlock2D(in_ch, out_ch, temb, not final, cfg["layers_per_block"], eps, groups)
            self.down_blocks.append(blk)
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb, heads, xdim, eps, groups)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        out_ch = rev[0]
        for i, kind in enumerate(cfg["up_block_types"]):
            prev = out_ch
            out_ch = rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            final = i == len(boc) - 1
            if kind == "CrossAttnUpBlock2D":
                blk = CrossAttnUpBlock2D(in_ch, out_ch, prev, temb, heads, xdim, not final,
                                         cfg["layers_per_block"] + 1, eps, groups)
            else:
                blk = UpBlock2D(in_ch, out_ch, prev, temb, not final, cfg["layers_per_block"] + 1, eps, groups)
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_pretrained(cls, path=None, subfolder=None, cache_dir=None, **kw):
        # no network / no checkpoints in this environment: a from-config random-init model
        return cls(**kw)

    def enable_gradient_checkpointing(self):
        pass

    def forward(self, sample, timestep, encoder_hidden_states=None, **_):
        """Stock diffusers forward (the LDMSeg subclass overrides it; kept for the stock model)."""
        t = timestep if torch.is_tensor(timestep) else torch.tensor([timestep], device=sample.device)
        t = t.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(t).to(self.dtype))
        sample = self.conv_in(sample)
        res = (sample,)
        for blk in self.down_blocks:
            if getattr(blk, "has_cross_attention", False):
                sample, r = blk(sample, emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample, r = blk(sample, emb)
            res += r
        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            r, res = res[-n:], res[:-n]
            if getattr(blk, "has_cross_attention", False):
                sample = blk(sample, r, emb, encoder_hidden_states=encoder_hidden_states)
            else:
                sample = blk(sample, r, emb)
        return SimpleNamespace(sample=self.conv_out(self.conv_act(self.conv_norm_out(sample))))


# --------------------------------------------------------------------------------------------
# AutoencoderKL encoder (App. A.4)
class AttentionBlock(nn.Module):
    """diffusers 0.16.1 single-head spatial attention (keys group_norm/query/key/value/proj_attn)."""

    def __init__(self, channels: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.query(t), self.key(t), self.value(t)
        scores = torch.baddbmm(torch.empty(b, h * w, h * w, dtype=q.dtype, device=q.device), q,
                               k.transpose(-1, -2), beta=0, alpha=1.0 / math.sqrt(c))
        probs = torch.softmax(scores.float(), dim=-1).type(scores.dtype)
        o = self.proj_attn(torch.bmm(probs, v))
        return o.transpose(-1, -2).reshape(b, c, h, w) + x


class UNetMidBlock2D(nn.Module):
    def __init__(self, in_channels: int, resnet_eps: float = 1e-6, resnet_groups: int = 32,
                 temb_channels=None, add_attention: bool = True, **_):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels, in_channels, temb_channels, resnet_groups, resnet_eps),
            ResnetBlock2D(in_channels, in_channels, temb_channels, resnet_groups, resnet_eps)])
        self.attentions = nn.ModuleList(
            [AttentionBlock(in_channels, resnet_groups, resnet_eps) if add_attention else None])

    def forward(self, x, temb=None):
        x = self.resnets[0](x, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            if attn is not None:
                x = attn(x)
            x = resnet(x, temb)
        return x


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, add_downsample, num_layers=2, eps=1e-6, groups=32):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, None, groups, eps)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, padding=0)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x


class Encoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, groups=32):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        oc = block_out_channels[0]
        for i, c in enumerate(block_out_channels):
            ic, oc = oc, c
            self.down_blocks.append(
                DownEncoderBlock2D(ic, oc, i != len(block_out_channels) - 1, layers_per_block, 1e-6, groups))
        self.mid_block = UNetMidBlock2D(block_out_channels[-1], 1e-6, groups, None)
        self.conv_norm_out = nn.GroupNorm(groups, block_out_channels[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(block_out_channels[-1], 2 * out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.mean + self.std * noise


class AutoencoderKL(nn.Module):
    """Encoder half only (+quant_conv): tools/main_ldm.py:138 replaces the decoder by nn.Identity."""

    def __init__(self, in_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, norm_num_groups=32, scaling_factor=0.18215, **_):
        super().__init__()
        self.encoder = Encoder(in_channels, latent_channels, tuple(block_out_channels), layers_per_block,
                               norm_num_groups)
        self.decoder = nn.Identity()
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.scaling_factor = scaling_factor
        self.config = SimpleNamespace(in_channels=in_channels, latent_channels=latent_channels,
                                      block_out_channels=tuple(block_out_channels),
                                      scaling_factor=scaling_factor)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @classmethod
    def from_pretrained(cls, path=None, subfolder=None, cache_dir=None, **kw):
        return cls(**kw)

    def encode(self, x):
        moments = self.quant_conv(self.encoder(x))
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(moments))


# stubs for names the reference imports but the sampling path never calls
class EMAModel:  # diffusers.training_utils.EMAModel (unet.py:14) -- training only
    def __init__(self, *a, **k):
        raise NotImplementedError("EMAModel is training-only and out of scope for the sampling oracle")
