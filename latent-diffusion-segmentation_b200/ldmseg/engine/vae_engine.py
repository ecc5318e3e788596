"""Launch plans for the two ends of the sampling path:

  * ImageEncoderPlan -- AutoencoderKL encoder (+quant_conv) on RGB, replaces
    `GeneralVAEImage.encode(...).latent_dist` (diffusers AutoencoderKL, called from
    /root/reference/ldmseg/trainers/trainers_ldm_cond.py:371-375).
  * SegDecoderPlan  -- the shallow segmentation decoder, replaces `GeneralVAESeg.decode`
    (/root/reference/ldmseg/models/vae.py:123-172, 267-271) incl. the bilinear x2, plus a fused
    fast path (bilinear + argmax + max-prob) that never writes the 134 MB/sample logits.
  * SegEncoderPlan  -- `GeneralVAESeg.encode` (vae.py:174-265), needed for inpainting ground truth.

All convolutions / linears run on the tcgen05 implicit-GEMM kernel, norms on the fused kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn as nn

from ldmseg import _native as nat
from ldmseg import _pack as pk
from .plan import PlanBase, WeightsBase, _Layer


def _pad_cin(conv: nn.Conv2d, cpad: int) -> torch.Tensor:
    w = conv.weight.detach().float()
    wp = torch.zeros(w.shape[0], cpad, 3, 3, device=w.device)
    wp[:, : w.shape[1]] = w
    return wp


# ================================================================================================
class ImageEncoderWeights(WeightsBase):
    def __init__(self, vae, device):
        super().__init__(device)
        enc = vae.encoder
        self.groups = enc.conv_norm_out.num_groups
        self.in_channels = enc.conv_in.in_channels
        self.cin_pad = 16
        self._gemm("conv_in", pk.pack_conv3x3(_pad_cin(enc.conv_in, self.cin_pad)), enc.conv_in.bias,
                   enc.conv_in.out_channels)
        self.blocks = []
        for i, blk in enumerate(enc.down_blocks):
            chans = []
            for j, r in enumerate(blk.resnets):
                self._resnet(f"down{i}.res{j}", r, [r.in_channels])
                chans.append((r.in_channels, r.out_channels))
            has_down = blk.downsamplers is not None
            if has_down:
                self._conv_s2(f"down{i}.down", blk.downsamplers[0].conv)
            self.blocks.append((chans, has_down))
        mid = enc.mid_block
        for j, r in enumerate(mid.resnets):
            self._resnet(f"mid.res{j}", r, [r.in_channels])
        a = mid.attentions[0]
        self.mid_c = a.channels
        self._norm("mid.attn.norm", a.group_norm)
        self._gemm("mid.attn.q", pk.pack_linear(a.query.weight), a.query.bias, a.channels)
        self._gemm("mid.attn.k", pk.pack_linear(a.key.weight), a.key.bias, a.channels)
        # V is produced transposed (V^T = W_v x^T) so it can be the K-major B operand of P.V; its bias
        # is added after the product (rows of P sum to one)
        self.wv = self._dev(pk.pack_linear(a.value.weight), torch.bfloat16)
        self.bv = self._dev(a.value.bias.float())
        self._gemm("mid.attn.proj", pk.pack_linear(a.proj_attn.weight), a.proj_attn.bias, a.channels)
        self._norm("norm_out", enc.conv_norm_out)
        self._gemm("conv_out", pk.pack_conv3x3(enc.conv_out.weight), enc.conv_out.bias, enc.conv_out.out_channels)
        self._gemm("quant", pk.pack_linear(vae.quant_conv.weight), vae.quant_conv.bias, vae.quant_conv.out_channels)
        self.moment_channels = vae.quant_conv.out_channels


class ImageEncoderPlan(PlanBase):
    """x_in bf16 [nb*S*S, 16] (RGB in channels 0..2, already scaled to [-1,1]) -> moments f32 [nb*(S/8)^2, 8]."""

    def __init__(self, W: ImageEncoderWeights, nb: int, size: int, allow_split: bool = True):
        super().__init__(W, nb, allow_split)
        self.size = size
        self.x_in = torch.zeros(nb * size * size, W.cin_pad, device=self.device, dtype=torch.bfloat16)
        self._build()

    def _attention(self, x, c, h):
        W, nb = self.W, self.nb
        hw, m = h * h, nb * h * h
        g = self._buf(m, c)
        self._gn("mid.attn.norm", x, c, None, 0, hw, False, g)
        q = self._buf(m, c)
        k = self._buf(m, c)
        self._gemm(W.L["mid.attn.q"], [g], [c], 1, 1, m, [(0, 1)], q)
        self._gemm(W.L["mid.attn.k"], [g], [c], 1, 1, m, [(0, 1)], k)
        vt = self._buf(c, hw)
        # scores are produced in query chunks so that the f32 S / bf16 P scratch stays bounded (2048 x hw: 128 MB +
        # 64 MB at a 1024x1024 image instead of 1 GB + 512 MB) and is shared by all images of the batch
        qc = min(hw, max(128, (32 * 1024 * 1024) // hw // 128 * 128))
        s = self._buf(qc, hw, torch.float32)
        p = self._buf(qc, hw)
        o = self._buf(m, c)
        scale = 1.0 / math.sqrt(c)
        for b in range(nb):
            gb, qb, kb, ob = (t[b * hw:(b + 1) * hw] for t in (g, q, k, o))
            # V^T [c, hw] = W_v [c, c] . g_b^T : "activations" = W_v rows, "weights" = g_b rows.  The B operands of
            # these three products (g_b, k_b, V^T) are written by earlier launches of this stream: they are NOT
            # static weights, so the kernel must not fetch them before its grid-dependency wait (_Layer without
            # `static` -> weight_static = 0)
            self._gemm(_Layer(gb, None, hw), [W.wv], [c], 1, 1, c, [(0, 1)], vt, allow_split=False)
            for r0 in range(0, hw, qc):
                rows = min(qc, hw - r0)
                sc, pc = s[:rows], p[:rows]
                self._gemm(_Layer(kb, None, hw), [qb[r0:r0 + rows]], [c], 1, 1, rows, [(0, 1)], sc, allow_split=False)
                self._op(lambda sc=sc, pc=pc, rows=rows: nat.softmax_rows(sc, rows, hw, scale, pc))
                self._gemm(_Layer(vt, W.bv, c), [pc], [hw], 1, 1, rows, [(0, 1)], ob[r0:r0 + rows], allow_split=False)
        out = self._buf(m, c)
        self._gemm(W.L["mid.attn.proj"], [o], [c], 1, 1, m, [(0, 1)], out, residual=x)
        return out

    def _build(self):
        W, nb = self.W, self.nb
        h = self.size
        c = W.L["conv_in"].n
        x = self._buf(nb * h * h, c)
        self._gemm(W.L["conv_in"], [self.x_in], [W.cin_pad], nb, h, h, [(0, 9)], x)
        for i, (chans, has_down) in enumerate(W.blocks):
            for j, (ci, co) in enumerate(chans):
                x = self._resnet(f"down{i}.res{j}", x, ci, None, 0, h)
                c = co
            if has_down:
                x = self._down(W.L[f"down{i}.down"], x, c, h, 0)     # F.pad(0,1,0,1) + stride-2 conv, padding 0
                h //= 2
        x = self._resnet("mid.res0", x, c, None, 0, h)
        x = self._attention(x, c, h)
        x = self._resnet("mid.res1", x, c, None, 0, h)
        a = self._buf(nb * h * h, c)
        self._gn("norm_out", x, c, None, 0, h * h, True, a)
        mc = W.moment_channels
        pre = self._buf(nb * h * h, mc)
        self._gemm(W.L["conv_out"], [a], [c], nb, h, h, [(0, 9)], pre)
        self.moments = self._buf(nb * h * h, mc, torch.float32)
        self._gemm(W.L["quant"], [pre], [mc], 1, 1, nb * h * h, [(0, 1)], self.moments)
        self.latent_size = h


# ================================================================================================
class SegVAEWeights(WeightsBase):
    """GeneralVAESeg encoder + decoder (nn.Sequential indices as in the reference's state-dict)."""

    def __init__(self, vae, device):
        super().__init__(device)
        dec = vae.decoder
        convs = [m for m in dec if isinstance(m, nn.Conv2d)]
        convts = [m for m in dec if isinstance(m, nn.ConvTranspose2d)]
        lns = [m for m in dec if type(m).__name__ == "LayerNorm2d"]
        gns = [m for m in dec if isinstance(m, nn.GroupNorm)]
        assert len(convs) == 2 and len(gns) == 1 and len(convts) == len(lns)
        self.latent_channels = convs[0].in_channels
        self.zpad = 16
        self.dec_c = convs[0].out_channels
        self._gemm("dec.conv_in", pk.pack_conv3x3(_pad_cin(convs[0], self.zpad)), convs[0].bias, convs[0].out_channels)
        self.n_up = len(convts)
        for i, (ct, ln) in enumerate(zip(convts, lns)):
            w, b = pk.pack_convT2x2(ct.weight.detach(), ct.bias.detach())
            self._gemm(f"dec.up{i}", w, b, 4 * ct.out_channels, cout=ct.out_channels)
            self._norm(f"dec.ln{i}", ln)
        self.groups = gns[0].num_groups
        self._norm("dec.gn", gns[0])
        self._gemm("dec.conv_out", pk.pack_conv3x3(convs[1].weight), convs[1].bias, convs[1].out_channels)
        self.num_classes = convs[1].out_channels
        # encoder (optional: skip_encoder variants are not built)
        self.enc = None
        enc = vae.encoder
        if isinstance(enc, nn.Sequential):
            spec = []
            mods = list(enc)
            for idx, m in enumerate(mods):
                if isinstance(m, nn.Conv2d):
                    nxt = mods[idx + 1] if idx + 1 < len(mods) else None
                    silu = isinstance(nxt, nn.SiLU)
                    name = f"enc.{idx}"
                    stride = m.stride[0]
                    cin = m.in_channels
                    cpad = (cin + 15) // 16 * 16 if cin % 8 else cin
                    if stride == 1:
                        self._gemm(name, pk.pack_conv3x3(_pad_cin(m, cpad)), m.bias, m.out_channels)
                    else:
                        self._conv_s2(name, m)
                    spec.append(("conv", name, cin, cpad, m.out_channels, stride, silu))
                elif isinstance(m, nn.GroupNorm):
                    self._norm(f"enc.{idx}", m)
                    spec.append(("gn", f"enc.{idx}", m.num_channels, isinstance(mods[idx + 1], nn.SiLU)))
            self.enc = spec
            self.enc_in = mods[0].in_channels


class SegDecoderPlan(PlanBase):
    """z_in bf16 [nb*L*L, 16] (latents in channels 0..3) -> logits f32 [nb*(4L)^2, classes] channel-last
    (before the final bilinear x2)."""

    def __init__(self, W: SegVAEWeights, nb: int, size: int):
        super().__init__(W, nb)
        self.size = size
        self.z_in = torch.zeros(nb * size * size, W.zpad, device=self.device, dtype=torch.bfloat16)
        h = size
        c = W.dec_c
        x = self._buf(nb * h * h, c)
        self._gemm(W.L["dec.conv_in"], [self.z_in], [W.zpad], nb, h, h, [(0, 9)], x)
        for i in range(W.n_up):
            lay = W.L[f"dec.up{i}"]
            co = lay.extra["cout"]
            t = self._buf(nb * h * h, 4 * co)
            self._gemm(lay, [x], [c], 1, 1, nb * h * h, [(0, 1)], t)
            y = self._buf(nb * 4 * h * h, co)
            g, b, eps = W.norms[f"dec.ln{i}"]
            self._op(lambda t=t, h=h, co=co, g=g, b=b, eps=eps, y=y:
                     nat.convt_shuffle_ln(t, nb, h, h, co, g, b, eps, True, y))
            x, c, h = y, co, 2 * h
        a = self._buf(nb * h * h, c)
        self._gn("dec.gn", x, c, None, 0, h * h, True, a)
        self.logits = self._buf(nb * h * h, W.num_classes, torch.float32)
        self._gemm(W.L["dec.conv_out"], [a], [c], nb, h, h, [(0, 9)], self.logits)
        self.out_size = h


class SegEncoderPlan(PlanBase):
    """bit-plane input bf16 [nb*S*S, cpad] -> moments f32 [nb*(S/8)^2, 2*latent]."""

    def __init__(self, W: SegVAEWeights, nb: int, size: int):
        super().__init__(W, nb)
        if W.enc is None:
            raise NotImplementedError("GeneralVAESeg.encode: only the default nn.Sequential encoder is built")
        first = W.enc[0]
        self.cpad = first[3]
        self.x_in = torch.zeros(nb * size * size, self.cpad, device=self.device, dtype=torch.bfloat16)
        x, c, h = self.x_in, self.cpad, size
        last_conv = [s for s in W.enc if s[0] == "conv"][-1][1]
        for s in W.enc:
            if s[0] == "conv":
                _, name, cin, cpad, cout, stride, silu = s
                act = nat.ACT_SILU if silu else nat.ACT_NONE
                final = name == last_conv
                if stride == 1:
                    y = self._buf(nb * h * h, cout, torch.float32 if final else torch.bfloat16)
                    self._gemm(W.L[name], [x], [c], nb, h, h, [(0, 9)], y, act=act)
                else:
                    y = self._down(W.L[name], x, c, h, 1, act=act)        # nn.Conv2d(stride=2, padding=1)
                    h //= 2
                x, c = y, cout
            else:
                _, name, cn, silu = s
                y = self._buf(nb * h * h, c)
                self._gn(name, x, c, None, 0, h * h, silu, y)
                x = y
        self.moments = x
        self.latent_size = h


# ================================================================================================
class ImageEncoderEngine:
    def __init__(self, vae):
        dev = next(vae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("ldmseg_b200: GeneralVAEImage must live on a CUDA device (no CPU fallback)")
        nat.load()
        self.device = dev
        with torch.cuda.device(dev):
            self.weights = ImageEncoderWeights(vae, dev)
        self.plans: Dict[Tuple[int, int, bool], ImageEncoderPlan] = {}

    def plan(self, nb, size, no_split: bool = False):
        key = (nb, size, no_split)
        if key not in self.plans:
            with torch.cuda.device(self.device):
                self.plans[key] = ImageEncoderPlan(self.weights, nb, size, allow_split=not no_split)
        return self.plans[key]

    @torch.no_grad()
    def encode(self, x: torch.Tensor, in_scale: float = 1.0, in_shift: float = 0.0,
               no_split: bool = False) -> torch.Tensor:
        """x f32 NCHW [B,3,S,S] -> moments f32 NCHW [B,8,S/8,S/8]; (in_scale, in_shift) fuses the
        caller's `2*x-1` (trainers_ldm_cond.py:369) into the layout conversion."""
        nb, c, h, w = x.shape
        if h != w or c != self.weights.in_channels:
            raise RuntimeError(f"GeneralVAEImage.encode expects [B,{self.weights.in_channels},S,S], got {tuple(x.shape)}")
        plan = self.plan(nb, h, no_split)
        with torch.cuda.device(self.device):
            nat.nchw_to_nhwc_bf16(x.float().contiguous(), nb, c, h * w, self.weights.cin_pad, 0, in_scale,
                                  in_shift, plan.x_in)
            plan.run()
            L = plan.latent_size
            mc = self.weights.moment_channels
            out = torch.empty(nb, mc, L, L, device=self.device, dtype=torch.float32)
            nat.nhwc_f32_to_nchw(plan.moments, nb, mc, L * L, mc, 1.0, out)
        return out


class SegVAEEngine:
    def __init__(self, vae):
        dev = next(vae.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("ldmseg_b200: GeneralVAESeg must live on a CUDA device (no CPU fallback)")
        nat.load()
        self.device = dev
        with torch.cuda.device(dev):
            self.weights = SegVAEWeights(vae, dev)
        self.dec_plans: Dict[Tuple[int, int], SegDecoderPlan] = {}
        self.enc_plans: Dict[Tuple[int, int], SegEncoderPlan] = {}

    def dec_plan(self, nb, size):
        if (nb, size) not in self.dec_plans:
            with torch.cuda.device(self.device):
                self.dec_plans[(nb, size)] = SegDecoderPlan(self.weights, nb, size)
        return self.dec_plans[(nb, size)]

    def _run_decoder(self, z: torch.Tensor, scale: float) -> SegDecoderPlan:
        nb, c, h, w = z.shape
        if h != w or c != self.weights.latent_channels:
            raise RuntimeError(f"GeneralVAESeg.decode expects [B,{self.weights.latent_channels},L,L]")
        plan = self.dec_plan(nb, h)
        nat.nchw_to_nhwc_bf16(z.float().contiguous(), nb, c, h * w, self.weights.zpad, 0, scale, 0.0, plan.z_in)
        plan.run()
        return plan

    @torch.no_grad()
    def decode(self, z: torch.Tensor, interpolate: bool = True, scale: float = 1.0) -> torch.Tensor:
        """logits f32 NCHW [B, classes, 8L, 8L] (or 4L without the bilinear x2)."""
        with torch.cuda.device(self.device):
            plan = self._run_decoder(z, scale)
            nb, s, k = plan.nb, plan.out_size, self.weights.num_classes
            if interpolate:
                out = torch.empty(nb, k, 2 * s, 2 * s, device=self.device, dtype=torch.float32)
                nat.bilinear2x_to_nchw(plan.logits, nb, s, s, k, k, out)
            else:
                out = torch.empty(nb, k, s, s, device=self.device, dtype=torch.float32)
                nat.nhwc_f32_to_nchw(plan.logits, nb, k, s * s, k, 1.0, out)
        return out

    @torch.no_grad()
    def decode_ids(self, z: torch.Tensor, scale: float = 1.0):
        """Fused fast path: (argmax ids u8 [B,8L,8L], max softmax prob f32 [B,8L,8L]); the full-resolution
        logits are never written to HBM."""
        with torch.cuda.device(self.device):
            plan = self._run_decoder(z, scale)
            nb, s, k = plan.nb, plan.out_size, self.weights.num_classes
            ids = torch.empty(nb, 2 * s, 2 * s, device=self.device, dtype=torch.uint8)
            prob = torch.empty(nb, 2 * s, 2 * s, device=self.device, dtype=torch.float32)
            nat.bilinear2x_argmax(plan.logits, nb, s, s, k, k, ids, prob)
        return ids, prob

    @torch.no_grad()
    def decode_panoptic(self, z: torch.Tensor, sizes, crops=None, scale: float = 1.0, mask_th: float = 0.5,
                        count_th: int = 512, overlap_th: float = 0.5, ignore_label: int = 0,
                        threshold_output: bool = True):
        """Decode + the whole per-image post-processing of `compute_pq`
        (/root/reference/ldmseg/trainers/trainers_ldm_cond.py:1243-1313) on the device.

        z [B,4,L,L]; sizes = [(h, w)] original image sizes; crops = [(y0, x0, ch, cw)] padding crops on the 8L x 8L
        grid (default: the whole grid).  Returns (ids u8 [B, max_hw] on the device -- image i is
        ids[i, :h*w].view(h, w), 0 = void, id = class + 1 -- and keep int32 [B, 128]: keep[i, c] = 1 iff segment
        id c + 1 of image i survives = the reference's `segments_info`)."""
        with torch.cuda.device(self.device):
            plan = self._run_decoder(z, scale)
            nb, s, k = plan.nb, plan.out_size, self.weights.num_classes
            if len(sizes) != nb:
                raise RuntimeError("decode_panoptic: one (h, w) per image")
            geom = []
            for i, (h, w) in enumerate(sizes):
                y0, x0, ch, cw = crops[i] if crops is not None else (0, 0, 2 * s, 2 * s)
                geom.append([int(h), int(w), int(y0), int(x0), int(ch), int(cw)])
            max_hw = max(g[0] * g[1] for g in geom)
            stride = (max_hw + 15) // 16 * 16
            gd = torch.tensor(geom, dtype=torch.int32).to(self.device, non_blocking=True)
            pred = torch.empty(nb, stride, device=self.device, dtype=torch.int16)
            area = torch.empty(nb, 128, device=self.device, dtype=torch.int32)
            orig = torch.empty(nb, 128, device=self.device, dtype=torch.int32)
            ids = torch.empty(nb, stride, device=self.device, dtype=torch.uint8)
            keep = torch.empty(nb, 128, device=self.device, dtype=torch.int32)
            nat.panoptic_resample(plan.logits, nb, s, k, k, gd, max_hw, stride, mask_th, threshold_output, pred,
                                  area, orig)
            nat.panoptic_filter(pred, nb, gd, max_hw, stride, area, orig, count_th, overlap_th, ignore_label, ids,
                                keep)
        return ids, keep

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> torch.Tensor:
        nb, c, h, w = x.shape
        if (nb, h) not in self.enc_plans:
            with torch.cuda.device(self.device):
                self.enc_plans[(nb, h)] = SegEncoderPlan(self.weights, nb, h)
        plan = self.enc_plans[(nb, h)]
        with torch.cuda.device(self.device):
            nat.nchw_to_nhwc_bf16(x.float().contiguous(), nb, c, h * w, plan.cpad, 0, 1.0, 0.0, plan.x_in)
            plan.run()
            L = plan.latent_size
            mc = plan.moments.shape[1]
            out = torch.empty(nb, mc, L, L, device=self.device, dtype=torch.float32)
            nat.nhwc_f32_to_nchw(plan.moments, nb, mc, L * L, mc, 1.0, out)
        return out
