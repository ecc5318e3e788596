"""Shared plumbing of the launch plans: tiling heuristic, packed-layer records and the PlanBase
helpers that append kernel launches (igemm / GroupNorm / LayerNorm / resnet block) over preallocated
channel-last buffers."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch

from ldmseg import _native as nat
from ldmseg import _pack as pk

SMS = 148


def choose_tiling(m: int, n: int, num_kb: int, sms: int = SMS) -> Tuple[int, int]:
    """Pick (block_n, split_k) for an igemm of M x N with num_kb 64-wide K blocks.

    Cost model (arbitrary units = MMA column-steps): a work item costs
    kb_per_split * max(block_n, 96) + a fixed fill/epilogue overhead; items run in waves of `sms`.
    Split-K is only considered when the tiles alone cannot fill the machine."""
    m_tiles = (m + 127) // 128
    best = (128, 1)
    best_cost = float("inf")
    for bn in (256, 160, 128, 64):
        tiles = m_tiles * ((n + bn - 1) // bn)
        splits = [1]
        if tiles < sms:
            s = 2
            while tiles * s <= sms and num_kb // s >= 4 and s <= 32:
                splits.append(s)
                s += 1
        for s in splits:
            waves = (tiles * s + sms - 1) // sms
            kb = (num_kb + s - 1) // s
            fixed = 24 * bn / 64 + 16 + (10 * bn / 64 if s > 1 else 0)
            cost = waves * (kb * max(bn, 96) / 64.0 + fixed)
            if cost < best_cost - 1e-9:
                best_cost, best = cost, (bn, s)
    return best


class _Layer:
    """Packed parameters of one GEMM-shaped layer."""
    __slots__ = ("w", "bias", "n", "extra")

    def __init__(self, w, bias, n, **extra):
        self.w, self.bias, self.n, self.extra = w, bias, n, extra




class WeightsBase:
    """bf16-packed GEMM weights + fp32 norm parameters, resident on the device."""

    def __init__(self, device):
        self.device = device
        self.L: Dict[str, _Layer] = {}
        self.norms: Dict[str, Tuple[torch.Tensor, torch.Tensor, float]] = {}
        self.groups = 32

    def _dev(self, t, dtype=None):
        return t.detach().to(device=self.device, dtype=dtype or t.dtype).contiguous()

    def _gemm(self, name, packed_f32, bias, n, **extra):
        self.L[name] = _Layer(self._dev(packed_f32, torch.bfloat16),
                              None if bias is None else self._dev(bias.float()), n, **extra)

    def _norm(self, name, mod):
        self.norms[name] = (self._dev(mod.weight.float()), self._dev(mod.bias.float()), float(mod.eps))

    def _resnet(self, name, r, src_split):
        """ResnetBlock2D: conv1, conv2 (+ the 1x1 shortcut as extra K segments, biases summed)."""
        self._norm(name + ".norm1", r.norm1)
        self._norm(name + ".norm2", r.norm2)
        self._gemm(name + ".conv1", pk.pack_conv3x3(r.conv1.weight), r.conv1.bias, r.out_channels)
        w2 = pk.pack_conv3x3(r.conv2.weight)
        b2 = r.conv2.bias.detach().float().clone()
        if r.conv_shortcut is not None:
            w2 = torch.cat([w2, pk.split_linear_k(r.conv_shortcut.weight, src_split)], dim=1)
            b2 = b2 + r.conv_shortcut.bias.detach().float()
        self._gemm(name + ".conv2", w2, b2, r.out_channels)


class PlanBase:
    """Static launch list + buffers.  Subclasses fill `self.ops` in `_build`."""

    def __init__(self, W: WeightsBase, nb: int):
        self.W, self.nb = W, nb
        self.device = W.device
        self.ops: List[Callable[[], None]] = []
        self.n_launch = 0
        self._keep: list = []
        self.ws = torch.zeros(8 * 1024 * 1024, device=self.device, dtype=torch.float32)  # split-K partials
        self.counters = torch.zeros(8192, device=self.device, dtype=torch.int32)
        self.gn_stats = torch.zeros(nb * 64 * 2, device=self.device, dtype=torch.float32)
        self.rowbias_ld = 0

    def _buf(self, rows, c, dtype=torch.bfloat16):
        t = torch.empty(rows, c, device=self.device, dtype=dtype)
        self._keep.append(t)
        return t

    def _op(self, fn, launches=1):
        self.ops.append(fn)
        self.n_launch += launches

    def _gemm(self, layer: _Layer, srcs, src_c, nb, h, w, segs, out, *, rowbias=None, residual=None,
              act=nat.ACT_NONE, out_ld=None, bias="layer", allow_split=True):
        num_kb = sum(taps * ((src_c[s] + 63) // 64) for s, taps in segs)
        bn, split = choose_tiling(nb * h * w, layer.n, num_kb)
        ws_need = nb * h * w * ((layer.n + 3) // 4 * 4)
        if split > 1 and (ws_need > self.ws.numel() or not allow_split):
            split = 1
        p = nat.make_igemm_params(srcs, src_c, nb, h, w, segs, layer.w, layer.n, out,
                                  out_ld if out_ld is not None else out.shape[1],
                                  bias=layer.bias if bias == "layer" else bias,
                                  rowbias=rowbias, rowbias_ld=self.rowbias_ld if rowbias is not None else 0,
                                  residual=residual, res_ld=residual.shape[1] if residual is not None else 0,
                                  act=act, block_n=bn, split_k=split, workspace=self.ws, counters=self.counters)
        self._keep.append(p)
        self._op(lambda p=p: nat.igemm(p))

    def _gn(self, name, src0, c0, src1, c1, hw, silu, out):
        g, b, eps = self.W.norms[name]
        nb, groups, stats = self.nb, self.W.groups, self.gn_stats
        self._op(lambda: nat.groupnorm(src0, c0, src1, c1, nb, hw, groups, g, b, eps, silu, out, stats), 2)

    def _ln(self, name, src, rows, c, out, silu=False):
        g, b, eps = self.W.norms[name]
        self._op(lambda: nat.layernorm(src, rows, c, g, b, eps, silu, out))

    def _resnet(self, name, x, cx, skip, cskip, h, w=None, rowbias=None):
        """ResnetBlock2D over x (+ virtually concatenated skip); returns the output buffer."""
        W, nb = self.W, self.nb
        w = h if w is None else w
        hw, m = h * w, self.nb * h * w
        cin = cx + (cskip if skip is not None else 0)
        l1, l2 = W.L[name + ".conv1"], W.L[name + ".conv2"]
        cout = l1.n
        a1 = self._buf(m, cin)
        self._gn(name + ".norm1", x, cx, skip, cskip if skip is not None else 0, hw, True, a1)
        h1 = self._buf(m, cout)
        # conv1 reads the normalised tensor; for the up blocks that IS the (only) materialised concat
        self._gemm(l1, [a1], [cin], nb, h, w, [(0, 9)], h1, rowbias=rowbias)
        a2 = self._buf(m, cout)
        self._gn(name + ".norm2", h1, cout, None, 0, hw, True, a2)
        out = self._buf(m, cout)
        if cin == cout and skip is None:
            self._gemm(l2, [a2], [cout], nb, h, w, [(0, 9)], out, residual=x)
        elif skip is None:
            self._gemm(l2, [a2, x], [cout, cx], nb, h, w, [(0, 9), (1, 1)], out)
        else:
            self._gemm(l2, [a2, x, skip], [cout, cx, cskip], nb, h, w, [(0, 9), (1, 1), (2, 1)], out)
        return out

    def run(self) -> None:
        for op in self.ops:
            op()
