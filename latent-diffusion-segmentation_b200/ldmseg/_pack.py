"""Weight packing for the tcgen05 implicit-GEMM kernel.

The kernel consumes bf16 weights [N, Ktot], K-contiguous, K ordered [segment][tap (ky,kx)][channel]
with every (segment, tap) slice zero-padded to a multiple of 64 channels (one 128-byte swizzle row
of the TMA box).  These helpers turn PyTorch-layout parameters into that layout once, at load time.
"""
from __future__ import annotations

from typing import Sequence

import torch


TILE_ROWS = 16


def _pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def pack_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """[N, C, 3, 3] -> [N, 9 * pad64(C)] (tap-major, then channel)."""
    n, c, kh, kw = w.shape
    assert kh == 3 and kw == 3
    cp = _pad64(c)
    out = torch.zeros(n, 9, cp, dtype=torch.float32, device=w.device)
    out[:, :, :c] = w.float().permute(0, 2, 3, 1).reshape(n, 9, c)
    return out.reshape(n, 9 * cp)


def pack_upsample2_conv3x3(w: torch.Tensor) -> torch.Tensor:
    """Nearest x2 up-sampling folded into the 3x3 convolution that follows it (diffusers Upsample2D):
    [N, C, 3, 3] -> [4 * pad16(N), 4 * pad64(C)], the four phase matrices stacked along N (phase = 2*py + px).
    Output pixel (2y+py, 2x+px) reads the up-sampled rows 2y+py+ky-1, i.e. the input rows y-1, y, y for py = 0 and
    y, y, y+1 for py = 1 (same along x), so it is a 2x2 convolution of the INPUT whose tap (a, b) reads
    (y+py+a-1, x+px+b-1) with the sum of the 3x3 taps that land there (ldmseg_igemm_params.upsample2)."""
    n, c, kh, kw = w.shape
    assert kh == 3 and kw == 3
    cp, npad = _pad64(c), (n + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS
    wf = w.detach().float()
    taps = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}      # phase bit -> 3x3 taps summed into 2x2 tap 0 / 1
    out = torch.zeros(4, npad, 4, cp, dtype=torch.float32, device=w.device)
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    acc = torch.zeros(n, c, dtype=torch.float32, device=w.device)
                    for ky in taps[py][a]:
                        for kx in taps[px][b]:
                            acc += wf[:, :, ky, kx]
                    out[2 * py + px, :n, 2 * a + b, :c] = acc
    return out.reshape(4 * npad, 4 * cp)


def pack_conv3x3_im2col(w: torch.Tensor) -> torch.Tensor:
    """[N, C, 3, 3] -> [N, pad64(9*C)] for convs fed by the im2col kernel (taps contiguous, no per-tap
    padding; identical to pack_conv3x3 when C is a multiple of 64)."""
    n, c, kh, kw = w.shape
    assert kh == 3 and kw == 3
    return pack_linear(w.float().permute(0, 2, 3, 1).reshape(n, 9 * c))


def pack_linear(w: torch.Tensor) -> torch.Tensor:
    """[N, K] (or [N, K, 1, 1]) -> [N, pad64(K)]."""
    if w.dim() == 4:
        w = w[:, :, 0, 0]
    n, k = w.shape
    kp = _pad64(k)
    out = torch.zeros(n, kp, dtype=torch.float32, device=w.device)
    out[:, :k] = w.float()
    return out


def split_linear_k(w: torch.Tensor, splits: Sequence[int]) -> torch.Tensor:
    """Linear / 1x1 weight whose K axis is a concat of sources: pad each source slice to 64."""
    if w.dim() == 4:
        w = w[:, :, 0, 0]
    parts, off = [], 0
    for c in splits:
        parts.append(pack_linear(w[:, off:off + c]))
        off += c
    assert off == w.shape[1]
    return torch.cat(parts, dim=1)


def split_conv3x3_k(w: torch.Tensor, splits: Sequence[int]) -> torch.Tensor:
    """conv3x3 weight over a channel concat: one 9-tap segment per source."""
    parts, off = [], 0
    for c in splits:
        parts.append(pack_conv3x3(w[:, off:off + c]))
        off += c
    assert off == w.shape[1]
    return torch.cat(parts, dim=1)


def interleave_geglu(w: torch.Tensor, b: torch.Tensor):
    """GEGLU projection [2*F, K] (h rows then g rows) -> rows interleaved in blocks of 16 so that
    every 32-column chunk of the GEMM output holds [16 h | 16 g] for the same 16 outputs."""
    f = w.shape[0] // 2
    assert f % 16 == 0
    wh, wg = w[:f].reshape(f // 16, 16, -1), w[f:].reshape(f // 16, 16, -1)
    wi = torch.cat([wh, wg], dim=1).reshape(2 * f, -1)
    bh, bg = b[:f].reshape(f // 16, 16), b[f:].reshape(f // 16, 16)
    bi = torch.cat([bh, bg], dim=1).reshape(2 * f)
    return wi, bi


def pack_convT2x2(w: torch.Tensor, b: torch.Tensor):
    """ConvTranspose2d(k=2, s=2) weight [Cin, Cout, 2, 2] -> GEMM weight [4*Cout, pad64(Cin)] with
    row block t = ky*2+kx, and the bias repeated per tap."""
    cin, cout, kh, kw = w.shape
    assert kh == 2 and kw == 2
    wt = w.float().permute(2, 3, 1, 0).reshape(4 * cout, cin)
    return pack_linear(wt), b.float().repeat(4)


def fold_layernorm(w: torch.Tensor, b, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm folded into the Linear that consumes it:  LN(x) W^T + b  =  rstd (x W'^T - mean * colsum) + c  with
    W' = W diag(gamma), c = W beta + b and colsum[n] = sum_k W'[n, k] taken over the bf16-ROUNDED W' (so that the mean
    term cancels exactly what the tensor cores accumulate).  Returns (W' f32, c f32, colsum f32)."""
    w = w.detach().float()
    wf = w * gamma.detach().float()[None, :]
    c = w @ beta.detach().float()
    if b is not None:
        c = c + b.detach().float()
    colsum = wf.to(torch.bfloat16).float().sum(dim=1)
    return wf, c, colsum


def to_bf16(w: torch.Tensor) -> torch.Tensor:
    return w.to(torch.bfloat16).contiguous()


def tile_pack(w2d: torch.Tensor) -> torch.Tensor:
    """Row-major packed weight [N, Ktot] (Ktot % 64 == 0) -> block-tiled [ceil(N/16), Ktot/64, 16, 64]:
    every 16-row x 64-k block is 2 KB contiguous, so the weight stream of a tile reads whole DRAM pages
    (ldmseg_igemm_params.weight_tiled).  16 rows: a CTA of a pair stages HALF a B tile, 80 rows at block_n 160."""
    n, k = w2d.shape
    assert k % 64 == 0
    npad = (n + TILE_ROWS - 1) // TILE_ROWS * TILE_ROWS
    if npad != n:
        w2d = torch.cat([w2d, torch.zeros(npad - n, k, dtype=w2d.dtype, device=w2d.device)], dim=0)
    return w2d.reshape(npad // TILE_ROWS, TILE_ROWS, k // 64, 64).permute(0, 2, 1, 3).contiguous()
