"""ctypes binding of libldmseg_b200.so (the C ABI declared in include/ldmseg_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, a RuntimeError is
raised.  torch tensors are used only as device-memory handles (data_ptr) and for the current
stream.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# LDMSEG_LIB: development override (A/B-timing two builds inside one GPU session); the default is the in-tree build
_LIB_PATH = os.environ.get("LDMSEG_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libldmseg_b200.so")

MAX_SRC = 3
MAX_SEG = 4
OUT_BF16, OUT_F32 = 0, 1
ACT_NONE, ACT_SILU, ACT_GEGLU = 0, 1, 2

EXPORTS = [
    "ldmseg_version", "ldmseg_last_error_string", "ldmseg_launch_count", "ldmseg_igemm",
    "ldmseg_igemm_simple", "ldmseg_groupnorm", "ldmseg_layernorm", "ldmseg_attention",
    "ldmseg_attention_simple", "ldmseg_geglu", "ldmseg_upsample2x", "ldmseg_im2col_s2",
    "ldmseg_nchw_to_nhwc_bf16", "ldmseg_nhwc_f32_to_nchw", "ldmseg_nhwc_bf16_to_nchw",
    "ldmseg_ddim_step", "ldmseg_sampler_step", "ldmseg_advance_step", "ldmseg_timestep_sinusoid",
    "ldmseg_small_linear", "ldmseg_convt_shuffle_ln", "ldmseg_bilinear2x_to_nchw",
    "ldmseg_bilinear2x_argmax", "ldmseg_select_row", "ldmseg_ddim_step_indexed", "ldmseg_softmax_rows", "ldmseg_nchw_f32_to_nhwc",
    "ldmseg_groupnorm_apply_cs", "ldmseg_set_pdl", "ldmseg_set_debug",
    "ldmseg_groupnorm_apply_cs_f32", "ldmseg_layernorm_f32", "ldmseg_noise_mix", "ldmseg_cross_attention",
    "ldmseg_panoptic_resample", "ldmseg_panoptic_filter", "ldmseg_igemm_max_split_clusters",
]


class IgemmParams(C.Structure):
    _fields_ = [
        ("src", C.c_void_p * MAX_SRC),
        ("src_c", C.c_int * MAX_SRC),
        ("nsrc", C.c_int),
        ("nb", C.c_int), ("h", C.c_int), ("w", C.c_int),
        ("nseg", C.c_int),
        ("seg_src", C.c_int * MAX_SEG),
        ("seg_taps", C.c_int * MAX_SEG),
        ("weight", C.c_void_p),
        ("n", C.c_int),
        ("ktot", C.c_int),
        ("bias", C.c_void_p),
        ("rowbias", C.c_void_p),
        ("rowbias_ld", C.c_int),
        ("residual", C.c_void_p),
        ("res_ld", C.c_int),
        ("out", C.c_void_p),
        ("out_ld", C.c_int),
        ("out_dtype", C.c_int),
        ("act", C.c_int),
        ("block_n", C.c_int),
        ("split_k", C.c_int),
        ("workspace", C.c_void_p),
        ("tile_counters", C.c_void_p),
        ("workspace_elems", C.c_longlong),
        ("stats", C.c_void_p),
        ("stats_hw", C.c_int),
        ("weight_tiled", C.c_int),
        ("pdl", C.c_int),
        ("pair", C.c_int),
        ("weight_static", C.c_int),
        ("next_weight", C.c_void_p),
        ("next_weight_bytes", C.c_longlong),
        ("residual_f32", C.c_int),
        ("out2", C.c_void_p),
        ("out2_ld", C.c_int),
        ("conv_stride", C.c_int),
        ("conv_pad", C.c_int),
        ("rowstats_out", C.c_void_p),
        ("ln_rowstats", C.c_void_p),
        ("ln_colsum", C.c_void_p),
        ("ln_channels", C.c_int),
        ("ln_eps", C.c_float),
        ("stream_k", C.c_int),
        ("split_cluster", C.c_int),
        ("upsample2", C.c_int),
    ]


_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (no GPU needed to load / resolve symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(there is no CPU fallback for the sampling hot path)")
    lib = C.CDLL(_LIB_PATH)
    lib.ldmseg_version.restype = C.c_int
    lib.ldmseg_last_error_string.restype = C.c_char_p
    lib.ldmseg_launch_count.restype = C.c_int64
    vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
    sig = {
        "ldmseg_igemm": [C.POINTER(IgemmParams), vp],
        "ldmseg_igemm_simple": [C.POINTER(IgemmParams), vp],
        "ldmseg_groupnorm": [vp, i32, vp, i32, i32, i32, i32, vp, vp, f32, i32, vp, vp, vp],
        "ldmseg_layernorm": [vp, i32, i32, vp, vp, f32, i32, vp, vp],
        "ldmseg_attention": [vp, i32, i32, i32, i32, vp, vp],
        "ldmseg_attention_simple": [vp, i32, i32, i32, i32, vp, vp],
        "ldmseg_cross_attention": [vp, vp, i32, i32, i32, i32, i32, vp, vp],
        "ldmseg_panoptic_resample": [vp, i32, i32, i32, i32, vp, i32, i32, f32, i32, vp, vp, vp, vp],
        "ldmseg_panoptic_filter": [vp, i32, vp, i32, i32, vp, vp, i32, C.c_double, i32, vp, vp, vp],
        "ldmseg_geglu": [vp, i32, i32, vp, vp],
        "ldmseg_upsample2x": [vp, i32, i32, i32, i32, vp, vp],
        "ldmseg_im2col_s2": [vp, i32, i32, i32, i32, i32, vp, vp],
        "ldmseg_nchw_to_nhwc_bf16": [vp, i32, i32, i32, i32, i32, f32, f32, vp, vp],
        "ldmseg_nhwc_f32_to_nchw": [vp, i32, i32, i32, i32, f32, vp, vp],
        "ldmseg_nhwc_bf16_to_nchw": [vp, i32, i32, i32, i32, f32, vp, vp],
        "ldmseg_nchw_f32_to_nhwc": [vp, i32, i32, i32, i32, f32, vp, vp],
        "ldmseg_ddim_step": [vp, vp, i64, f32, f32, i32, i32, f32, i32, f32, vp, vp, vp, vp],
        "ldmseg_sampler_step": [vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, vp, vp, vp, vp, i32, i32, f32, i32, f32, vp],
        "ldmseg_noise_mix": [vp, vp, vp, vp, i32, i64, f32, i32, vp, vp, i32, i32, vp],
        "ldmseg_groupnorm_apply_cs_f32": [vp, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, f32, i32, vp, vp],
        "ldmseg_layernorm_f32": [vp, i32, i32, vp, vp, f32, i32, vp, vp],
        "ldmseg_advance_step": [vp, vp],
        "ldmseg_timestep_sinusoid": [vp, i32, i32, i32, f32, vp, vp],
        "ldmseg_small_linear": [vp, i32, i32, vp, vp, i32, i32, i32, vp, i32, vp],
        "ldmseg_convt_shuffle_ln": [vp, i32, i32, i32, i32, vp, vp, f32, i32, vp, vp],
        "ldmseg_bilinear2x_to_nchw": [vp, i32, i32, i32, i32, i32, i32, vp, vp],
        "ldmseg_bilinear2x_argmax": [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
        "ldmseg_select_row": [vp, i32, vp, i32, vp, vp],
        "ldmseg_groupnorm_apply_cs": [vp, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, f32, i32, vp, vp],
        "ldmseg_set_pdl": [i32],
        "ldmseg_igemm_max_split_clusters": [i32, i32, i32],
        "ldmseg_set_debug": [i32],
        "ldmseg_softmax_rows": [vp, i32, i32, f32, vp, vp],
        "ldmseg_ddim_step_indexed": [vp, vp, i64, vp, vp, i32, f32, i32, i32, f32, i32, vp, vp, vp],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ldmseg_last_error_string().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "ldmseg_b200: the sampling hot path runs only on CUDA (sm_100a); got a CPU tensor. "
                "There is deliberately no CPU fallback.")


_split_cluster_cap = {}


def max_split_clusters(block_n: int, geglu: bool, cluster_size: int) -> int:
    """Clusters of `cluster_size` split-K CTAs the device holds at once (0 without a CUDA device)."""
    key = (block_n, bool(geglu), cluster_size)
    if key not in _split_cluster_cap:
        n = 0
        if torch.cuda.is_available():
            n = int(load().ldmseg_igemm_max_split_clusters(block_n, 1 if geglu else 0, cluster_size))
        _split_cluster_cap[key] = max(n, 0)
    return _split_cluster_cap[key]


def launch_count() -> int:
    return int(load().ldmseg_launch_count())


# --------------------------------------------------------------------------------------------
def make_igemm_params(srcs: Sequence[torch.Tensor], src_c: Sequence[int], nb: int, h: int, w: int,
                      segs: Sequence[tuple], weight: torch.Tensor, n: int, out: torch.Tensor,
                      out_ld: int, *, bias: Optional[torch.Tensor] = None,
                      rowbias: Optional[torch.Tensor] = None, rowbias_ld: int = 0,
                      residual: Optional[torch.Tensor] = None, res_ld: int = 0, act: int = ACT_NONE,
                      block_n: int = 0, split_k: int = 0, workspace: Optional[torch.Tensor] = None,
                      counters: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None,
                      stats_hw: int = 0, pdl: bool = False, weight_tiled: bool = False,
                      pair: bool = False, weight_static: bool = False, out2: Optional[torch.Tensor] = None,
                      conv_stride: int = 1, conv_pad: int = 1, rowstats_out: Optional[torch.Tensor] = None,
                      ln_rowstats: Optional[torch.Tensor] = None, ln_colsum: Optional[torch.Tensor] = None,
                      ln_channels: int = 0, ln_eps: float = 0.0, stream_k: bool = False,
                      split_cluster: bool = False, upsample2: bool = False) -> IgemmParams:
    p = IgemmParams()
    for i, s in enumerate(srcs):
        p.src[i] = s.data_ptr()
        p.src_c[i] = int(src_c[i])
    p.nsrc = len(srcs)
    p.nb, p.h, p.w = nb, h, w
    p.nseg = len(segs)
    ktot = 0
    for i, (si, taps) in enumerate(segs):
        p.seg_src[i] = si
        p.seg_taps[i] = taps
        ktot += taps * ((src_c[si] + 63) // 64 * 64)
    p.weight = weight.data_ptr()
    p.n = n
    p.ktot = ktot
    assert weight.numel() >= (4 if upsample2 else 1) * (((n + 15) // 16 * 16) if weight_tiled else n) * ktot, \
        (weight.shape, n, ktot)
    p.bias = _ptr(bias)
    p.rowbias = _ptr(rowbias)
    p.rowbias_ld = rowbias_ld
    p.residual = _ptr(residual)
    p.res_ld = res_ld
    p.out = out.data_ptr()
    p.out_ld = out_ld
    p.out_dtype = OUT_F32 if out.dtype == torch.float32 else OUT_BF16
    p.act = act
    p.block_n = block_n
    p.split_k = split_k
    p.workspace = _ptr(workspace)
    p.tile_counters = _ptr(counters)
    p.workspace_elems = workspace.numel() if workspace is not None else 0
    p.stats = _ptr(stats)
    p.stats_hw = stats_hw
    p.weight_tiled = int(weight_tiled)
    p.pdl = int(pdl)
    p.pair = int(pair)
    p.weight_static = int(weight_static)
    p.next_weight = None
    p.next_weight_bytes = 0
    p.residual_f32 = int(residual is not None and residual.dtype == torch.float32)
    p.out2 = _ptr(out2)
    p.out2_ld = out2.shape[1] if out2 is not None else 0
    p.conv_stride = conv_stride
    p.conv_pad = conv_pad
    p.rowstats_out = _ptr(rowstats_out)
    p.ln_rowstats = _ptr(ln_rowstats)
    p.ln_colsum = _ptr(ln_colsum)
    p.ln_channels = ln_channels
    p.ln_eps = ln_eps
    p.stream_k = 1 if stream_k else 0
    p.split_cluster = 1 if split_cluster else 0
    p.upsample2 = 1 if upsample2 else 0
    return p


def igemm(p: IgemmParams, simple: bool = False) -> None:
    lib = load()
    fn = lib.ldmseg_igemm_simple if simple else lib.ldmseg_igemm
    _check(fn(C.byref(p), _stream()), "ldmseg_igemm")


def groupnorm(src0, c0, src1, c1, nb, hw, groups, gamma, beta, eps, silu, out, stats) -> None:
    _check(load().ldmseg_groupnorm(_ptr(src0), c0, _ptr(src1), c1, nb, hw, groups, _ptr(gamma),
                                   _ptr(beta), eps, int(silu), _ptr(out), _ptr(stats), _stream()),
           "ldmseg_groupnorm")


def layernorm(src, rows, c, gamma, beta, eps, silu, out) -> None:
    fn = load().ldmseg_layernorm_f32 if src.dtype == torch.float32 else load().ldmseg_layernorm
    _check(fn(_ptr(src), rows, c, _ptr(gamma), _ptr(beta), eps, int(silu), _ptr(out), _stream()),
           "ldmseg_layernorm")


def attention(qkv, nb, ntok, heads, d, out, simple: bool = False) -> None:
    lib = load()
    fn = lib.ldmseg_attention_simple if simple else lib.ldmseg_attention
    _check(fn(_ptr(qkv), nb, ntok, heads, d, _ptr(out), _stream()), "ldmseg_attention")


def cross_attention(q, kv, nb, ntok_q, ntok_kv, heads, d, out) -> None:
    _check(load().ldmseg_cross_attention(_ptr(q), _ptr(kv), nb, ntok_q, ntok_kv, heads, d, _ptr(out), _stream()),
           "ldmseg_cross_attention")


def geglu(x, rows, c, out) -> None:
    _check(load().ldmseg_geglu(_ptr(x), rows, c, _ptr(out), _stream()), "ldmseg_geglu")


def upsample2x(src, nb, h, w, c, out) -> None:
    _check(load().ldmseg_upsample2x(_ptr(src), nb, h, w, c, _ptr(out), _stream()), "ldmseg_upsample2x")


def im2col_s2(src, nb, h, w, c, pad_lo, out) -> None:
    _check(load().ldmseg_im2col_s2(_ptr(src), nb, h, w, c, pad_lo, _ptr(out), _stream()),
           "ldmseg_im2col_s2")


def nchw_to_nhwc_bf16(src, nb, c, hw, cpad, coff, scale, shift, out) -> None:
    _check(load().ldmseg_nchw_to_nhwc_bf16(_ptr(src), nb, c, hw, cpad, coff, scale, shift, _ptr(out),
                                           _stream()), "ldmseg_nchw_to_nhwc_bf16")


def nhwc_f32_to_nchw(src, nb, c, hw, ld, scale, out) -> None:
    _check(load().ldmseg_nhwc_f32_to_nchw(_ptr(src), nb, c, hw, ld, scale, _ptr(out), _stream()),
           "ldmseg_nhwc_f32_to_nchw")


def nhwc_bf16_to_nchw(src, nb, c, hw, ld, scale, out) -> None:
    _check(load().ldmseg_nhwc_bf16_to_nchw(_ptr(src), nb, c, hw, ld, scale, _ptr(out), _stream()),
           "ldmseg_nhwc_bf16_to_nchw")


def ddim_step(model_out, sample, alpha_t, alpha_prev, ptype, clip, clip_range, use_clipped, prev, x0,
              sigma: float = 0.0, noise=None) -> None:
    _check(load().ldmseg_ddim_step(_ptr(model_out), _ptr(sample), model_out.numel(), alpha_t,
                                   alpha_prev, ptype, int(clip), clip_range, int(use_clipped), sigma,
                                   _ptr(noise), _ptr(prev), _ptr(x0), _stream()), "ldmseg_ddim_step")


def sampler_step(eps, latents, x0, rgb, unet_in, m, coef, step_ptr, nsteps, self_cond, mask=None,
                 known=None, noise=None, sigma=None, ptype: int = 0, clip: bool = False,
                 clip_range: float = 1.0, cfg: bool = False, guidance: float = 1.0) -> None:
    _check(load().ldmseg_sampler_step(_ptr(eps), _ptr(latents), _ptr(x0), _ptr(rgb), _ptr(unet_in), m,
                                      _ptr(coef), _ptr(step_ptr), nsteps, int(self_cond), _ptr(mask),
                                      _ptr(known), _ptr(noise), _ptr(sigma), ptype, int(clip), clip_range,
                                      int(cfg), guidance, _stream()),
           "ldmseg_sampler_step")


def noise_mix(x, noise, timesteps_dev, acp_dev, nb, per_sample, scale, mode, out, unet_in=None, hw=0,
              cpad=0) -> None:
    _check(load().ldmseg_noise_mix(_ptr(x), _ptr(noise), _ptr(timesteps_dev), _ptr(acp_dev), nb, per_sample,
                                   scale, mode, _ptr(out), _ptr(unet_in), hw, cpad, _stream()),
           "ldmseg_noise_mix")


def advance_step(step_ptr) -> None:
    _check(load().ldmseg_advance_step(_ptr(step_ptr), _stream()), "ldmseg_advance_step")


def timestep_sinusoid(t, rows, dim, flip, freq_shift, out) -> None:
    _check(load().ldmseg_timestep_sinusoid(_ptr(t), rows, dim, int(flip), freq_shift, _ptr(out),
                                           _stream()), "ldmseg_timestep_sinusoid")


def small_linear(x, rows, k, w, b, n, silu_in, silu_out, out, out_ld) -> None:
    _check(load().ldmseg_small_linear(_ptr(x), rows, k, _ptr(w), _ptr(b), n, int(silu_in),
                                      int(silu_out), _ptr(out), out_ld, _stream()),
           "ldmseg_small_linear")


def convt_shuffle_ln(src, nb, h, w, c, gamma, beta, eps, silu, out) -> None:
    _check(load().ldmseg_convt_shuffle_ln(_ptr(src), nb, h, w, c, _ptr(gamma), _ptr(beta), eps,
                                          int(silu), _ptr(out), _stream()), "ldmseg_convt_shuffle_ln")


def bilinear2x_to_nchw(src, nb, h, w, c, ld, out) -> None:
    _check(load().ldmseg_bilinear2x_to_nchw(_ptr(src), int(src.dtype == torch.float32), nb, h, w, c,
                                            ld, _ptr(out), _stream()), "ldmseg_bilinear2x_to_nchw")


def bilinear2x_argmax(src, nb, h, w, c, ld, ids, maxprob=None) -> None:
    _check(load().ldmseg_bilinear2x_argmax(_ptr(src), int(src.dtype == torch.float32), nb, h, w, c,
                                           ld, _ptr(ids), _ptr(maxprob), _stream()),
           "ldmseg_bilinear2x_argmax")


def panoptic_resample(logits, nb, s, c, ld, geom, max_hw, out_stride, mask_th, threshold_output, pred, area,
                      orig_area) -> None:
    _check(load().ldmseg_panoptic_resample(_ptr(logits), nb, s, c, ld, _ptr(geom), max_hw, out_stride, mask_th,
                                           int(threshold_output), _ptr(pred), _ptr(area), _ptr(orig_area),
                                           _stream()), "ldmseg_panoptic_resample")


def panoptic_filter(pred, nb, geom, max_hw, out_stride, area, orig_area, count_th, overlap_th, ignore_label, ids,
                    keep) -> None:
    _check(load().ldmseg_panoptic_filter(_ptr(pred), nb, _ptr(geom), max_hw, out_stride, _ptr(area),
                                         _ptr(orig_area), count_th, overlap_th, ignore_label, _ptr(ids),
                                         _ptr(keep), _stream()), "ldmseg_panoptic_filter")


def select_row(table, ncols, step_ptr, nb, dst) -> None:
    _check(load().ldmseg_select_row(_ptr(table), ncols, _ptr(step_ptr), nb, _ptr(dst), _stream()),
           "ldmseg_select_row")


def ddim_step_indexed(model_out, sample, timestep_dev, acp_dev, step_ratio, final_alpha, ptype, clip,
                      clip_range, use_clipped, prev, x0) -> None:
    _check(load().ldmseg_ddim_step_indexed(_ptr(model_out), _ptr(sample), model_out.numel(),
                                           _ptr(timestep_dev), _ptr(acp_dev), step_ratio, final_alpha,
                                           ptype, int(clip), clip_range, int(use_clipped), _ptr(prev),
                                           _ptr(x0), _stream()), "ldmseg_ddim_step_indexed")


def softmax_rows(s, rows, cols, scale, out) -> None:
    _check(load().ldmseg_softmax_rows(_ptr(s), rows, cols, scale, _ptr(out), _stream()),
           "ldmseg_softmax_rows")


def nchw_f32_to_nhwc(src, nb, c, hw, ld, scale, out) -> None:
    _check(load().ldmseg_nchw_f32_to_nhwc(_ptr(src), nb, c, hw, ld, scale, _ptr(out), _stream()),
           "ldmseg_nchw_f32_to_nhwc")


def groupnorm_apply_cs(src0, c0, cs0, src1, c1, cs1, nb, hw, groups, gamma, beta, eps, silu, out) -> None:
    f32 = src0.dtype == torch.float32
    assert src1 is None or (src1.dtype == torch.float32) == f32, "GroupNorm sources must share a dtype"
    fn = load().ldmseg_groupnorm_apply_cs_f32 if f32 else load().ldmseg_groupnorm_apply_cs
    _check(fn(_ptr(src0), c0, _ptr(cs0), _ptr(src1), c1, _ptr(cs1), nb, hw, groups, _ptr(gamma), _ptr(beta), eps,
              int(silu), _ptr(out), _stream()), "ldmseg_groupnorm_apply_cs")


def set_pdl(enable: bool) -> bool:
    """Enable programmatic dependent launch for every kernel of the library; returns the old setting."""
    return bool(load().ldmseg_set_pdl(int(enable)))
