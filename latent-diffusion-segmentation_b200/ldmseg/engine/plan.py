"""Shared plumbing of the launch plans: tiling heuristic, packed-layer records and the PlanBase
helpers that append kernel launches (igemm / GroupNorm / LayerNorm / resnet block) over preallocated
channel-last buffers."""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Optional, Tuple

import torch

from ldmseg import _native as nat
from ldmseg import _pack as pk

SMS = 148
# measured on B200 (tools/bench_igemm.py): operand ingest of one SM sustains ~59 B/cycle, i.e. one
# 128-byte swizzle row of a TMA box every ~2.17 cycles; an M=128 x N=bn x K=16 tcgen05.mma takes bn/2 cycles
CYC_PER_TMA_ROW = 2.17
USE_PDL = os.environ.get("LDMSEG_PDL", "1") != "0"
FUSE_GN_STATS = os.environ.get("LDMSEG_FUSE_GN_STATS", "1") != "0"
USE_PAIR = os.environ.get("LDMSEG_PAIR", "1") != "0"     # CTA pairs (tcgen05 cta_group::2) where the model prefers them
# stride-2 convolutions read their input through the TMA traversal stride (no im2col buffer); 0 = im2col + GEMM
USE_S2_TMA = os.environ.get("LDMSEG_S2_TMA", "1") != "0"
# fp32 residual / skip stream (SURVEY.md §7 hard part 3): every tensor that is added back later (resnet outputs,
# transformer hidden states) is kept in f32, with a bf16 shadow only where a later launch reads it through TMA
RESID_F32 = os.environ.get("LDMSEG_RESID_F32", "0") != "0"
# each igemm launch pulls the NEXT launch's weights into L2 while its own tail runs (weight streaming at small
# batch); only below this many output rows per forward level-0 launch (large batches are compute-bound)
NEXTW_MAX_ROWS = int(os.environ.get("LDMSEG_NEXTW_MAX_ROWS", "8192"))
# LayerNorm folded into the GEMMs around it (row moments from the producer's epilogue, gamma / beta in the consumer's
# weights): one launch per LayerNorm less (32 per UNet forward); 0 = separate LayerNorm kernels
LN_FOLD = os.environ.get("LDMSEG_LN_FOLD", "1") != "0"
# stream-K tail: the tiles past the last whole wave of a persistent grid are cut along K into one piece per CTA
# (csrc/igemm.cu, TailSeg); 0 = ragged last waves run as whole tiles
STREAM_K = os.environ.get("LDMSEG_STREAM_K", "1") != "0"
# split-K exchange through distributed shared memory (the splits of a tile launched as one thread-block cluster, every
# unit of the partial tile pushed into its owner's shared memory) wherever the device can hold all the tiles' clusters
# at once.  OFF by default: measured in-graph at batch 1 it is no faster than the global workspace (2793 vs 2756 us per
# UNet forward; profiles/r02_ab_cluster_splitk.log) -- moving a 64-128 KB fp32 partial tile per CTA over the SM-to-SM
# network takes ~3 us, as long as the three L2 round trips it replaces; 1 = use it
SPLIT_CLUSTER = os.environ.get("LDMSEG_SPLIT_CLUSTER", "0") != "0"
# nearest x2 up-sampling folded into the 3x3 convolution after it (four 2x2 phase GEMMs over the input pixels, 4/9 of
# the multiply-adds, no up-sampled tensor; ldmseg_igemm_params.upsample2); 0 = upsample2x kernel + plain 3x3 conv
UP2_FOLD = os.environ.get("LDMSEG_UP2_FOLD", "1") != "0"


# 320-wide pair tiles (csrc/igemm.cu, IgemmCfg: two N = 160 tcgen05.mma per k-step over three accumulator slots) for
# the multi-wave N = 320 / 640 / 1280 convolutions of batches >= 4; 0 = at most 256-wide tiles
USE_BN320 = os.environ.get("LDMSEG_BN320", "1") != "0"
USE_TUNED = os.environ.get("LDMSEG_TUNED", "1") != "0"
_TUNED: Optional[Dict[str, list]] = None


def _tuned_table() -> Dict[str, list]:
    """Measured (block_n, split_k, pair) per (M, N, k-blocks) from tools/tune_tiling.py; empty if absent."""
    global _TUNED
    if _TUNED is None:
        path = os.environ.get("LDMSEG_TUNED_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                                   "tuned_b200.json")
        try:
            with open(path) as f:
                _TUNED = json.load(f).get("entries", {})
        except (OSError, ValueError):
            _TUNED = {}
    return _TUNED


# measured on B200 (profiles/r02_ablate_unet_b8_streamk_all.log against r02_ablate_unet_b8_c.log): publishing a partial
# tile, waiting for the peers and adding their partials costs the last wave ~5.8 us whatever the tile's K: launches
# with >= 90 k-blocks gained 15-30 %, launches with <= 45 k-blocks lost 5-20 %.  Per 160 tile columns (the partial tiles
# cross L2 twice).
TAIL_EXCHANGE_CYC = 11000.0


def stream_k_shape_ok(unit_tiles: int, num_kb: int, units: int) -> bool:
    """Shapes the stream-K tail is defined / sensible for: a ragged last wave whose K steps give every unit a
    non-trivial piece, and -- without a whole wave before it -- pieces of at least half a tile (below that uniform
    split-K, whose splits share the final reduction, is the better cut)."""
    whole, rem = divmod(unit_tiles, units)
    if rem == 0 or rem * num_kb < 4 * units:
        return False
    return whole > 0 or 2 * rem >= units


# cycles of the split-K exchange per launch: through the global workspace (partial tile to L2, release atomic, acquire
# spin, partials back, counter re-arm; measured in-graph at batch 1 as 7.6 us per launch) and inside a cluster
SPLIT_EXCHANGE_CYC = 9500.0
CSPLIT_EXCHANGE_CYC = 3500.0


def choose_tiling_ex(m: int, n: int, num_kb: int, sms: int = SMS, allow_split: bool = True,
                     allow_pair: bool = False, use_tuned: bool = True, allow_tail: bool = False,
                     cluster_cap: Optional[Callable[[int, int], int]] = None,
                     allow_320: bool = False) -> Tuple[int, int, bool, bool, bool]:
    """Pick (block_n, split_k, pair, stream_k_tail, split_cluster) for an igemm of M x N with num_kb 64-wide K blocks.
    `cluster_cap(block_n, split)`: how many clusters of `split` CTAs the device holds at once (None: no clusters).

    Cycle model per work item (one CTA, one 128 x bn output tile, one K split):
        k-blocks * max(MMA = 4 * bn/2, ingest = 2.17 * (128 + bn))  +  prologue  +  epilogue
    items run in waves of `sms` CTAs.  A CTA pair (cta_group::2, 256 x bn tile) stages only bn/2 rows of B per CTA
    (a measured per-k-block saving, against a fixed cluster-launch cost); it needs an even number of 128-row tiles.
    Split-K (partials to a workspace, cooperative reduce) is considered only when the tiles cannot fill the machine,
    and is charged for the partial store and the reduction.  The stream-K tail (csrc/igemm.cu, TailSeg) replaces the
    ragged last wave by its share of a wave plus a fixed exchange cost."""
    if use_tuned and USE_TUNED and allow_split and sms == SMS:
        hit = _tuned_table().get(f"{m},{n},{num_kb}")
        if hit is not None and (allow_pair or not hit[2]):
            bn, s, pair = int(hit[0]), int(hit[1]), bool(hit[2])
            tiles = ((m + 127) // 128) * ((n + bn - 1) // bn)
            clus = bool(s > 1 and not pair and cluster_cap is not None and tiles <= cluster_cap(bn, s))
            return bn, s, pair, False, clus
    m_tiles = (m + 127) // 128
    best, best_cost = (128, 1, False, False, False), float("inf")
    cands = [(bn, False) for bn in (256, 160, 128, 64)]
    # (not for short K: there the epilogue bounds the kernel and coupling two CTAs' accumulator hand-over costs
    # 10-14 %, measured at K = 320)
    if allow_pair and m_tiles >= 2 and m_tiles % 2 == 0 and num_kb >= 10:
        # 320-wide pair tiles (two N = 160 MMAs per k-step sharing the staged A rows; whole tiles or stream-K tail
        # only, not GEGLU): a third less shared-memory traffic per multiply-add than 160-wide tiles
        # -- from ~32 k-blocks on: per layer in the batch-8 forward (profiles/r02_ablate_unet_b8_i.log against
        # r02_ablate_unet_b8_h.log) every convolution and the K = 2560 feed-forward output gained 6-44 us, the K = 640 /
        # 1280 projections (10 / 20 k-blocks, epilogue-bound) lost 0.6-2.4 us each
        wide = allow_320 and n >= 320 and num_kb >= 32
        cands = [(bn, True) for bn in ((320, 256, 160, 128) if wide else (256, 160, 128))] + cands
    for bn, pair in cands:
        tiles = m_tiles * ((n + bn - 1) // bn)
        t_kb = max(2.0 * bn, CYC_PER_TMA_ROW * (128 + bn))
        if pair:
            # measured (tools/bench_ingest.py, in-graph A/B): halving the B staging saves ~45 cycles per k-block at
            # bn = 160 (the MMA rate of narrow tiles, not ingest, is the larger term), and a cluster launch costs
            # ~1.5 us more -- pairs win once a CTA runs a few hundred k-blocks (batch 8: -3 % on the forward) and
            # lose in the single-wave launches of batch 1
            t_kb -= 45.0 * bn / 160.0
        if bn != 64 and ((m_tiles + 1) // 2 if pair else m_tiles) * ((n + bn - 1) // bn) > (sms // 2 if pair else sms):
            # multi-wave launches, measured per k-block at batch 8 under the power cap (pairs:
            # profiles/r02_ablate_unet_b8_streamk*.log; single CTAs from the pair / single ratios of
            # profiles/r01_bench_ingest_v11.log): the fit above is 7 % low at bn 160 and 15 % high at bn 256
            # (320: tools/bench_bn320.py, profiles/r02_bench_bn320.log -- 834 cycles against 2 x 604 for the same
            # columns as two 160-wide tiles)
            t_kb = ({128: 545.0, 160: 620.0, 256: 665.0, 320: 834.0}[bn] if pair
                    else {128: 580.0, 160: 660.0, 256: 719.0}[bn])
        chunks = bn / 32.0
        splits = [1]
        if allow_split and tiles < sms and bn != 320:
            s = 2
            while tiles * s <= sms and num_kb // s >= 4 and s <= 16:
                splits.append(s)
                s += 1
        units = sms // 2 if pair else sms
        unit_tiles = ((m_tiles + 1) // 2 if pair else m_tiles) * ((n + bn - 1) // bn)
        for s in splits:
            waves = (tiles * s + sms - 1) // sms
            kb = (num_kb + s - 1) // s
            # measured (tools/bench_split.py, bench_epi.py): ~5 us of launch + prologue + pipeline fill + epilogue
            # tail per wave, ~5 us more for the split-K publish / wait-for-peers / reduce pass
            item = kb * t_kb + 9500 + chunks * 120
            clus = False
            if s > 1:
                clus = bool(not pair and cluster_cap is not None and tiles <= cluster_cap(bn, s))
                item += (CSPLIT_EXCHANGE_CYC if clus else SPLIT_EXCHANGE_CYC) + chunks * 150
            cost = waves * item + (3000 if pair else 0)
            if cost < best_cost - 1e-9:
                best_cost, best = cost, (bn, s, pair, False, clus)
            if s == 1 and allow_tail and bn != 64 and stream_k_shape_ok(unit_tiles, num_kb, units):
                # tail against whole tiles for THIS tiling, with the per-launch overhead counted once (a persistent
                # CTA pays prologue and pipeline fill once, not per wave); taken when it is worth at least 3 %, and
                # ranked against the other tilings by scaling this tiling's whole-tile cost
                whole, rem = divmod(unit_tiles, units)
                plain = waves * (kb * t_kb + chunks * 120) + 9500
                tail = (whole * (kb * t_kb + chunks * 120) + rem / units * kb * t_kb + 9500
                        + TAIL_EXCHANGE_CYC * bn / 160.0 + chunks * 150 * (1 + min(4.0, units / rem)))
                if tail < 0.97 * plain and cost * tail / plain < best_cost - 1e-9:
                    best_cost, best = cost * tail / plain, (bn, 1, pair, True, False)
    return best


def choose_tiling(m: int, n: int, num_kb: int, sms: int = SMS, allow_split: bool = True,
                  allow_pair: bool = False, use_tuned: bool = True) -> Tuple[int, int, bool]:
    """(block_n, split_k, pair) without the stream-K tail (tools/tune_tiling.py, tests)."""
    return choose_tiling_ex(m, n, num_kb, sms, allow_split, allow_pair, use_tuned, False)[:3]


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
        self._shared: Dict[str, torch.Tensor] = {}

    def shared(self, name: str, numel: int, dtype) -> torch.Tensor:
        """Scratch shared by every plan of this engine (plans run one at a time on one stream): split-K
        workspace and tile counters.  Grows on demand."""
        t = self._shared.get(name)
        if t is None or t.numel() < numel:
            t = torch.zeros(numel, device=self.device, dtype=dtype)
            self._shared[name] = t
        return t

    def _dev(self, t, dtype=None):
        return t.detach().to(device=self.device, dtype=dtype or t.dtype).contiguous()

    def _gemm(self, name, packed_f32, bias, n, **extra):
        extra.setdefault("ktot", packed_f32.shape[1])
        extra.setdefault("tiled", True)
        extra.setdefault("name", name)
        extra.setdefault("static", True)      # a real parameter: nothing on the stream writes it
        self.L[name] = _Layer(self._dev(pk.tile_pack(packed_f32), torch.bfloat16),
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
        has_sc = r.conv_shortcut is not None
        if has_sc:
            w2 = torch.cat([w2, pk.split_linear_k(r.conv_shortcut.weight, src_split)], dim=1)
            b2 = b2 + r.conv_shortcut.bias.detach().float()
        self._gemm(name + ".conv2", w2, b2, r.out_channels, shortcut=has_sc)

    def _conv_s2(self, name, conv):
        """3x3 stride-2 convolution: per-tap packing for the TMA-strided kernel, im2col packing otherwise."""
        w = conv.weight
        if USE_S2_TMA:
            c = w.shape[1]
            if c % 8:
                cp = (c + 15) // 16 * 16
                wp = torch.zeros(w.shape[0], cp, 3, 3, device=w.device)
                wp[:, :c] = w.detach().float()
                w = wp
            self._gemm(name, pk.pack_conv3x3(w), conv.bias, conv.out_channels, s2_tma=True)
        else:
            self._gemm(name, pk.pack_conv3x3_im2col(w), conv.bias, conv.out_channels, s2_tma=False)


class PlanBase:
    """Static launch list + buffers.  Subclasses fill `self.ops` in `_build`.

    GroupNorm statistics are fused into the producing igemm whenever the GroupNorm input was written by
    an igemm of this plan: `_gn` then assigns that producer a slice of `stats_arena` (per-image,
    per-channel sum / sum of squares, zeroed once per run) and emits a single apply kernel."""

    def __init__(self, W: WeightsBase, nb: int, allow_split: bool = True):
        self.W, self.nb = W, nb
        self.allow_split = allow_split   # False: no launch of this plan waits for sibling CTAs (side-stream use)
        self.device = W.device
        self.ops: List[Callable[[], None]] = []
        self.tags: List[str] = []            # one per op: "<family>:<rows>:<layer>" (profiling / ablation only)
        self.n_launch = 0
        self._keep: list = []
        self.ws = W.shared("splitk_ws", 16 * 1024 * 1024, torch.float32)   # split-K partials (self-cleaning)
        self.counters = W.shared("splitk_counters", 8192, torch.int32)
        self.resid_f32 = RESID_F32
        self._f32: Dict[int, torch.Tensor] = {}     # bf16 handle ptr -> f32 tensor of the residual stream
        self._igemm_params: list = []
        self.gn_stats = torch.zeros(nb * 64 * 2, device=self.device, dtype=torch.float32)
        # per-(image, channel) GroupNorm moments and per-row LayerNorm moments written by producer epilogues; one
        # memset per run
        self.stats_arena = torch.zeros(nb * 1024 * 1024, device=self.device, dtype=torch.float32)
        self._arena_used = 0
        self._producer: Dict[int, tuple] = {}   # out.data_ptr() -> (IgemmParams, n, rows_per_image)
        self._chan_stats: Dict[int, torch.Tensor] = {}
        self.rowbias_ld = 0
        self.pdl = USE_PDL
        # only the part of the arena the finished plan uses is cleared (known when the first run happens)
        self._op(lambda: self.stats_arena[: max(self._arena_used, 1)].zero_(), 1, tag="memset:0:stats_arena")

    def _buf(self, rows, c, dtype=torch.bfloat16):
        t = torch.empty(rows, c, device=self.device, dtype=dtype)
        self._keep.append(t)
        return t

    def _op(self, fn, launches=1, tag="misc:0:"):
        self.ops.append(fn)
        self.tags.append(tag)
        self.n_launch += launches

    def _gemm(self, layer: _Layer, srcs, src_c, nb, h, w, segs, out, *, rowbias=None, residual=None,
              act=nat.ACT_NONE, out_ld=None, bias="layer", allow_split=True, stream=False, shadow=True,
              conv_stride=1, conv_pad=1, rowstats=None, ln=None, upsample2=False):
        """Append one igemm launch.  `stream=True` marks `out` as a tensor of the residual stream: under
        RESID_F32 it is written as f32 (+ a bf16 shadow in `out` when `shadow`), and later `residual=` /
        GroupNorm / LayerNorm reads of `out` use the f32 copy."""
        num_kb = sum(taps * ((src_c[s] + 63) // 64) for s, taps in segs)
        m = nb * h * w
        tiled = bool(layer.extra.get("tiled", False))
        allow_split = allow_split and self.allow_split
        allow_pair = USE_PAIR and tiled
        if upsample2:
            # (nb, h, w) is the INPUT geometry; the kernel runs 4 phases x ceil(m / 128) tiles of 128 input pixels and
            # a CTA pair must stay inside one phase
            assert layer.extra.get("up2", False) and tiled and segs == [(0, 4)], layer.extra.get("name")
            allow_pair = allow_pair and ((m + 127) // 128) % 2 == 0
            m = 4 * ((m + 127) // 128) * 128
        # the stream-K tail shares the split-K premise (a co-resident grid) and its scratch buffers
        tail_ok = STREAM_K and allow_split and tiled and act != nat.ACT_GEGLU
        geglu = act == nat.ACT_GEGLU
        cap = (lambda bn_, s_: nat.max_split_clusters(bn_, geglu, s_)) if SPLIT_CLUSTER and allow_split else None
        allow_320 = USE_BN320 and not geglu
        bn, split, pair, stream_k, split_cluster = choose_tiling_ex(
            m, layer.n, num_kb, allow_split=allow_split, allow_pair=allow_pair, allow_tail=tail_ok,
            cluster_cap=cap, allow_320=allow_320)
        tiles = ((m + 127) // 128) * ((layer.n + bn - 1) // bn)
        if split > 1 and not split_cluster and tiles * split * 128 * bn > self.ws.numel():
            bn, split, pair, stream_k, split_cluster = choose_tiling_ex(
                m, layer.n, num_kb, allow_split=False, allow_pair=allow_pair, allow_tail=False, allow_320=allow_320)
        out_main, out2 = out, None
        if stream and self.resid_f32 and out.dtype == torch.bfloat16:
            f = self._buf(out.shape[0], out.shape[1], torch.float32)
            self._f32[out.data_ptr()] = f
            out_main, out2 = f, (out if shadow else None)
        if residual is not None:
            residual = self._f32.get(residual.data_ptr(), residual)
        p = nat.make_igemm_params(srcs, src_c, nb, h, w, segs, layer.w, layer.n, out_main,
                                  out_ld if out_ld is not None else out_main.shape[1],
                                  bias=layer.bias if bias == "layer" else bias,
                                  rowbias=rowbias, rowbias_ld=self.rowbias_ld if rowbias is not None else 0,
                                  residual=residual, res_ld=residual.shape[1] if residual is not None else 0,
                                  act=act, block_n=bn, split_k=split, workspace=self.ws, counters=self.counters,
                                  pdl=self.pdl, weight_tiled=tiled, pair=pair,
                                  weight_static=bool(layer.extra.get("static", False)), out2=out2,
                                  conv_stride=conv_stride, conv_pad=conv_pad, rowstats_out=rowstats,
                                  ln_rowstats=None if ln is None else ln[0],
                                  ln_colsum=None if ln is None else layer.extra["ln_colsum"],
                                  ln_channels=0 if ln is None else ln[1], ln_eps=0.0 if ln is None else ln[2],
                                  stream_k=stream_k, split_cluster=split_cluster, upsample2=upsample2)
        self._keep.append(p)
        self._igemm_params.append((p, layer, m))
        if out.dtype == torch.bfloat16 and act != nat.ACT_GEGLU and out.is_contiguous():
            self._producer[out.data_ptr()] = (p, layer.n, out.shape[0])
        self._op(lambda p=p: nat.igemm(p), tag=f"igemm:{out.shape[0] if upsample2 else m}:{layer.extra.get('name', '')}:n{layer.n}:kb{num_kb}:bn{bn}:s{split}:p{int(pair)}:t{int(stream_k)}:c{int(split_cluster)}")

    def _down(self, layer: _Layer, x, c, h, pad_lo, act=nat.ACT_NONE):
        """3x3 stride-2 convolution of x [nb*h*h, c] -> [nb*(h/2)^2, n] (a tensor of the residual stream)."""
        nb = self.nb
        ho = h // 2
        y = self._buf(nb * ho * ho, layer.n)
        if layer.extra.get("s2_tma", False):
            self._gemm(layer, [x], [c], nb, ho, ho, [(0, 9)], y, stream=True, conv_stride=2, conv_pad=pad_lo,
                       act=act)
        else:
            col = self._buf(nb * ho * ho, 9 * c)
            self._op(lambda: nat.im2col_s2(x, nb, h, h, c, pad_lo, col), tag=f"im2col:{nb * h * h}:")
            self._gemm(layer, [col], [9 * c], 1, 1, nb * ho * ho, [(0, 1)], y, stream=True, act=act)
        return y

    def link_weight_prefetch(self, max_rows: int = NEXTW_MAX_ROWS) -> None:
        """Give every igemm launch the weights of the next one (static parameters only) as an L2 prefetch hint;
        the last launch of the plan points at the first (the plan is replayed step after step).  Skipped for
        plans whose first launch already has more than `max_rows` output rows (compute-bound batches)."""
        ps = self._igemm_params
        if len(ps) < 2 or ps[0][2] > max_rows:
            return
        for i, (p, _, _) in enumerate(ps):
            _, nxt, _ = ps[(i + 1) % len(ps)]
            if not nxt.extra.get("static", False):
                continue
            nbytes = nxt.w.numel() * nxt.w.element_size()
            if nbytes < (1 << 16) or nbytes > (96 << 20):
                continue
            p.next_weight = nxt.w.data_ptr()
            p.next_weight_bytes = nbytes

    def _arena(self, n: int) -> Optional[torch.Tensor]:
        if self._arena_used + n > self.stats_arena.numel():
            return None
        sl = self.stats_arena[self._arena_used:self._arena_used + n]
        self._arena_used += n
        return sl

    def _stats_for(self, src, c, hw) -> Optional[torch.Tensor]:
        """Channel-statistics slice for a GroupNorm source written by an igemm of this plan (or None)."""
        if not FUSE_GN_STATS or hw % 32 != 0:
            return None
        key = src.data_ptr()
        if key in self._chan_stats:
            return self._chan_stats[key]
        prod = self._producer.get(key)
        if prod is None or prod[1] != c or prod[2] != self.nb * hw:
            return None
        if prod[0].upsample2 and (hw // 4) % 32 != 0:
            return None     # the producer's rows are INPUT pixels: a warp's 32 rows must stay inside one image
        need = self.nb * c * 2
        if self._arena_used + need > self.stats_arena.numel():
            return None
        sl = self.stats_arena[self._arena_used:self._arena_used + need]
        self._arena_used += need
        prod[0].stats = sl.data_ptr()          # the producer's launch happens later, at run time
        prod[0].stats_hw = hw // 4 if prod[0].upsample2 else hw   # the kernel's rows: input pixels when it up-samples
        self._chan_stats[key] = sl
        return sl

    def _gn(self, name, src0, c0, src1, c1, hw, silu, out):
        g, b, eps = self.W.norms[name]
        nb, groups = self.nb, self.W.groups
        cs0 = self._stats_for(src0, c0, hw)
        cs1 = self._stats_for(src1, c1, hw) if src1 is not None else None
        if cs0 is not None and (src1 is None or cs1 is not None):
            # fp32 residual stream: read the f32 copies (both sources or neither)
            f0 = self._f32.get(src0.data_ptr())
            f1 = self._f32.get(src1.data_ptr()) if src1 is not None else None
            if f0 is not None and (src1 is None or f1 is not None):
                src0, src1 = f0, f1
            self._op(lambda: nat.groupnorm_apply_cs(src0, c0, cs0, src1, c1, cs1, nb, hw, groups, g, b, eps, silu,
                                                    out), tag=f"gn:{nb * hw}:{name}:c{c0 + c1}")
            return
        stats = self.gn_stats
        self._op(lambda: nat.groupnorm(src0, c0, src1, c1, nb, hw, groups, g, b, eps, silu, out, stats), 3,
                 tag=f"gn3:{nb * hw}:{name}:c{c0 + c1}")

    def _ln(self, name, src, rows, c, out, silu=False):
        g, b, eps = self.W.norms[name]
        src = self._f32.get(src.data_ptr(), src)
        self._op(lambda: nat.layernorm(src, rows, c, g, b, eps, silu, out), tag=f"ln:{rows}:{name}")

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
        # the 1x1 shortcut rides in conv2 as extra K segments exactly when the layer was packed with it
        # (WeightsBase._resnet); otherwise the block input is added back as the residual
        if not l2.extra.get("shortcut", False):
            assert skip is None and cin == cout, f"{name}: no packed shortcut but cin {cin} != cout {cout}"
            self._gemm(l2, [a2], [cout], nb, h, w, [(0, 9)], out, residual=x, stream=True)
        elif skip is None:
            self._gemm(l2, [a2, x], [cout, cx], nb, h, w, [(0, 9), (1, 1)], out, stream=True)
        else:
            self._gemm(l2, [a2, x, skip], [cout, cx, cskip], nb, h, w, [(0, 9), (1, 1), (2, 1)], out, stream=True)
        return out

    def run(self) -> None:
        old = nat.set_pdl(self.pdl)
        try:
            for op in self.ops:
                op()
        finally:
            nat.set_pdl(old)
