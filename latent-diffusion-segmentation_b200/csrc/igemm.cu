// Implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// One kernel covers every contraction of the UNet / VAE hot path that is not attention:
// conv3x3 (9 shifted TMA box loads of the channel-last activation, zero padding by TMA
// out-of-bounds fill), stride-2 conv3x3 (TMA traversal stride), nearest x2 up-sampling + conv3x3 (four 2x2 phase
// GEMMs over the low-resolution tensor, see IgemmKParams::up2), conv1x1 / linear (1 tap), concat-free skip
// connections and the fused 1x1 shortcut (several K segments accumulating into one TMEM tile).
//
// Replaces F.conv2d / F.linear inside diffusers' ResnetBlock2D / Transformer2DModel / Attention /
// FeedForward / Downsample2D / Upsample2D as called from /root/reference/ldmseg/models/unet.py:357,
// 361-373, 388-395, 401-425, 431, and the convs of GeneralVAESeg.decode (models/vae.py:133-172).
//
// CTA = 320 threads (576 for the GEGLU epilogue), persistent over (tile, k-split) work items:
//   warp 0       TMA producer   (one elected lane)
//   warp 1       TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9   epilogue: tcgen05.ld -> bias / row-bias / residual / SiLU / GEGLU / GroupNorm statistics -> global
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA; two TMEM accumulator stages
// (tmem_full/tmem_empty) between MMA and epilogue so tile i's epilogue overlaps tile i+1's MMAs.
//
// Tile: BLOCK_M = 128 output pixels x BN output channels, BLOCK_K = 64 (one 128-byte swizzle row).
// PAIR variant: two CTAs of a cluster form one 256 x BN tile (tcgen05 cta_group::2), BN up to 320 (two N = 160
// instructions per k-step over three accumulator slots) -- see IgemmCfg.
#include "common.h"
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

constexpr int BM = 128;
constexpr int BK = 64;
// TMA warp, MMA warp, 8 epilogue warps -- 16 for the GEGLU epilogue without split-K: with K of only 320-1280 the
// epilogue of a 128 x 256 tile (two MUFU and ~9 packed-FMA instructions per output) took longer than the tile's MMAs
// and two warps per scheduler left it latency-bound (ncu at batch 8: issue 52 %, tensor 29 %; 87 -> 64 us at
// M = 32768, K = 320, N = 2560).  The plain epilogue gains nothing from 16 warps (measured: it is bound by the
// rate of its 32-byte sector stores at wide N, not by latency) and the batch-1 forward loses 0.7 %.
// The 320-wide pair tile: its accumulator slots are single-buffered halves, so the time the epilogue needs to drain a
// half stalls the next tile's MMAs (and the last tile's epilogue is exposed in full: 17 of 57 us at N = K = 320, batch
// 8, profiles/r02_bench_bn320_stages.log).  16 epilogue warps (-DLDMSEG_BN320_EPI_WARPS=16) were measured and do not
// help (57.9 against 56.1 us, profiles/r02_bench_bn320_epi16.log): like the plain epilogue above it is not latency-bound.
#ifndef LDMSEG_BN320_EPI_WARPS
#define LDMSEG_BN320_EPI_WARPS 8
#endif
__host__ __device__ constexpr int igemm_epi_warps(bool geglu, bool split, int bn = 0) {
  return bn == 320 ? LDMSEG_BN320_EPI_WARPS : (geglu && !split) ? 16 : 8;
}
__host__ __device__ constexpr int igemm_threads(bool geglu, bool split, int bn = 0) {
  return 64 + 32 * igemm_epi_warps(geglu, split, bn);
}
constexpr int kABytes = BM * BK * 2;  // 16 KB per stage

struct alignas(64) IgemmKParams {
  CUtensorMap a_map[LDMSEG_MAX_SRC];
  CUtensorMap b_map;
  int nseg;
  int seg_src[LDMSEG_MAX_SEG];
  int seg_taps[LDMSEG_MAX_SEG];
  int seg_cblocks[LDMSEG_MAX_SEG];
  int M, N, H, W, HW;
  int num_m_tiles, num_n_tiles, num_kb;
  const float* bias;
  const float* rowbias;
  int rowbias_ld;
  const void* residual;   // bf16, or f32 when res_f32 (fp32 residual stream)
  int res_ld;
  int res_f32;
  void* out;
  int out_ld;
  int out_f32;
  __nv_bfloat16* out2;    // optional bf16 shadow of an f32 output (the copy TMA consumers read)
  int out2_ld;
  int act;
  int a_stride;           // 1, or 2: stride-2 conv through the TMA traversal stride (input is 2h x 2w)
  int a_pad;              // zero padding before (top / left): tap (ky, kx) reads input (s*y + ky - pad, s*x + kx - pad)
  const uint8_t* next_w;  // weights of the NEXT igemm launch on the stream: pulled into L2 once this CTA's own
  unsigned long long next_w_bytes;  // operand loads are all in flight (weight streaming at small batch)
  int split_k;
  float* workspace;   // split-K partial tiles: [tile][split][128][BN] f32
  int* counters;
  float* stats;       // optional per-(image, channel) {sum, sum of squares} of the stored output
  int stats_hw;       // rows per image for the statistics (the producer may be a plain [M, K] GEMM)
  float* rowstats;            // optional [M, 2]: += per-row {sum, sum of squares} of the stored bf16 output (LayerNorm
                              // statistics for a consumer GEMM that folds the LayerNorm)
  const float* ln_rowstats;   // LayerNorm folded into THIS GEMM: A is the un-normalised row x, the weights carry gamma,
  const float* ln_colsum;     // out = rstd_m * (acc - mean_m * colsum[n]) + bias[n]   (bias carries W beta + b)
  float ln_inv_c, ln_eps;
  int w_tiled;        // weights stored as [N/16][K/64][16][64] blocks (2 KB contiguous per block)
  int coop_reduce;    // split-K: all epilogue warps share the reduction of each owned chunk (see splitk_final_coop)
  int prefetch_b;     // issue the first work item's weight loads before the grid-dependency wait
  int csplit;          // split-K inside a thread-block cluster: the splits of a tile exchange partials through
                       // distributed shared memory (see splitk_final_cluster)
  int tail;            // stream-K tail (see TailSeg): tiles [0, tail_full_tiles) run whole, in waves of the grid; the
  int tail_full_tiles; // tail_tiles after them are cut along K into one contiguous piece per CTA
  int tail_tiles;
  int tail_slots;      // workspace slots (partial tiles) per tail tile
  int vec_ok;         // out / residual pointers and leading dimensions allow 32-byte vector accesses
  int up2;            // nearest x2 up-sampling folded into the 3x3 convolution: the launch runs 4 phase GEMMs, phase
  int up_m_tiles;     // (py, px) = the output pixels (2y + py, 2x + px) as a 2x2 convolution of the h x w input;
  int up_nblk;        // num_m_tiles = 4 * up_m_tiles (phase-major); up_nblk = 16-row weight blocks per phase
  int debug;          // development only (ldmseg_set_debug; timing experiments, results are garbage):
                      // bit0 split-K without the final reduce, bit1 without publish / wait / reduce, bit2 without the
                      // partial store; bit3 / bit4 leave the A / B loads out after a work item's first k-block; bit5 drop
                      // the epilogue; bit6 launch the workspace split-K as clusters; bit7 no fused GroupNorm statistics
                      // (valid off the power cap only); bit9 no first-tile epilogue prefetch
};

// PAIR: two CTAs of a cluster work as one 256 x BN tile (tcgen05 cta_group::2): each CTA stages its own 128 rows of
// A and HALF of the B tile, so a k-block costs 16 KB + BN * 64 B of shared-memory ingest per SM instead of
// 16 KB + BN * 128 B -- operand ingest, not the tensor pipe, is what bounds the 1-CTA kernel (plan.py).
//
// BN = 320 (PAIR only): one tcgen05.mma is at most 256 wide, so a k-step issues TWO N = 160 instructions that share the
// staged A rows -- per multiply-add a third less shared-memory traffic than two 160-wide tiles (A is filled and
// fetched once for 320 columns), which is what bounds the N = 320 / 640 / 1280 convolutions.  2 x 320 accumulator
// columns do not fit the 512 of tensor memory, so the accumulators live in THREE 160-column slots used as a ring:
// tile i takes ring positions 2i, 2i + 1 (slots mod 3).  Tile i + 1's first half lands in the slot tile i never
// touched, its second half in the slot tile i's epilogue drains FIRST, and the MMA issuer runs the first-half
// instructions of a tile's first k-blocks ahead of the second-half ones (the stages stay occupied meanwhile), so the
// tensor core keeps working while that slot drains.
template <int BN, bool PAIR = false>
struct IgemmCfg {
  static_assert(BN != 320 || PAIR, "the 320-wide tile is built for CTA pairs");
  static constexpr int kSub = BN == 320 ? 2 : 1;      // tcgen05.mma instructions per k-step and tile
  static constexpr int kSubN = BN / kSub;             // their width
  static constexpr int kAccSlots = BN == 320 ? 3 : 2; // accumulator slots of kSubN columns in tensor memory
  static constexpr int kBRows = PAIR ? BN / 2 : BN;   // B rows staged by one CTA
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // the epilogue needs no shared memory: everything but the barriers goes to the operand ring
  static constexpr int kStages = PAIR ? ((BN <= 160) ? 8 : 6)
                                      : (BN <= 64) ? 9 : (BN <= 128) ? 7 : (BN <= 160) ? 6 : 4;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // stages + barriers (2*stages + 2 + slots) * 8 + tmem ptr, + 1024 alignment slack
  static constexpr int kSmemBytes = kStages * kStageBytes + (2 * kStages + 2 + kAccSlots) * 8 + 16 + 1024;
  static_assert(kSmemBytes <= 232448, "over the 227 KB of shared memory a CTA may use");
};

// ---- epilogue ------------------------------------------------------------------------------
// Thread = accumulator row.  tcgen05.ld (32x32b.x32) hands every thread 32 consecutive columns of ONE
// output row, and everything downstream keeps that layout:
//   * bias / per-image (time-embedding) bias: warp-broadcast float4 loads;
//   * residual and output: 256-bit vector accesses, i.e. every thread reads / writes whole 32-byte
//     sectors of its own row (no shared-memory transposition, no warp barriers, ~3 instructions per value;
//     the previous transposing epilogue spent ~16 and did not fit the instruction cache);
//   * split-K partials: workspace tiles are laid out [8-column group][row][8] so the same thread = row
//     pattern is perfectly coalesced both when the partial is stored and when it is reduced;
//   * fused GroupNorm statistics: a 5-step exchange butterfly turns 32 per-row values x 32 lanes into one
//     column sum per lane (31 shuffles per quantity), then one red.global per (lane, quantity).
// Eight epilogue warps: two per TMEM lane quadrant, taking alternate 32-column chunks of the tile.
enum { EPI_DIRECT = 0, EPI_PARTIAL = 1, EPI_FINAL = 2, EPI_DIRECT_ADD = 3 };
// -DLDMSEG_EPI_LEAN builds the epilogue without the LayerNorm-fold code (A/B of its cost on launches that do not use it)
#ifdef LDMSEG_EPI_LEAN
constexpr bool kEpiLnFold = false;
#else
constexpr bool kEpiLnFold = true;
#endif

__device__ __forceinline__ float fast_silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// GELU(x) = x/2 (1 + erf(x/sqrt2)), erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// bf16 rounding of the stored product); 2 MUFU + ~10 FMA instead of the branchy erff
__device__ __forceinline__ float fast_gelu(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = 1.061405429f;
  poly = fmaf(poly, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// Two GELUs per call on the packed-fp32x2 pipe (same Abramowitz-Stegun formula): 18 instructions per PAIR instead of
// ~17 per value.  gelu(x) = x/2 + |x|/2 * erf(|x|/sqrt2), |x|/2 = z/sqrt2 with z = |x|/sqrt2.
__device__ __forceinline__ float2 fast_gelu2(float2 x) {
  const float2 z = make_float2(fabsf(x.x) * 0.70710678118654752440f, fabsf(x.y) * 0.70710678118654752440f);
  const float2 den = f2_fma(z, f2_splat(0.3275911f), f2_splat(1.f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 poly = f2_fma(f2_splat(-1.061405429f), t, f2_splat(1.453152027f));   // negated polynomial: -(a5 t + a4) ...
  poly = f2_fma(poly, t, f2_splat(-1.421413741f));
  poly = f2_fma(poly, t, f2_splat(0.284496736f));
  poly = f2_fma(poly, t, f2_splat(-0.254829592f));
  poly = f2_mul(poly, t);
  const float2 arg = f2_mul(f2_mul(z, f2_splat(-1.4426950408889634f)), z);     // -z^2 * log2(e)
  const float2 e = make_float2(ex2_approx_f(arg.x), ex2_approx_f(arg.y));
  const float2 erf_abs = f2_fma(poly, e, f2_splat(1.f));                        // 1 - poly(t) * exp(-z^2)
  const float2 hx = f2_mul(x, f2_splat(0.5f));
  const float2 hax = f2_mul(z, f2_splat(0.70710678118654752440f));              // |x| / 2
  return f2_fma(hax, erf_abs, hx);
}

// Epilogue arguments held in registers (reading them through the parameter block from inside the loops
// made every access a load the compiler had to repeat after each global store).
struct EpiArgs {
  int M, N, HW, out_ld, res_ld, rowbias_ld, act, out_f32, split_k, stats_hw, vec_ok, res_f32, out2_ld;
  int up_w, up_off;   // folded x2 up-sampling: input width (0 = off) and this phase's 2 * py * w + px
  float ln_inv_c, ln_eps;
  float* rowstats;
  const float* ln_rowstats;
  const float* ln_colsum;
  const float* bias;
  const float* rowbias;
  const void* residual;
  void* out;
  __nv_bfloat16* out2;
  float* stats;
};

// Column sums over the warp's 32 rows: on return lane l holds sum_rows a[l].
__device__ __forceinline__ float warp_column_sums(float (&a)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool hi = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = hi ? a[i] : a[i + w];
      const float keep = hi ? a[i + w] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
  return a[0];
}

// Finish one 32-column chunk: v[i] = accumulator of (row m, column col0 + i).
// ln_mu / ln_r: LayerNorm statistics of this row when the LayerNorm is folded into this GEMM; rs1 / rs2 accumulate
// the row's {sum, sum of squares} of the stored bf16 values for a LayerNorm folded into the NEXT GEMM.
template <bool GEGLU>
__device__ __forceinline__ void epi_finish(const EpiArgs& p, float (&v)[32], int m, int img, int img_stats,
                                           int col0, int lane, float ln_mu, float ln_r, float& rs1, float& rs2) {
  const bool row_ok = m < p.M;
  const bool full = (col0 + 32 <= p.N) && p.vec_ok;
  // row of `out` this accumulator row lands in: m itself, or -- folded x2 up-sampling, m = (n, y, x) of the INPUT --
  // pixel (2y + py, 2x + px) of the output:  n 4hw + (2y + py) 2w + 2x + px  =  4m - 2x + (2 py w + px)
  const size_t mo = p.up_w ? static_cast<size_t>(4 * m - 2 * (m % p.up_w) + p.up_off) : static_cast<size_t>(m);
  if (full) {
    if (kEpiLnFold && p.ln_colsum != nullptr) {
      const float nmr = -ln_mu * ln_r;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + col0) + j);
        v[4 * j] = fmaf(ln_r, v[4 * j], nmr * s4.x);
        v[4 * j + 1] = fmaf(ln_r, v[4 * j + 1], nmr * s4.y);
        v[4 * j + 2] = fmaf(ln_r, v[4 * j + 2], nmr * s4.z);
        v[4 * j + 3] = fmaf(ln_r, v[4 * j + 3], nmr * s4.w);
      }
    }
    if (p.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
    }
    if (p.rowbias != nullptr && row_ok) {
      const float4* rb = reinterpret_cast<const float4*>(p.rowbias + static_cast<size_t>(img) * p.rowbias_ld + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b4 = __ldg(rb + j);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
    }
    if constexpr (GEGLU) {
      // chunk columns are [16 x h | 16 x g] -> 16 outputs = one 32-byte sector
      uint32_t o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 r = f2_mul(make_float2(v[2 * i], v[2 * i + 1]),
                                fast_gelu2(make_float2(v[16 + 2 * i], v[17 + 2 * i])));
        o[i] = pack_bf16x2(r.x, r.y);
      }
      if (row_ok)
        stg256(reinterpret_cast<__nv_bfloat16*>(p.out) + mo * p.out_ld + (col0 >> 1), o);
      return;
    } else {
      if (p.residual != nullptr && row_ok) {
        if (p.res_f32) {   // fp32 residual stream
          const float* rp = reinterpret_cast<const float*>(p.residual) + static_cast<size_t>(m) * p.res_ld + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t r[8];
            ldg256_nc(rp + 8 * j, r);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[8 * j + i] += __uint_as_float(r[i]);
          }
        } else {
          const __nv_bfloat16* rp =
              reinterpret_cast<const __nv_bfloat16*>(p.residual) + static_cast<size_t>(m) * p.res_ld + col0;
          uint32_t r0[8], r1[8];
          ldg256_nc(rp, r0);
          ldg256_nc(rp + 16, r1);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 a = unpack_bf16x2(r0[i]), b = unpack_bf16x2(r1[i]);
            v[2 * i] += a.x; v[2 * i + 1] += a.y; v[16 + 2 * i] += b.x; v[17 + 2 * i] += b.y;
          }
        }
      }
      if (p.act == LDMSEG_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fast_silu(v[i]);
      }
      if (kEpiLnFold && p.rowstats != nullptr && row_ok) {
        // row moments for a LayerNorm folded into the next GEMM: taken from the f32 values (the bf16 rounding of
        // the stored row is unbiased: its effect on mean / variance is ~1e-4 relative, far below the output's own
        // rounding), packed fp32x2 arithmetic
        float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 x = make_float2(v[2 * i], v[2 * i + 1]);
          a1 = f2_fma(x, f2_splat(1.f), a1);
          a2 = f2_fma(x, x, a2);
        }
        rs1 += a1.x + a1.y;
        rs2 += a2.x + a2.y;
      }
      if (p.out_f32) {
        if (row_ok) {
          float* op = reinterpret_cast<float*>(p.out) + mo * p.out_ld + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) stg256(op + 8 * j, *reinterpret_cast<uint32_t(*)[8]>(&v[8 * j]));
          if (p.out2 != nullptr) {   // bf16 shadow for the consumers that read this tensor through TMA
            uint32_t o0[8], o1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              o0[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
              o1[i] = pack_bf16x2(v[16 + 2 * i], v[17 + 2 * i]);
            }
            __nv_bfloat16* o2 = p.out2 + mo * p.out2_ld + col0;
            stg256(o2, o0);
            stg256(o2 + 16, o1);
          }
        }
      } else {
        uint32_t o0[8], o1[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o0[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          o1[i] = pack_bf16x2(v[16 + 2 * i], v[17 + 2 * i]);
        }
        if (row_ok) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + mo * p.out_ld + col0;
          stg256(op, o0);
          stg256(op + 16, o1);
        }
        if (p.stats != nullptr) {  // statistics of what the consumer will read (bf16-rounded)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 a = unpack_bf16x2(o0[i]), b = unpack_bf16x2(o1[i]);
            v[2 * i] = a.x; v[2 * i + 1] = a.y; v[16 + 2 * i] = b.x; v[17 + 2 * i] = b.y;
          }
        }
      }
    }
  } else {
    // generic path: partial chunks (N = 4, 8, ...) or unaligned leading dimensions; never GEGLU (N % 32 == 0).
    // Rare and tiny: a rolled loop over a local copy keeps it out of the instruction-cache budget.
    float tmp[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tmp[i] = v[i];
    const float* rb = p.rowbias != nullptr ? p.rowbias + static_cast<size_t>(img) * p.rowbias_ld : nullptr;
    const int ncol = row_ok ? min(32, p.N - col0) : 0;
#pragma unroll 1
    for (int i = 0; i < ncol; ++i) {
      const int col = col0 + i;
      float x = tmp[i];
      if (p.ln_colsum != nullptr) x = ln_r * (x - ln_mu * __ldg(p.ln_colsum + col));
      if (p.bias != nullptr) x += __ldg(p.bias + col);
      if (rb != nullptr) x += __ldg(rb + col);
      if (p.residual != nullptr)
        x += p.res_f32 ? reinterpret_cast<const float*>(p.residual)[static_cast<size_t>(m) * p.res_ld + col]
                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(
                             p.residual)[static_cast<size_t>(m) * p.res_ld + col]);
      if (p.act == LDMSEG_ACT_SILU) x = fast_silu(x);
      if (p.out_f32) {
        reinterpret_cast<float*>(p.out)[mo * p.out_ld + col] = x;
        if (p.out2 != nullptr) p.out2[mo * p.out2_ld + col] = __float2bfloat16(x);
      } else {
        const __nv_bfloat16 o = __float2bfloat16(x);
        reinterpret_cast<__nv_bfloat16*>(p.out)[mo * p.out_ld + col] = o;
        x = __bfloat162float(o);
      }
      if (p.rowstats != nullptr) {
        rs1 += x;
        rs2 = fmaf(x, x, rs2);
      }
      tmp[i] = x;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < ncol ? tmp[i] : 0.f;
  }
  if constexpr (!GEGLU) {
    if (p.stats != nullptr) {
      float sq[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (!row_ok) v[i] = 0.f;
        sq[i] = v[i] * v[i];
      }
      const float su = warp_column_sums(v, lane);
      const float sqs = warp_column_sums(sq, lane);
      // the warp's 32 rows belong to one image (rows per image % 32 == 0, checked on the host)
      if (col0 + lane < p.N && m - lane < p.M) {
        float* g = p.stats + (static_cast<size_t>(img_stats) * p.N + col0 + lane) * 2;
        red_add_f32x2(g, su, sqs);   // {sum, sum of squares} are adjacent: one L2 reduction instead of two
      }
    }
  }
}

// One warp, one output tile: processes the 32-column chunks half, half + NH, ...  (NH = epilogue warps per TMEM
// lane quadrant)
//   DIRECT : TMEM -> epilogue -> global            PARTIAL: TMEM -> split-K workspace
//   FINAL  : sum of the workspace partials -> epilogue -> global, for the chunks dealt to this split
//   DIRECT_ADD: TMEM + the first `split_idx` workspace partials -> epilogue -> global (stream-K tail: the CTA that
//               holds a tile's last k-blocks finishes it)
// ws_stride: f32 elements between two partials of the tile in the workspace (BM x the tile's full width)
template <int BN, bool GEGLU, int MODE, int NH>
__device__ __forceinline__ void epilogue_warp(const EpiArgs& p, uint32_t t_row, float* ws_tile, int split_idx,
                                              int m_base, int n0, int q, int half, int lane, float2 ln_rs,
                                              int ws_stride = BM * BN) {
  constexpr int kChunks = BN / 32;
  const size_t kTileElems = static_cast<size_t>(ws_stride);
  const int m = m_base + lane;
  const int row_in_tile = q * 32 + lane;
  const int img = m / p.HW;                                       // image of this row (per-image bias)
  const int img_stats = MODE == EPI_PARTIAL ? 0 : m_base / p.stats_hw;  // image of the warp's rows (statistics)
  float ln_mu = 0.f, ln_r = 1.f, rs1 = 0.f, rs2 = 0.f;
  if constexpr (MODE != EPI_PARTIAL) {
    if (p.ln_rowstats != nullptr) {
      ln_mu = ln_rs.x * p.ln_inv_c;
      ln_r = rsqrtf(fmaxf(fmaf(-ln_mu, ln_mu, ln_rs.y * p.ln_inv_c), 0.f) + p.ln_eps);
    }
  }
#pragma unroll 1
  for (int ch = half; ch < kChunks; ch += NH) {
    const int col0 = n0 + ch * 32;
    if (col0 >= p.N) break;
    float v[32];
    if constexpr (MODE == EPI_FINAL) {
      if ((q * kChunks + ch) % p.split_k != split_idx) continue;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
      const float* src = ws_tile + (static_cast<size_t>(ch) * 4 * BM + row_in_tile) * 8;
#pragma unroll 2
      for (int s = 0; s < (m < p.M ? p.split_k : 0); ++s) {
        uint32_t t[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) ldg256_cg(src + static_cast<size_t>(s) * kTileElems + g * (BM * 8), t[g]);
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * g + i] += __uint_as_float(t[g][i]);
      }
    } else {
      uint32_t r[32];
      tmem_ld_32x32(t_row + ch * 32, r);
      tmem_wait_ld();
      if constexpr (MODE == EPI_DIRECT_ADD) {
        if (m < p.M) {
          const float* src = ws_tile + (static_cast<size_t>(ch) * 4 * BM + row_in_tile) * 8;
#pragma unroll 1
          for (int s = 0; s < split_idx; ++s) {
            uint32_t t[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) ldg256_cg(src + static_cast<size_t>(s) * kTileElems + g * (BM * 8), t[g]);
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
              for (int i = 0; i < 8; ++i) r[8 * g + i] = __float_as_uint(__uint_as_float(r[8 * g + i]) + __uint_as_float(t[g][i]));
          }
        }
      }
      if constexpr (MODE == EPI_PARTIAL) {
        // rows past M (the 8x8 level fills half a tile at batch 1) are never read back
        if (m < p.M) {
          float* dst = ws_tile + static_cast<size_t>(split_idx) * kTileElems +
                       (static_cast<size_t>(ch) * 4 * BM + row_in_tile) * 8;
#pragma unroll
          for (int g = 0; g < 4; ++g) stg256(dst + g * (BM * 8), *reinterpret_cast<uint32_t(*)[8]>(&r[8 * g]));
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    }
    epi_finish<GEGLU>(p, v, m, img, img_stats, col0, lane, ln_mu, ln_r, rs1, rs2);
  }
  if constexpr (MODE != EPI_PARTIAL && !GEGLU) {
    if (p.rowstats != nullptr && m < p.M) {
      atomicAdd(p.rowstats + 2 * static_cast<size_t>(m), rs1);
      atomicAdd(p.rowstats + 2 * static_cast<size_t>(m) + 1, rs2);
    }
  }
}

// Split-K final pass, cooperative form.  A split CTA owns the (row quadrant, 32-column chunk) units u = split (mod S)
// of its tile.  In EPI_FINAL above ONE warp sums all S partials of a unit, two splits per trip: with S = 14 that is a
// chain of seven dependent L2 round trips on a single warp (~5 us of the ~14 us a weight-streaming 8x8 / 16x16
// convolution takes at batch 1) while the other seven epilogue warps idle.  Here the 8 epilogue warps form teams of
// T = 8 / (owned units, rounded up to a power of two): member j of a team sums the splits s = j (mod T) of the
// team's unit, the partial sums meet in shared memory (the operand ring is idle: a split launch gives every CTA
// exactly one work item), and member 0 finishes the chunk.  One round trip instead of S / 2.
template <int BN, bool GEGLU>
__device__ __forceinline__ void splitk_final_coop(const EpiArgs& p, const float* ws_tile, int split_idx, int tile_m0,
                                                  int n0, int ew, int lane, float* stage) {
  constexpr int kChunks = BN / 32;
  constexpr int kUnits = 4 * kChunks;
  constexpr int kTileElems = BM * BN;
  const int S = p.split_k;
  const int n_own = (kUnits - split_idx + S - 1) / S;
  if (n_own <= 0) return;
  int T = 8;
  while (T > 1 && T * n_own > 8) T >>= 1;
  const int n_teams = 8 / T, team = ew / T, j = ew - team * T;
  float4* st4 = reinterpret_cast<float4*>(stage);   // [warp][8 float4 groups][32 lanes]
  for (int k = team; k < n_own; k += n_teams) {
    const int u = split_idx + k * S;
    const int q = u / kChunks, ch = u - q * kChunks;
    const int col0 = n0 + ch * 32;
    const bool live = col0 < p.N;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    if (live && tile_m0 + q * 32 + lane < p.M) {
      const float* src = ws_tile + (static_cast<size_t>(ch) * 4 * BM + q * 32 + lane) * 8;
#pragma unroll 2
      for (int s = j; s < S; s += T) {
        uint32_t t[4][8];
#pragma unroll
        for (int g = 0; g < 4; ++g) ldg256_cg(src + static_cast<size_t>(s) * kTileElems + g * (BM * 8), t[g]);
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int i = 0; i < 8; ++i) v[8 * g + i] += __uint_as_float(t[g][i]);
      }
    }
    if (T > 1) {
      if (j != 0) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          st4[(ew * 8 + g) * 32 + lane] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(T * 32) : "memory");
      if (j == 0) {
        for (int jj = 1; jj < T; ++jj) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 a = st4[((ew + jj) * 8 + g) * 32 + lane];
            v[4 * g] += a.x; v[4 * g + 1] += a.y; v[4 * g + 2] += a.z; v[4 * g + 3] += a.w;
          }
        }
      }
      // the team's staging slices are rewritten by its next unit
      if (k + n_teams < n_own) asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(T * 32) : "memory");
    }
    if (j == 0 && live) {
      const int m = tile_m0 + q * 32 + lane;
      float ln_mu = 0.f, ln_r = 1.f, rs1 = 0.f, rs2 = 0.f;
      if (p.ln_rowstats != nullptr && m < p.M) {
        const float2 rs = __ldcg(reinterpret_cast<const float2*>(p.ln_rowstats) + m);
        ln_mu = rs.x * p.ln_inv_c;
        ln_r = rsqrtf(fmaxf(fmaf(-ln_mu, ln_mu, rs.y * p.ln_inv_c), 0.f) + p.ln_eps);
      }
      epi_finish<GEGLU>(p, v, m, m / p.HW, (tile_m0 + q * 32) / p.stats_hw, col0, lane, ln_mu, ln_r, rs1, rs2);
      if constexpr (!GEGLU) {
        if (p.rowstats != nullptr && m < p.M) {
          atomicAdd(p.rowstats + 2 * static_cast<size_t>(m), rs1);
          atomicAdd(p.rowstats + 2 * static_cast<size_t>(m) + 1, rs2);
        }
      }
    }
  }
}

// ---- split-K through distributed shared memory ---------------------------------------------------
// The S split CTAs of a tile are launched as one thread-block cluster (cluster rank = split index).  The tile's
// 4 x BN/32 (row quadrant, 32-column chunk) units are dealt round-robin to the splits; every CTA PUSHES each unit of
// its partial accumulator straight from TMEM into the shared memory of the unit's owner (st.shared::cluster, 512
// contiguous bytes per warp instruction, fire and forget), the cluster meets at a hardware barrier, and every owner
// sums its units' S slots out of its OWN shared memory and finishes them.  The operand ring is idle by then (a split
// launch gives every CTA exactly one work item).  Against the global-memory exchange above (partial tile to L2, release
// atomic, acquire spin, partials back from L2, counter re-arm: three dependent L2 round trips plus a gpu-scope
// release, measured in-graph at batch 1 as 7.6 us per split launch, 17 % of the UNet forward -- tools/ablate_unet.py
// with LDMSEG_DEBUG_FLAGS=1/2/6) nothing leaves the cluster and nothing waits on a remote load.  (Pulling the partials
// with ld.shared::cluster instead cost 3 us per launch in exposed latency, 12 us with a row-major slot layout whose
// loads were 32 separate 16-byte requests.)
// Slot layout in the owner's shared memory: [own unit k][split s][16-byte piece g 0..7][row 0..31][4 f32].
__device__ __forceinline__ uint32_t csplit_slot(uint32_t part_u32, int k, int s, int S, int lane) {
  return part_u32 + static_cast<uint32_t>(((k * S + s) * 8) * 512 + lane * 16);
}
template <int BN, bool GEGLU>
__device__ __forceinline__ void splitk_final_cluster(const EpiArgs& p, uint32_t part_u32, int S, int split_idx,
                                                     int tile_m0, int n0, int ew, int lane, float* stage, int debug) {
  constexpr int kChunks = BN / 32;
  constexpr int kUnits = 4 * kChunks;
  const int n_own = (kUnits - split_idx + S - 1) / S;
  if (n_own <= 0) return;
  int T = 8;
  while (T > 1 && T * n_own > 8) T >>= 1;
  const int n_teams = 8 / T, team = ew / T, j = ew - team * T;
  float4* st4 = reinterpret_cast<float4*>(stage);   // [warp][8 float4 groups][32 lanes]
  for (int k = team; k < n_own; k += n_teams) {
    const int u = split_idx + k * S;
    const int q = u / kChunks, ch = u - q * kChunks;
    const int col0 = n0 + ch * 32;
    const bool live = col0 < p.N;
    const int row = q * 32 + lane;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
    if (live && tile_m0 + row < p.M && !(debug & 1)) {
#pragma unroll 2
      for (int s = j; s < S; s += T) {
        const uint32_t slot = csplit_slot(part_u32, k, s, S, lane);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 a;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                       : "r"(slot + static_cast<uint32_t>(g * 512))
                       : "memory");
          v[4 * g] += a.x; v[4 * g + 1] += a.y; v[4 * g + 2] += a.z; v[4 * g + 3] += a.w;
        }
      }
    }
    if (T > 1) {
      if (j != 0) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          st4[(ew * 8 + g) * 32 + lane] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      }
      asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(T * 32) : "memory");
      if (j == 0) {
        for (int jj = 1; jj < T; ++jj) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 a = st4[((ew + jj) * 8 + g) * 32 + lane];
            v[4 * g] += a.x; v[4 * g + 1] += a.y; v[4 * g + 2] += a.z; v[4 * g + 3] += a.w;
          }
        }
      }
      // the team's staging slices are rewritten by its next unit
      if (k + n_teams < n_own) asm volatile("bar.sync %0, %1;" ::"r"(2 + team), "r"(T * 32) : "memory");
    }
    if (j == 0 && live) {
      const int m = tile_m0 + row;
      float ln_mu = 0.f, ln_r = 1.f, rs1 = 0.f, rs2 = 0.f;
      if (p.ln_rowstats != nullptr && m < p.M) {
        const float2 rs = __ldcg(reinterpret_cast<const float2*>(p.ln_rowstats) + m);
        ln_mu = rs.x * p.ln_inv_c;
        ln_r = rsqrtf(fmaxf(fmaf(-ln_mu, ln_mu, rs.y * p.ln_inv_c), 0.f) + p.ln_eps);
      }
      epi_finish<GEGLU>(p, v, m, m / p.HW, (tile_m0 + q * 32) / p.stats_hw, col0, lane, ln_mu, ln_r, rs1, rs2);
      if constexpr (!GEGLU) {
        if (p.rowstats != nullptr && m < p.M) {
          atomicAdd(p.rowstats + 2 * static_cast<size_t>(m), rs1);
          atomicAdd(p.rowstats + 2 * static_cast<size_t>(m) + 1, rs2);
        }
      }
    }
  }
}
// Pull the epilogue operands of the units this CTA will finish towards the SM while the MMAs are still running: the
// final pass is a chain of short dependent steps and every L2 round trip in it is exposed.
template <int BN>
__device__ __forceinline__ void csplit_prefetch_epilogue(const EpiArgs& p, int S, int split_idx, int tile_m0, int n0,
                                                         int ew, int lane) {
  constexpr int kChunks = BN / 32;
  constexpr int kUnits = 4 * kChunks;
  for (int u = split_idx + ew * S; u < kUnits; u += 8 * S) {
    const int q = u / kChunks, ch = u - q * kChunks;
    const int col0 = n0 + ch * 32;
    const int m = tile_m0 + q * 32 + lane;
    if (col0 >= p.N || m >= p.M) continue;
    if (p.residual != nullptr)
      prefetch_l1(reinterpret_cast<const uint8_t*>(p.residual) +
                  (static_cast<size_t>(m) * p.res_ld + col0) * (p.res_f32 ? 4 : 2));
    if (lane == 0) {
      if (p.bias != nullptr) prefetch_l1(p.bias + col0);
      if (p.rowbias != nullptr) prefetch_l1(p.rowbias + static_cast<size_t>(m / p.HW) * p.rowbias_ld + col0);
      if (p.ln_colsum != nullptr) prefetch_l1(p.ln_colsum + col0);
    }
  }
}

// ---- stream-K tail ---------------------------------------------------------------------------
// A persistent grid of G units (CTAs, or CTA pairs) runs T tiles in ceil(T / G) waves; the last wave is only
// (T mod G) / G full (N = 320 at a 64x64 latent, batch 8: 512 tiles on 148 SMs = 3.46 waves paid as 4; 256 tiles as 2).
// With `tail` set, only the floor(T / G) whole waves run as whole tiles.  The R = T mod G tiles left are laid end to
// end as U = R * num_kb k-block steps and cut into G equal contiguous pieces, one per unit (U >= G, so no piece is
// empty, and R < G, so a piece is shorter than a tile and touches at most two).  A piece that ends inside a tile is a
// PARTIAL: its accumulator goes to a workspace slot and a per-tile counter is bumped.  The unit whose piece holds a
// tile's LAST k-block finishes the tile: it waits for the counter, adds the partials to its own accumulator and runs
// the normal epilogue.  A unit whose piece spans two tiles does the head of the second tile (a partial somebody else
// waits for) BEFORE the end of the first (which it finishes itself), so no unit waits before it has published
// everything others need from it: no chains of waits, no deadlock as long as the grid is co-resident (1 CTA per SM,
// grid <= #SMs -- the same premise as split-K).
struct TailSeg {
  int tile;        // unit-tile index (a pair tile for PAIR)
  int kb0, kb1;    // k-block range
  int slot;        // workspace slot of a partial, or the number of partials to add for the finishing piece
  int finish;
};
// the unit whose piece holds step x:  the largest g with floor(g * U / G) <= x
__device__ __forceinline__ int tail_owner(long long x, int G, long long U) {
  return static_cast<int>(((x + 1) * G - 1) / U);
}
__device__ __forceinline__ int tail_plan(int tail_full_tiles, int tail_tiles, int num_kb, int unit, int G,
                                         TailSeg& s0, TailSeg& s1) {
  const long long U = static_cast<long long>(tail_tiles) * num_kb;
  const long long u0 = unit * U / G, u1 = (unit + 1) * U / G;
  if (u1 <= u0) return 0;
  const int t0 = static_cast<int>(u0 / num_kb), t1 = static_cast<int>((u1 - 1) / num_kb);
  auto fill = [&](TailSeg& s, int t, long long a, long long b) {
    const long long base = static_cast<long long>(t) * num_kb;
    s.tile = tail_full_tiles + t;
    s.kb0 = static_cast<int>(a - base);
    s.kb1 = static_cast<int>(b - base);
    s.finish = s.kb1 == num_kb;
    s.slot = unit - tail_owner(base, G, U);   // pieces of this tile before mine = my slot = what the finisher adds
  };
  if (t0 == t1) {
    fill(s0, t0, u0, u1);
    return 1;
  }
  fill(s0, t1, static_cast<long long>(t1) * num_kb, u1);   // head of the next tile first
  fill(s1, t0, u0, static_cast<long long>(t1) * num_kb);   // then the end of this one
  return 2;
}

template <int BN, bool GEGLU, bool SPLIT, bool PAIR, bool TAIL = false, bool CSPLIT = false>
__global__ void __launch_bounds__(igemm_threads(GEGLU, SPLIT, BN), 1)
igemm_kernel(const __grid_constant__ IgemmKParams p) {
  static_assert(!(TAIL && (SPLIT || GEGLU)), "the stream-K tail replaces split-K; it is not built for GEGLU");
  static_assert(!CSPLIT || (SPLIT && !PAIR && !TAIL), "the cluster exchange is a form of split-K for single CTAs");
  static_assert(BN != 320 || (PAIR && !SPLIT && !GEGLU), "the 320-wide tile: CTA pairs, whole tiles or the stream-K tail");
  using Cfg = IgemmCfg<BN, PAIR>;
  constexpr int kEpiWarps = igemm_epi_warps(GEGLU, SPLIT, BN);
  constexpr int kEpiThreads = kEpiWarps * 32;
  constexpr int NH = kEpiWarps / 4;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;   // one per accumulator slot
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + Cfg::kAccSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CTA pair: rank 0 (the leader) issues the MMAs; its full / tmem_empty barriers count both CTAs
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // CTA or CTA pair
  const int num_units = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.a_map[p.seg_src[s]]);
    tma_prefetch_desc(&p.b_map);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&tmem_full[i], 1);
    for (int i = 0; i < Cfg::kAccSlots; ++i) mbar_init(&tmem_empty[i], PAIR ? 2 * kEpiThreads : kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_ptr_smem, Cfg::kTmemCols);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here
  // on we touch memory it produced.  (No-ops when the launch carries no PDL attribute.)
  // Programmatic dependent launch.  Trigger only now that this CTA owns its TMEM columns: a dependent CTA that
  // became co-resident earlier could otherwise take them and starve this (prerequisite) grid forever.
  pdl_trigger();

  // work item = (tile, k-split); PAIR: "tile" is a pair tile (two vertically adjacent 128-row tiles x BN) and each
  // CTA of the pair owns the 128-row tile 2 * m_pair + rank (a phantom one past the end when num_m_tiles is odd:
  // its loads are zero-filled by TMA and its rows are masked in the epilogue)
  const int num_tiles = (PAIR ? (p.num_m_tiles + 1) / 2 : p.num_m_tiles) * p.num_n_tiles;
  const int total_work = num_tiles * p.split_k;
  // The it-th work item of this unit: (unit tile, k-split index, k-block range).  TAIL: whole tiles first, then the
  // unit's one or two pieces of the tail (split = 0 whole tile, 1 partial, 2 finishing piece; slot as in TailSeg).
  TailSeg ts0 = {0, 0, 0, 0, 0}, ts1 = {0, 0, 0, 0, 0};
  int tail_n = 0, n_whole = 0;
  if constexpr (TAIL) {
    n_whole = p.tail_full_tiles / num_units;
    tail_n = tail_plan(p.tail_full_tiles, p.tail_tiles, p.num_kb, unit, num_units, ts0, ts1);
  }
  auto get_item = [&](int it, int& tile, int& split, int& kb_begin, int& kb_end, int& slot) -> bool {
    if constexpr (TAIL) {
      if (it >= n_whole + tail_n) return false;
      if (it < n_whole) {
        tile = unit + it * num_units;
        split = 0;
        kb_begin = 0;
        kb_end = p.num_kb;
        slot = 0;
      } else {
        const bool first = it == n_whole;
        tile = first ? ts0.tile : ts1.tile;
        kb_begin = first ? ts0.kb0 : ts1.kb0;
        kb_end = first ? ts0.kb1 : ts1.kb1;
        slot = first ? ts0.slot : ts1.slot;
        split = (first ? ts0.finish : ts1.finish) ? 2 : 1;
      }
      return true;
    } else {
      const int wi = unit + it * num_units;
      if (wi >= total_work) return false;
      tile = wi / p.split_k;
      split = wi - tile * p.split_k;
      kb_begin = static_cast<int>(static_cast<long long>(split) * p.num_kb / p.split_k);
      kb_end = static_cast<int>(static_cast<long long>(split + 1) * p.num_kb / p.split_k);
      slot = 0;
      return true;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      // PAIR: the loads of BOTH CTAs complete on the leader's full barrier (shared::cluster address); only the
      // leader arms it, with the bytes of both.  A peer load may land before the leader has armed the phase: the
      // transaction count just goes negative until the leader's arrive.expect_tx, the phase cannot complete early.
      const uint32_t full0 = PAIR ? map_to_cta(&full_bar[0], 0) : 0u;
      auto arm = [&](int stage, uint32_t cta_bytes = Cfg::kStageBytes) {
        if (!PAIR || rank == 0) mbar_expect_tx(&full_bar[stage], PAIR ? 2 * cta_bytes : cta_bytes);
      };
      // blk0: first 16-row weight block of the launch's phase (folded up-sampling: the 4 phase matrices are stacked
      // along N, block-tiled weights only), else 0
      auto load_b = [&](int stage, int kb, int n_tile, int blk0) {
        if constexpr (PAIR) {
          // per tcgen05.mma of the k-step (one, or two for BN = 320): this CTA's half of that instruction's B rows
#pragma unroll
          for (int j = 0; j < Cfg::kSub; ++j)
            tma_load_4d_2sm(smem_b + stage * Cfg::kBBytes + j * (Cfg::kBBytes / Cfg::kSub), &p.b_map,
                            full0 + stage * 8, 0, 0, kb,
                            blk0 + n_tile * (BN / 16) + j * (Cfg::kSubN / 16) +
                                static_cast<int>(rank) * (Cfg::kSubN / 32));
        } else {
          if (p.w_tiled)
            tma_load_4d(smem_b + stage * Cfg::kBBytes, &p.b_map, &full_bar[stage], 0, 0, kb,
                        blk0 + n_tile * (BN / 16));
          else
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.b_map, &full_bar[stage], kb * BK, n_tile * BN);
        }
      };
      // (virtual) 128-row tile -> up-sampling phase 2 py + px (0 for every other launch)
      auto phase_of = [&](int unit_tile) {
        const int m_unit = unit_tile / p.num_n_tiles;
        return p.up2 ? (PAIR ? 2 * m_unit + static_cast<int>(rank) : m_unit) / p.up_m_tiles : 0;
      };
      // Weights are never written inside the stream, so their loads need not wait for the previous kernel:
      // the first work item's weight tiles go into the (still empty) ring while the predecessor is still running
      // (-3 % on the batch-1 forward).  Also pulling the rest of the weight slice into L2 from here was measured
      // and dropped: at batch 8 many M-tiles share a slice and the redundant prefetches cost more than they hide.
      int prefetched = 0;
      int tile, split, kb_begin, kb_end, slot;
      if (p.prefetch_b && get_item(0, tile, split, kb_begin, kb_end, slot)) {
        const int n_tile = tile % p.num_n_tiles;
        const int blk0 = phase_of(tile) * p.up_nblk;
        prefetched = min(kStages, kb_end - kb_begin);
        for (int i = 0; i < prefetched; ++i) {
          arm(i);
          load_b(i, kb_begin + i, n_tile, blk0);
        }
      }
      pdl_wait();  // activations (and everything else the predecessor wrote) from here on
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; get_item(it, tile, split, kb_begin, kb_end, slot); ++it) {
        const int m_unit = tile / p.num_n_tiles;
        const int n_tile = tile - m_unit * p.num_n_tiles;
        int m_tile = PAIR ? 2 * m_unit + static_cast<int>(rank) : m_unit;
        int up_ph = 0;   // up-sampling phase 2 py + px
        if (p.up2) {     // phase-major virtual tiles: 128 INPUT pixels, output pixels (2y + py, 2x + px)
          up_ph = m_tile / p.up_m_tiles;
          m_tile -= up_ph * p.up_m_tiles;
        }
        const int blk0 = up_ph * p.up_nblk;
        const int m0 = m_tile * BM;
        const int x0 = (m0 % p.W) * p.a_stride;
        const int y0 = ((m0 / p.W) % p.H) * p.a_stride;
        const int b0 = m0 / p.HW;
        // decode kb_begin -> (segment, tap, channel block)
        int seg = 0, rem = kb_begin;
        while (seg < p.nseg - 1 && rem >= p.seg_taps[seg] * p.seg_cblocks[seg]) {
          rem -= p.seg_taps[seg] * p.seg_cblocks[seg];
          ++seg;
        }
        int tap = rem / p.seg_cblocks[seg];
        int cb = rem - tap * p.seg_cblocks[seg];
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const bool b_done = prefetched > 0;  // this stage's weight tile is already in flight
          // development only (timing experiments, results are garbage): leave an operand's stage contents stale
          const bool skip_a = (p.debug & 8) && kb > kb_begin, skip_b = (p.debug & 16) && kb > kb_begin;
          if (b_done) {
            --prefetched;
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            arm(stage, (skip_a ? 0 : kABytes) + (skip_b ? 0 : Cfg::kBBytes));
          }
          int dx = 0, dy = 0;
          if (p.seg_taps[seg] == 9) {
            dy = tap / 3 - p.a_pad;
            dx = tap - (tap / 3) * 3 - p.a_pad;
          } else if (p.seg_taps[seg] == 4) {   // 2x2 phase kernel: tap (a, b) reads input (y + py + a - 1, x + px + b - 1)
            dy = (tap >> 1) + (up_ph >> 1) - 1;
            dx = (tap & 1) + (up_ph & 1) - 1;
          }
          if (!skip_a) {
            if constexpr (PAIR)
              tma_load_4d_2sm(smem_a + stage * kABytes, &p.a_map[p.seg_src[seg]], full0 + stage * 8, cb * BK,
                              x0 + dx, y0 + dy, b0);
            else
              tma_load_4d(smem_a + stage * kABytes, &p.a_map[p.seg_src[seg]], &full_bar[stage],
                          cb * BK, x0 + dx, y0 + dy, b0);
          }
          if (!b_done && !skip_b) load_b(stage, kb, n_tile, blk0);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
          if (++cb == p.seg_cblocks[seg]) {
            cb = 0;
            if (++tap == p.seg_taps[seg]) {
              tap = 0;
              ++seg;
            }
          }
        }
      }
      // All of this CTA's operand loads are in flight.  From here to the end of the launch (MMA tail, split-K
      // exchange, epilogue, drain, the norm kernel in between, the successor's prologue) DRAM would idle: pull this
      // CTA's slice of the NEXT igemm launch's weights into L2 (a hint: no completion, no dependency).
      if (p.next_w != nullptr) {
        constexpr unsigned long long kPiece = 8192;
        const unsigned long long pieces = (p.next_w_bytes + kPiece - 1) / kPiece;
        for (unsigned long long i = blockIdx.x; i < pieces; i += gridDim.x) {
          const unsigned long long off = i * kPiece;
          const unsigned long long len = p.next_w_bytes - off < kPiece ? p.next_w_bytes - off : kPiece;
          bulk_prefetch_l2(p.next_w + off, static_cast<uint32_t>(len & ~15ull));
        }
      }
    }
    __syncwarp();
    if constexpr (CSPLIT) {   // barriers 1 and 2 of 3 (see the epilogue warps)
      cluster_sync_all();
      cluster_sync_all();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if ((!PAIR || rank == 0) && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, Cfg::kSubN, 0, 0);
      // descriptor = constant fields + start address >> 4 (stage s adds s * stage bytes >> 4; never carries out
      // of the 14-bit field: shared memory is < 256 KB)
      const uint64_t desc_base = make_smem_desc_sw128(0, 16, 1024);
      const uint32_t a_lo = (smem_u32(smem_a) & 0x3FFFF) >> 4, b_lo = (smem_u32(smem_b) & 0x3FFFF) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int tile, split, kb_begin, kb_end, slot;
      for (int it = 0; get_item(it, tile, split, kb_begin, kb_end, slot); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        if constexpr (Cfg::kSub == 2) {
          // 320-wide tile: halves into the ring slots r0 = 2 it and r1 = 2 it + 1 (mod 3).  Slot r1 is the one the
          // previous tile's epilogue drains first; until it is free, run the first-half MMAs of the tile's first
          // k-blocks ahead (their stages stay occupied), then catch up with the second halves and release the stages.
          const int r0 = 2 * it, r1 = 2 * it + 1;
          const int s0 = r0 % 3, s1 = r1 % 3;
          mbar_wait(&tmem_empty[s0], ((r0 / 3) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d0 = tmem_base + s0 * Cfg::kSubN, d1 = tmem_base + s1 * Cfg::kSubN;
          constexpr uint32_t kSubDesc = (Cfg::kBBytes / 2) >> 4;   // second half of the stage's B rows
          const int nkb = kb_end - kb_begin;
          const int ahead = nkb < kStages - 1 ? nkb : kStages - 1;
          int st = stage;
          uint32_t ph = phase;
          for (int i = 0; i < ahead; ++i) {
            mbar_wait(&full_bar[st], ph);
            tc_fence_after();
            const uint64_t adesc = desc_base + (a_lo + st * (kABytes >> 4));
            const uint64_t bdesc = desc_base + (b_lo + st * (Cfg::kBBytes >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(d0, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            if (++st == kStages) {
              st = 0;
              ph ^= 1;
            }
          }
          mbar_wait(&tmem_empty[s1], ((r1 / 3) & 1) ^ 1);
          tc_fence_after();
          for (int i = 0; i < ahead; ++i) {
            const uint64_t adesc = desc_base + (a_lo + stage * (kABytes >> 4));
            const uint64_t bdesc = desc_base + (b_lo + stage * (Cfg::kBBytes >> 4)) + kSubDesc;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_2sm(d1, adesc + 2 * k, bdesc + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          for (int kb = kb_begin + ahead; kb < kb_end; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t adesc = desc_base + (a_lo + stage * (kABytes >> 4));
            const uint64_t bdesc = desc_base + (b_lo + stage * (Cfg::kBBytes >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16_2sm(d0, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
              umma_bf16_2sm(d1, adesc + 2 * k, bdesc + kSubDesc + 2 * k, idesc, 1u);
            }
            umma_commit_2sm(&empty_bar[stage]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          umma_commit_2sm(&tmem_full[acc]);
          continue;
        }
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        // One k-block (4 MMAs) per barrier wait.  Issuing two k-blocks per trip was measured and dropped: the
        // MMA-only rate (operand loads left out, tools/bench_ingest.py) stays at 57 % / 65 % / 77 % of 4096 MAC/clk
        // for BN 128 / 160 / 256 either way -- a fixed ~58 cycles per M=128, K=16 instruction on top of 0.42 * BN,
        // also with cta_group::2 -- and the batch-1 forward got 1.4 % slower (the first MMA waits for two stages).
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = desc_base + (a_lo + stage * (kABytes >> 4));
          const uint64_t bdesc = desc_base + (b_lo + stage * (Cfg::kBBytes >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the >>4 field
            if constexpr (PAIR)
              umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
            else
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          if constexpr (PAIR) umma_commit_2sm(&empty_bar[stage]);   // frees the stage in both CTAs
          else umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (PAIR) umma_commit_2sm(&tmem_full[acc]);
        else umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
    if constexpr (CSPLIT) {   // barriers 1 and 2 of 3
      cluster_sync_all();
      cluster_sync_all();
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    pdl_wait();
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;    // which 32-column chunks of the tile this warp takes
    const int et = threadIdx.x - 64;     // 0..255
    EpiArgs ea;
    ea.M = p.M; ea.N = p.N; ea.HW = p.HW; ea.out_ld = p.out_ld; ea.res_ld = p.res_ld;
    ea.rowbias_ld = p.rowbias_ld; ea.act = p.act; ea.out_f32 = p.out_f32; ea.split_k = p.split_k;
    ea.stats_hw = p.stats_hw; ea.vec_ok = p.vec_ok; ea.bias = p.bias; ea.rowbias = p.rowbias;
    ea.residual = p.residual; ea.out = p.out;
    // debug bit7: no fused GroupNorm statistics (timing only -- and only where the GPU is not power-capped: the
    // garbage activations that follow toggle fewer bits, and at batch 8 the clocks rise by more than the work saved)
    ea.stats = (p.debug & 128) ? nullptr : p.stats;
    ea.res_f32 = p.res_f32; ea.out2 = p.out2; ea.out2_ld = p.out2_ld;
    ea.rowstats = p.rowstats; ea.ln_rowstats = p.ln_rowstats; ea.ln_colsum = p.ln_colsum;
    ea.ln_inv_c = p.ln_inv_c; ea.ln_eps = p.ln_eps;
    ea.up_w = p.up2 ? p.W : 0; ea.up_off = 0;
    const uint32_t tmem_empty0 = PAIR ? map_to_cta(&tmem_empty[0], 0) : 0u;   // the leader's barrier
    auto release_acc = [&](int acc) {
      if constexpr (PAIR) mbar_arrive_cluster(tmem_empty0 + acc * 8);
      else mbar_arrive(&tmem_empty[acc]);
    };
    // One tile's accumulator through epilogue_warp<MODE> and back to the MMA issuer.  320-wide tiles: the two 160-wide
    // halves one after the other (ring slots 2 it, 2 it + 1 mod 3), each slot released as soon as it is drained -- the
    // next tile's second half is waiting for the FIRST one; the warps of a quadrant rotate their chunk residue for the
    // second half (5 chunks per half dealt to NH warps: whoever took two of the first half takes one of the second).
    auto run_epilogue = [&](auto mode, int it, uint32_t t_row, float* ws_tile, int sidx, int m_base, int n0, int q,
                            int half, float2 ln_rs) {
      constexpr int MODE = decltype(mode)::value;
      if constexpr (Cfg::kSub == 1) {
        epilogue_warp<BN, GEGLU, MODE, NH>(ea, t_row, ws_tile, sidx, m_base, n0, q, half, lane, ln_rs);
        tc_fence_before();
        release_acc(it & 1);
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int s = (2 * it + j) % 3;
          epilogue_warp<Cfg::kSubN, GEGLU, MODE, NH>(
              ea, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * Cfg::kSubN,
              ws_tile != nullptr ? ws_tile + j * (BM * Cfg::kSubN) : nullptr, sidx, m_base, n0 + j * Cfg::kSubN, q,
              (half + j * (NH / 2)) % NH, lane, ln_rs, BM * BN);
          tc_fence_before();
          release_acc(s);
        }
      }
    };
    using ModeDirect = std::integral_constant<int, EPI_DIRECT>;
    using ModePartial = std::integral_constant<int, EPI_PARTIAL>;
    using ModeDirectAdd = std::integral_constant<int, EPI_DIRECT_ADD>;
    int unit_tile, split, kb_begin_, kb_end_, slot;
    for (int it = 0; get_item(it, unit_tile, split, kb_begin_, kb_end_, slot); ++it) {
      const int m_unit = unit_tile / p.num_n_tiles;
      const int n_tile = unit_tile - m_unit * p.num_n_tiles;
      const int vm_tile = PAIR ? 2 * m_unit + static_cast<int>(rank) : m_unit;
      const int tile = vm_tile * p.num_n_tiles + n_tile;   // 128-row output tile (split-K workspace / counters)
      int m_tile = vm_tile;   // 128 rows of M (folded up-sampling: of the INPUT pixels; vm_tile = phase-major)
      if (p.up2) {
        const int up_ph = vm_tile / p.up_m_tiles;
        m_tile -= up_ph * p.up_m_tiles;
        ea.up_off = (up_ph >> 1) * 2 * p.W + (up_ph & 1);
      }
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_base = m_tile * BM + q * 32;
      const int n0 = n_tile * BN;
      // a folded LayerNorm's row moments come from an earlier launch: fetch them while the MMAs are still running
      float2 ln_rs = make_float2(0.f, 1.f);
      if (ea.ln_rowstats != nullptr && m_base + lane < ea.M)
        ln_rs = __ldcg(reinterpret_cast<const float2*>(ea.ln_rowstats) + m_base + lane);
      if constexpr (CSPLIT) csplit_prefetch_epilogue<BN>(ea, p.split_k, split, m_tile * BM, n0, warp - 2, lane);
      if constexpr (!SPLIT) {
        // First tile of this CTA: the epilogue warps have nothing to do until the MMAs finish, and what follows is a
        // chain of short dependent steps per 32-column chunk (TMEM load -> bias -> residual -> store).  Start the
        // residual rows and the bias vectors of the tile on their way to L1 now.  (Later tiles of a persistent CTA
        // overlap their epilogue with the next tile's MMAs anyway.)
        if (it == 0 && !(p.debug & 512)) {
          const int m = m_base + lane;
          const int ncols = min(BN, ea.N - n0);
          if (ea.residual != nullptr && m < ea.M) {
            const int esz = ea.res_f32 ? 4 : 2;
            const uint8_t* rp = reinterpret_cast<const uint8_t*>(ea.residual) +
                                (static_cast<size_t>(m) * ea.res_ld + n0) * esz;
            for (int off = half * 128; off < ncols * esz; off += NH * 128) prefetch_l1(rp + off);
          }
          if (q == 0 && lane < (ncols * 4 + 127) / 128 && (lane % NH) == half) {
            if (ea.bias != nullptr) prefetch_l1(ea.bias + n0 + lane * 32);
            if (ea.rowbias != nullptr && m_base < ea.M)
              prefetch_l1(ea.rowbias + static_cast<size_t>(m_base / ea.HW) * ea.rowbias_ld + n0 + lane * 32);
            if (ea.ln_colsum != nullptr) prefetch_l1(ea.ln_colsum + n0 + lane * 32);
          }
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if constexpr (CSPLIT) {
        // split-K inside a cluster (one work item per CTA; cluster rank = split): every unit of the partial goes to
        // the shared memory of the split that owns the unit
        const uint32_t part_u32 = smem_u32(smem_a);
        constexpr int kChunksC = BN / 32;
        const int S = p.split_k;
        // barrier 1 of 3: this CTA's MMAs are done (tmem_full), its operand ring may be overwritten; nobody pushes
        // before every CTA of the cluster has said so (a fast split would otherwise write into a ring its slower
        // peer is still multiplying from).  Arrive now, wait just before the first remote store.
        cluster_arrive();
        bool waited = false;
#pragma unroll 1
        for (int ch = half; ch < kChunksC; ch += NH) {
          if (n0 + ch * 32 >= p.N) break;
          uint32_t r[32];
          tmem_ld_32x32(t_row + ch * 32, r);
          tmem_wait_ld();
          if (!waited) {
            cluster_wait();
            waited = true;
          }
          const int u = q * kChunksC + ch;
          const int owner = u % S, k_own = u / S;
          uint32_t dst;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;"
                       : "=r"(dst)
                       : "r"(csplit_slot(part_u32, k_own, split, S, lane)), "r"(static_cast<uint32_t>(owner)));
#pragma unroll
          for (int g = 0; g < 8; ++g)
            st_dsmem_b4(dst + static_cast<uint32_t>(g * 512), r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
        }
        if (!waited) cluster_wait();
        tc_fence_before();
        release_acc(acc);
        cluster_sync_all();   // barrier 2 of 3 (release / acquire: every split's pushes have landed)
        if (!(p.debug & 2))   // debug bit1: no final pass
          splitk_final_cluster<BN, GEGLU>(ea, part_u32, S, split, m_tile * BM, n0, warp - 2, lane,
                                          reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes - 32768), p.debug);
      } else if constexpr (TAIL) {
        // stream-K tail: per 128-row tail tile a run of workspace slots and one arrival counter
        const int tt = (unit_tile - p.tail_full_tiles) * (PAIR ? 2 : 1) + static_cast<int>(rank);
        float* ws_tile = p.workspace + static_cast<size_t>(tt > 0 ? tt : 0) * p.tail_slots * (BM * BN);
        if (split == 1) {
          // a piece that ends inside the tile: publish the partial accumulator
          run_epilogue(ModePartial{}, it, t_row, ws_tile, slot, m_base, n0, q, half, ln_rs);
          asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");   // every thread's stores before thread 0's release
          if (et == 0) {
            int seen;
            asm volatile("atom.release.gpu.global.add.s32 %0, [%1], 1;" : "=r"(seen) : "l"(p.counters + tt) : "memory");
          }
        } else if (split == 2 && slot > 0) {
          // the piece with the tile's last k-block: add the `slot` partials published before it
          if (et == 0) {
            int seen;
            uint32_t spins = 0;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.counters + tt) : "memory");
              if (++spins > (1u << 26)) __trap();
            } while (seen < slot);
          }
          asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");
          run_epilogue(ModeDirectAdd{}, it, t_row, ws_tile, slot, m_base, n0, q, half, ln_rs);
          asm volatile("bar.sync 1, %0;" ::"r"(kEpiThreads) : "memory");
          if (et == 0) p.counters[tt] = 0;   // re-armed for the next launch (all `slot` arrivals were consumed)
        } else {
          run_epilogue(ModeDirect{}, it, t_row, nullptr, 0, m_base, n0, q, half, ln_rs);
        }
      } else if constexpr (!SPLIT) {
        if (p.debug & 32) {   // development only: drop the tile (is the epilogue the bottleneck?)
          tc_fence_before();
          if constexpr (Cfg::kSub == 1) {
            release_acc(acc);
          } else {
            release_acc((2 * it) % 3);
            release_acc((2 * it + 1) % 3);
          }
          continue;
        }
        run_epilogue(ModeDirect{}, it, t_row, nullptr, 0, m_base, n0, q, half, ln_rs);
      } else {
        // split-K: every split stores its partial tile (coalesced, no atomics); once all splits of the
        // tile have arrived, each split CTA reduces and finishes its share of the tile's chunks.
        float* ws_tile = p.workspace + static_cast<size_t>(tile) * p.split_k * (BM * BN);
        if (!(p.debug & 4))
          epilogue_warp<BN, GEGLU, EPI_PARTIAL, NH>(ea, t_row, ws_tile, split, m_base, n0, q, half, lane, ln_rs);
        tc_fence_before();
        release_acc(acc);
        // publish + wait for the peers: the CTA barrier orders every thread's partial stores before
        // thread 0's release-atomic (release is cumulative); thread 0 then polls with acquire loads
        // until all splits of this tile have arrived.  Peers are CTAs of the same persistent grid in
        // the same round, so they are running or about to be scheduled (1 CTA per SM, grid <= #SMs).
        if (p.debug & 2) continue;
        // the operands of the final pass that do not depend on the peers (bias, per-image bias, residual rows of the
        // units this split will finish) start their trip from L2 now, under the publish / wait below
        csplit_prefetch_epilogue<BN>(ea, p.split_k, split, m_tile * BM, n0, warp - 2, lane);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          // One self-re-arming counter per tile, counting modulo 2 S with wrapping increments: S arrivals (release),
          // then S departures.  A split spins until the count is >= S; it departs only after that, so the count wraps
          // to 0 -- the state the next launch expects -- only once nobody can still be spinning.  The departure is a
          // fire-and-forget reduction: no thread waits for an atomic's round trip before the CTA retires (the earlier
          // scheme -- a second counter, atom.acq_rel, the last split resets both -- held every CTA for that trip).
          const unsigned wrap = 2u * static_cast<unsigned>(p.split_k) - 1u;
          unsigned seen;
          asm volatile("red.release.gpu.global.inc.u32 [%0], %1;" ::"l"(p.counters + tile), "r"(wrap) : "memory");
          uint32_t spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.counters + tile) : "memory");
            if (++spins > (1u << 26)) __trap();
          } while (seen < static_cast<unsigned>(p.split_k));
          asm volatile("red.relaxed.gpu.global.inc.u32 [%0], %1;" ::"l"(p.counters + tile), "r"(wrap) : "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!(p.debug & 1)) {
          // one work item per CTA (always the case for the planner's split launches): nothing else touches the
          // operand ring any more, its first 32 KB stage the cooperative reduction
          if (p.coop_reduce && total_work <= num_units)
            splitk_final_coop<BN, GEGLU>(ea, ws_tile, split, m_tile * BM, n0, warp - 2, lane,
                                         reinterpret_cast<float*>(smem_a));
          else
            epilogue_warp<BN, GEGLU, EPI_FINAL, NH>(ea, 0, ws_tile, split, m_base, n0, q, half, lane, ln_rs);
        }
      }
    }
  }

  tc_fence_before();
  // (CSPLIT: barrier 3 of 3 -- no CTA of a cluster exits while barrier traffic or pushes could still target it)
  if constexpr (PAIR || CSPLIT) cluster_sync_all();   // the peer may still signal this CTA's barriers / read its operands
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// Reference-grade CUDA-core kernel with the same contract (one thread per output element).
__global__ void igemm_simple_kernel(ldmseg_igemm_params p, int M, int HW) {
  pdl_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int ncols = p.n;
  if (idx >= static_cast<long long>(M) * ncols) return;
  const int m = static_cast<int>(idx / ncols);
  const int n = static_cast<int>(idx - static_cast<long long>(m) * ncols);
  const int x = m % p.w, y = (m / p.w) % p.h, b = m / HW;
  const int stride = p.conv_stride == 2 ? 2 : 1;
  const int pad = p.conv_stride == 2 ? p.conv_pad : 1;
  const int hin = p.h * stride, win = p.w * stride;
  const __nv_bfloat16* wbase = reinterpret_cast<const __nv_bfloat16*>(p.weight);
  const int kblocks = p.ktot / 64;
  auto wat = [&](int k) -> float {
    const size_t off = p.weight_tiled
                           ? ((static_cast<size_t>(n >> 4) * kblocks + (k >> 6)) * 16 + (n & 15)) * 64 + (k & 63)
                           : static_cast<size_t>(n) * p.ktot + k;
    return __bfloat162float(wbase[off]);
  };
  float acc = 0.f;
  int koff = 0;
  for (int s = 0; s < p.nseg; ++s) {
    const int src = p.seg_src[s];
    const int C = p.src_c[src];
    const int Cpad = (C + 63) / 64 * 64;
    const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(p.src[src]);
    for (int tap = 0; tap < p.seg_taps[s]; ++tap) {
      int dy = 0, dx = 0;
      if (p.seg_taps[s] == 9) {
        dy = tap / 3 - pad;
        dx = tap % 3 - pad;
      }
      const int yy = y * stride + dy, xx = x * stride + dx;
      if (yy >= 0 && yy < hin && xx >= 0 && xx < win) {
        const __nv_bfloat16* arow =
            a + (static_cast<size_t>(b) * hin * win + static_cast<size_t>(yy) * win + xx) * C;
        for (int c = 0; c < C; ++c)
          acc += __bfloat162float(arow[c]) * wat(koff + c);
      }
      koff += Cpad;
    }
  }
  // epilogue identical to the tcgen05 kernel (GEGLU handled by the pair thread layout below)
  if (p.ln_colsum) {
    const float mu = p.ln_rowstats[2 * static_cast<size_t>(m)] / static_cast<float>(p.ln_channels);
    const float var = fmaxf(p.ln_rowstats[2 * static_cast<size_t>(m) + 1] / static_cast<float>(p.ln_channels) - mu * mu, 0.f);
    acc = rsqrtf(var + p.ln_eps) * (acc - mu * p.ln_colsum[n]);
  }
  if (p.bias) acc += p.bias[n];
  if (p.rowbias) acc += p.rowbias[static_cast<size_t>(b) * p.rowbias_ld + n];
  if (p.act == LDMSEG_ACT_GEGLU) {
    // interleaved [16 h | 16 g] per 32 columns; the g thread combines with its h partner through
    // a recomputation-free trick: only h threads write, after fetching g via shuffle.
    const int j = n & 31;
    const float other = __shfl_xor_sync(0xffffffffu, acc, 16);
    if (j < 16) {
      const float o = acc * gelu_erf_f(other);
      reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<size_t>(m) * p.out_ld + (n >> 5) * 16 +
                                              j] = __float2bfloat16(o);
    }
    return;
  }
  if (p.residual)
    acc += p.residual_f32
               ? reinterpret_cast<const float*>(p.residual)[static_cast<size_t>(m) * p.res_ld + n]
               : __bfloat162float(
                     reinterpret_cast<const __nv_bfloat16*>(p.residual)[static_cast<size_t>(m) * p.res_ld + n]);
  if (p.act == LDMSEG_ACT_SILU) acc = silu_f(acc);
  if (p.out_dtype == LDMSEG_OUT_F32) {
    reinterpret_cast<float*>(p.out)[static_cast<size_t>(m) * p.out_ld + n] = acc;
    if (p.out2) reinterpret_cast<__nv_bfloat16*>(p.out2)[static_cast<size_t>(m) * p.out2_ld + n] = __float2bfloat16(acc);
  } else {
    const __nv_bfloat16 o = __float2bfloat16(acc);
    reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<size_t>(m) * p.out_ld + n] = o;
    acc = __bfloat162float(o);
  }
  if (p.rowstats_out) {
    const float xb = __bfloat162float(__float2bfloat16(acc));
    atomicAdd(p.rowstats_out + 2 * static_cast<size_t>(m), xb);
    atomicAdd(p.rowstats_out + 2 * static_cast<size_t>(m) + 1, xb * xb);
  }
  if (p.stats) {
    const int sb = p.stats_hw > 0 ? m / p.stats_hw : b;
    atomicAdd(p.stats + (static_cast<size_t>(sb) * p.n + n) * 2, acc);
    atomicAdd(p.stats + (static_cast<size_t>(sb) * p.n + n) * 2 + 1, acc * acc);
  }
}

// ------------------------------------------------------------------------------------------
int g_debug = 0;
static int validate(const ldmseg_igemm_params* p) {
  LDM_REQUIRE(p != nullptr, "igemm: null params");
  LDM_REQUIRE(p->nsrc >= 1 && p->nsrc <= LDMSEG_MAX_SRC, "igemm: nsrc out of range");
  LDM_REQUIRE(p->nseg >= 1 && p->nseg <= LDMSEG_MAX_SEG, "igemm: nseg out of range");
  LDM_REQUIRE(p->nb > 0 && p->h > 0 && p->w > 0 && p->n > 0, "igemm: bad geometry");
  const long long hw = static_cast<long long>(p->h) * p->w;
  if (p->w >= BM) {
    LDM_REQUIRE(p->w % BM == 0 || (p->h == 1 && p->nb == 1), "igemm: w >= 128 must be a multiple of 128");
  } else {
    LDM_REQUIRE(BM % p->w == 0, "igemm: w < 128 must divide 128 (got %d)", p->w);
    if (hw >= BM)
      LDM_REQUIRE(hw % BM == 0, "igemm: h*w must be a multiple of 128");
    else
      LDM_REQUIRE(BM % hw == 0, "igemm: h*w must divide 128");
  }
  int ktot = 0;
  for (int s = 0; s < p->nseg; ++s) {
    LDM_REQUIRE(p->seg_src[s] >= 0 && p->seg_src[s] < p->nsrc, "igemm: bad seg_src");
    LDM_REQUIRE(p->seg_taps[s] == 1 || p->seg_taps[s] == 9 || (p->upsample2 && p->seg_taps[s] == 4),
                "igemm: taps must be 1 or 9 (4 with upsample2)");
    const int c = p->src_c[p->seg_src[s]];
    LDM_REQUIRE(c > 0 && c % 8 == 0, "igemm: source channels must be a multiple of 8 (got %d)", c);
    ktot += p->seg_taps[s] * ((c + BK - 1) / BK * BK);
  }
  LDM_REQUIRE(ktot == p->ktot, "igemm: ktot mismatch (expected %d, got %d)", ktot, p->ktot);
  for (int i = 0; i < p->nsrc; ++i)
    LDM_REQUIRE(p->src[i] != nullptr && (reinterpret_cast<uintptr_t>(p->src[i]) & 15) == 0,
                "igemm: source %d null or not 16-byte aligned", i);
  LDM_REQUIRE(p->weight && (reinterpret_cast<uintptr_t>(p->weight) & 15) == 0,
              "igemm: weight null or misaligned");
  LDM_REQUIRE(p->out != nullptr, "igemm: null out");
  if (p->act == LDMSEG_ACT_GEGLU) {
    LDM_REQUIRE(p->n % 32 == 0 && p->out_dtype == LDMSEG_OUT_BF16 && p->residual == nullptr,
                "igemm: GEGLU needs n %% 32 == 0, bf16 out, no residual");
    LDM_REQUIRE(p->out_ld % 8 == 0, "igemm: GEGLU out_ld must be a multiple of 8");
  }
  if (p->out_dtype == LDMSEG_OUT_BF16 && p->act != LDMSEG_ACT_GEGLU)
    LDM_REQUIRE(p->out_ld % 4 == 0, "igemm: bf16 out_ld must be a multiple of 4");
  if (p->out_dtype == LDMSEG_OUT_F32)
    LDM_REQUIRE(p->out_ld % 4 == 0, "igemm: f32 out_ld must be a multiple of 4");
  if (p->residual) LDM_REQUIRE(p->res_ld % 4 == 0, "igemm: res_ld must be a multiple of 4");
  LDM_REQUIRE(p->n % 4 == 0, "igemm: n must be a multiple of 4 (got %d)", p->n);
  LDM_REQUIRE(p->conv_stride == 0 || p->conv_stride == 1 || p->conv_stride == 2, "igemm: conv_stride must be 1 or 2");
  if (p->conv_stride == 2) {
    LDM_REQUIRE(p->conv_pad == 0 || p->conv_pad == 1, "igemm: stride-2 conv_pad must be 0 or 1");
    for (int s = 0; s < p->nseg; ++s) LDM_REQUIRE(p->seg_taps[s] == 9, "igemm: stride 2 is defined for 3x3 segments only");
  }
  if (p->upsample2) {
    LDM_REQUIRE(p->nseg == 1 && p->seg_taps[0] == 4, "igemm: upsample2 takes one segment with 4 taps");
    LDM_REQUIRE(p->weight_tiled, "igemm: upsample2 needs block-tiled weights (4 phase matrices stacked along n)");
    LDM_REQUIRE(p->conv_stride != 2 && p->act != LDMSEG_ACT_GEGLU && !p->residual && !p->rowbias && !p->rowstats_out &&
                    !p->ln_colsum && !p->ln_rowstats,
                "igemm: upsample2 supports bias, SiLU, f32 / shadow outputs and fused statistics only");
    if (p->pair)
      LDM_REQUIRE(((static_cast<long long>(p->nb) * p->h * p->w + BM - 1) / BM) % 2 == 0,
                  "igemm: upsample2 in pair mode needs an even number of 128-row input tiles (a pair shares one phase)");
  }
  if (p->out2) {
    LDM_REQUIRE(p->out_dtype == LDMSEG_OUT_F32 && p->act != LDMSEG_ACT_GEGLU, "igemm: out2 (bf16 shadow) needs an f32 out");
    LDM_REQUIRE(p->out2_ld % 4 == 0, "igemm: out2_ld must be a multiple of 4");
  }
  if (p->rowstats_out) LDM_REQUIRE(p->act != LDMSEG_ACT_GEGLU, "igemm: rowstats_out is not defined for GEGLU");
  if (p->ln_colsum || p->ln_rowstats) {
    LDM_REQUIRE(p->ln_colsum && p->ln_rowstats && p->ln_channels > 0, "igemm: a folded LayerNorm needs ln_rowstats, ln_colsum and ln_channels");
    LDM_REQUIRE((reinterpret_cast<uintptr_t>(p->ln_colsum) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->ln_rowstats) & 7) == 0,
                "igemm: ln_colsum / ln_rowstats misaligned");
  }
  if (p->next_weight) LDM_REQUIRE((reinterpret_cast<uintptr_t>(p->next_weight) & 15) == 0, "igemm: next_weight misaligned");
  if (p->rowbias) LDM_REQUIRE(p->rowbias_ld % 4 == 0, "igemm: rowbias_ld must be a multiple of 4");
  if (p->stats) {
    LDM_REQUIRE((p->stats_hw > 0 ? p->stats_hw : hw) % 32 == 0, "igemm: fused statistics need rows per image %% 32 == 0");
    LDM_REQUIRE(p->act != LDMSEG_ACT_GEGLU, "igemm: fused statistics are not defined for GEGLU");
  }
  return 0;
}

template <int BN, bool GEGLU, bool SPLIT, bool PAIR, bool TAIL = false, bool CSPLIT = false>
static int launch_igemm_v(const IgemmKParams& kp, int grid, cudaStream_t stream, int pdl) {
  using Cfg = IgemmCfg<BN, PAIR>;
  auto kernel = igemm_kernel<BN, GEGLU, SPLIT, PAIR, TAIL, CSPLIT>;
  static bool configured = false;
  if (!configured) {
    LDM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (CSPLIT) LDM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(igemm_threads(GEGLU, SPLIT, BN));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  const bool dbg_cluster = !PAIR && !CSPLIT && SPLIT && (g_debug & 64) && grid % kp.split_k == 0 && kp.split_k <= 8;
  if (PAIR || CSPLIT || dbg_cluster) {   // the two CTAs of a pair must share a TPC; the splits of a tile form one cluster
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (CSPLIT || dbg_cluster) ? kp.split_k : 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl || g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (PAIR) {
    // The persistent pair grid must be co-resident (split-K CTAs wait for their peers): never launch more
    // clusters than the device can hold at once (a part with a half-populated TPC holds fewer than #SMs / 2).
    static int max_clusters = 0;
    if (max_clusters == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms() / 2;
      }
      max_clusters = n;
    }
    if (grid > 2 * max_clusters) {
      // the stream-K tail was planned for `grid` units: a smaller grid would change every unit's piece
      if (TAIL) {
        set_error("igemm: the device holds only %d CTA pairs at once (planned for %d)", max_clusters, grid / 2);
        return -3;
      }
      cfg.gridDim = dim3(2 * max_clusters);
    }
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, kp);
  if (e != cudaSuccess) {
    set_error("igemm_kernel launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return check_launch("igemm_kernel");
}

// How many clusters of `cluster_size` split CTAs the device holds at once (0: such a cluster cannot be formed).  A
// split launch is a single wave by construction, so its tile count must not exceed this.
template <int BN, bool GEGLU>
static int csplit_max_clusters_v(int cluster_size) {
  using Cfg = IgemmCfg<BN, false>;
  auto kernel = igemm_kernel<BN, GEGLU, true, false, false, true>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster_size * num_sms());
  cfg.blockDim = dim3(igemm_threads(GEGLU, true));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
static int csplit_max_clusters(int bn, bool geglu, int cluster_size) {
  if (cluster_size < 2 || cluster_size > 16) return 0;
  static int cache[4][2][17];
  static bool have[4][2][17];
  const int bi = bn == 64 ? 0 : bn == 128 ? 1 : bn == 160 ? 2 : 3;
  if (!have[bi][geglu][cluster_size]) {
    int n = 0;
    switch (bn) {
      case 64: n = geglu ? csplit_max_clusters_v<64, true>(cluster_size) : csplit_max_clusters_v<64, false>(cluster_size); break;
      case 128: n = geglu ? csplit_max_clusters_v<128, true>(cluster_size) : csplit_max_clusters_v<128, false>(cluster_size); break;
      case 160: n = geglu ? csplit_max_clusters_v<160, true>(cluster_size) : csplit_max_clusters_v<160, false>(cluster_size); break;
      default: n = geglu ? csplit_max_clusters_v<256, true>(cluster_size) : csplit_max_clusters_v<256, false>(cluster_size); break;
    }
    cache[bi][geglu][cluster_size] = n;
    have[bi][geglu][cluster_size] = true;
  }
  return cache[bi][geglu][cluster_size];
}

template <int BN, bool PAIR>
static int launch_igemm(const IgemmKParams& kp, int grid, cudaStream_t stream, int pdl) {
  const bool geglu = kp.act == LDMSEG_ACT_GEGLU, split = kp.split_k > 1;
  if constexpr (BN == 320) {   // whole tiles or the stream-K tail only (validated by the caller)
    (void)geglu;
    (void)split;
    return kp.tail ? launch_igemm_v<BN, false, false, true, true>(kp, grid, stream, pdl)
                   : launch_igemm_v<BN, false, false, true, false>(kp, grid, stream, pdl);
  } else {
    if (kp.tail) return launch_igemm_v<BN, false, false, PAIR, true>(kp, grid, stream, pdl);
    if constexpr (!PAIR) {
      if (kp.csplit)
        return geglu ? launch_igemm_v<BN, true, true, false, false, true>(kp, grid, stream, pdl)
                     : launch_igemm_v<BN, false, true, false, false, true>(kp, grid, stream, pdl);
    }
    if (geglu) return split ? launch_igemm_v<BN, true, true, PAIR>(kp, grid, stream, pdl)
                            : launch_igemm_v<BN, true, false, PAIR>(kp, grid, stream, pdl);
    return split ? launch_igemm_v<BN, false, true, PAIR>(kp, grid, stream, pdl)
                 : launch_igemm_v<BN, false, false, PAIR>(kp, grid, stream, pdl);
  }
}

static int choose_block_n(int m_tiles, int n, int sms) {
  // pick the tile width with the lowest (waves x per-tile MMA time) estimate; ties -> wider
  const int cands[4] = {256, 160, 128, 64};
  int best = 128;
  double best_cost = 1e30;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const int n_tiles = (n + bn - 1) / bn;
    const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
    const long long waves = (tiles + sms - 1) / sms;
    // narrow tiles are smem-bandwidth-bound: never cheaper than a ~96-wide tile
    const double tile_cost = bn < 96 ? 96.0 : static_cast<double>(bn);
    const double cost = static_cast<double>(waves) * tile_cost + 8.0;  // + fixed per-wave overhead
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = bn;
    }
  }
  return best;
}

}  // namespace ldm

using namespace ldm;

extern "C" int ldmseg_igemm(const ldmseg_igemm_params* p, void* stream) {
  if (int rc = validate(p)) return rc;
  IgemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const int M = p->nb * p->h * p->w;
  const int HW = p->h * p->w;
  // A box: 128 consecutive pixels of the flattened (n, y, x) order
  uint32_t bw = p->w >= BM ? BM : p->w;
  uint32_t bh = (p->w >= BM) ? 1 : (HW >= BM ? BM / p->w : p->h);
  uint32_t bb = BM / (bw * bh);
  // stride-2 convs read the 2h x 2w input through the TMA traversal stride: a box of (2 bw) x (2 bh) traversed
  // elements lands as bw x bh pixels in shared memory -- no im2col buffer
  const uint32_t cs = p->conv_stride == 2 ? 2 : 1;
  for (int i = 0; i < p->nsrc; ++i) {
    const uint64_t C = p->src_c[i];
    const uint64_t win = static_cast<uint64_t>(p->w) * cs, hin = static_cast<uint64_t>(p->h) * cs;
    uint64_t dims[4] = {C, win, hin, static_cast<uint64_t>(p->nb)};
    uint64_t strides[3] = {C * 2, C * 2 * win, C * 2 * win * hin};
    uint32_t box[4] = {BK, bw * cs, bh * cs, bb};
    uint32_t estr[4] = {1, cs, cs, 1};
    if (int rc = encode_tmap_bf16(&kp.a_map[i], p->src[i], 4, dims, strides, box, cs == 2 ? estr : nullptr)) return rc;
  }
  kp.a_stride = static_cast<int>(cs);
  kp.a_pad = cs == 2 ? p->conv_pad : 1;
  // folded x2 up-sampling: 4 phase GEMMs over the input pixels, laid out as 4 x as many (phase-major) 128-row tiles
  const int up2 = p->upsample2 ? 1 : 0;
  const int m_tiles = (up2 ? 4 : 1) * ((M + BM - 1) / BM);
  kp.up2 = up2;
  kp.up_m_tiles = (M + BM - 1) / BM;
  kp.up_nblk = up2 ? (p->n + 15) / 16 : 0;
  int bn = p->block_n;
  if (bn == 0) bn = choose_block_n(m_tiles, p->n, num_sms());
  LDM_REQUIRE(bn == 64 || bn == 128 || bn == 160 || bn == 256 || bn == 320, "igemm: unsupported block_n %d", bn);
  const bool pair = p->pair != 0;
  if (pair) {
    LDM_REQUIRE(p->weight_tiled, "igemm: pair mode needs block-tiled weights");
    LDM_REQUIRE(bn == 128 || bn == 160 || bn == 256 || bn == 320,
                "igemm: pair mode supports block_n 128 / 160 / 256 / 320 (got %d)", bn);
    LDM_REQUIRE(m_tiles >= 2, "igemm: pair mode needs at least two 128-row tiles");
  }
  if (bn == 320)
    LDM_REQUIRE(pair && p->split_k <= 1 && p->act != LDMSEG_ACT_GEGLU && !p->split_cluster,
                "igemm: block_n 320 needs pair mode, no split_k, no GEGLU");
  {
    if (p->weight_tiled) {
      // [N/16][K/64][16][64]: every 16-row x 64-k block is 2 KB contiguous -> weight streaming reads
      // whole DRAM pages instead of 128-byte pieces at a K-row stride; a CTA of a pair stages half a tile
      // (bn / 2 rows: 80 for bn = 160, hence 16-row blocks)
      const uint64_t kblocks = static_cast<uint64_t>(p->ktot) / BK;
      uint64_t dims[4] = {BK, 16, kblocks, static_cast<uint64_t>((up2 ? 4 : 1) * ((p->n + 15) / 16))};
      uint64_t strides[3] = {128, 2048, kblocks * 2048};
      // (block_n 320: two N = 160 instructions per k-step, one box of 80 rows per CTA and instruction)
      uint32_t box[4] = {BK, 16, 1, static_cast<uint32_t>(pair ? (bn == 320 ? 160 : bn) / 32 : bn / 16)};
      if (int rc = encode_tmap_bf16(&kp.b_map, p->weight, 4, dims, strides, box)) return rc;
    } else {
    uint64_t dims[2] = {static_cast<uint64_t>(p->ktot), static_cast<uint64_t>(p->n)};
    uint64_t strides[1] = {static_cast<uint64_t>(p->ktot) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn)};
    if (int rc = encode_tmap_bf16(&kp.b_map, p->weight, 2, dims, strides, box)) return rc;
    }
  }
  kp.nseg = p->nseg;
  int num_kb = 0;
  for (int s = 0; s < p->nseg; ++s) {
    kp.seg_src[s] = p->seg_src[s];
    kp.seg_taps[s] = p->seg_taps[s];
    kp.seg_cblocks[s] = (p->src_c[p->seg_src[s]] + BK - 1) / BK;
    num_kb += kp.seg_taps[s] * kp.seg_cblocks[s];
  }
  kp.M = M;
  kp.N = p->n;
  kp.H = p->h;
  kp.W = p->w;
  kp.HW = HW;
  kp.num_m_tiles = m_tiles;
  kp.num_n_tiles = (p->n + bn - 1) / bn;
  kp.num_kb = num_kb;
  kp.bias = p->bias;
  kp.rowbias = p->rowbias;
  kp.rowbias_ld = p->rowbias_ld;
  kp.residual = p->residual;
  kp.res_ld = p->res_ld;
  kp.res_f32 = p->residual_f32 ? 1 : 0;
  kp.out2 = reinterpret_cast<__nv_bfloat16*>(p->out2);
  kp.out2_ld = p->out2_ld;
  {
    static int mode = -1;  // LDMSEG_NEXTW=0 disables the next-launch weight prefetch (A/B timing)
    if (mode < 0) {
      const char* e = getenv("LDMSEG_NEXTW");
      mode = e ? atoi(e) : 1;
    }
    if (mode != 0 && p->next_weight != nullptr && p->next_weight_bytes >= 16) {
      kp.next_w = reinterpret_cast<const uint8_t*>(p->next_weight);
      kp.next_w_bytes = static_cast<unsigned long long>(p->next_weight_bytes);
    }
  }
  kp.out = p->out;
  kp.out_ld = p->out_ld;
  kp.out_f32 = p->out_dtype == LDMSEG_OUT_F32;
  kp.act = p->act;
  kp.split_k = p->split_k > 1 ? p->split_k : 1;
  if (kp.split_k > num_kb) kp.split_k = num_kb;
  if (kp.split_k > 1) {
    LDM_REQUIRE(p->workspace != nullptr && p->tile_counters != nullptr,
                "igemm: split_k > 1 needs workspace and tile_counters");
    const int ws_m_tiles = pair ? (kp.num_m_tiles + 1) / 2 * 2 : kp.num_m_tiles;   // incl. the phantom tile
    const long long need = static_cast<long long>(ws_m_tiles) * kp.num_n_tiles * kp.split_k * BM * bn;
    LDM_REQUIRE(p->workspace_elems >= need, "igemm: split-K workspace too small (%lld f32 needed, %lld given)",
                need, static_cast<long long>(p->workspace_elems));
    LDM_REQUIRE(static_cast<long long>(ws_m_tiles) * kp.num_n_tiles <= 4096,
                "igemm: split-K supports at most 4096 output tiles (tile_counters holds 2 x 4096 int32)");
    kp.workspace = p->workspace;
    kp.counters = p->tile_counters;
  }
  if (p->split_cluster && kp.split_k > 1 && !pair) {
    // the splits of a tile as one cluster: only if every tile's cluster fits on the device at once
    const long long tiles = static_cast<long long>(kp.num_m_tiles) * kp.num_n_tiles;
    // the owner's slots ([own units][splits] x 4 KB) and 32 KB of staging must fit in the idle operand ring
    const int units = 4 * bn / 32, n_own = (units + kp.split_k - 1) / kp.split_k;
    const long long ring = bn == 64 ? 9LL * (kABytes + 64 * 128) : bn == 128 ? 7LL * (kABytes + 128 * 128)
                           : bn == 160 ? 6LL * (kABytes + 160 * 128) : 4LL * (kABytes + 256 * 128);
    if (static_cast<long long>(n_own) * kp.split_k * 4096 + 32768 <= ring &&
        tiles <= csplit_max_clusters(bn, p->act == LDMSEG_ACT_GEGLU, kp.split_k))
      kp.csplit = 1;
  }
  if (p->stream_k && kp.split_k == 1 && p->act != LDMSEG_ACT_GEGLU && bn != 64) {
    // stream-K tail (see TailSeg): only when the last wave is ragged and its pieces are non-empty; otherwise the
    // launch silently runs as whole tiles
    const int units = pair ? num_sms() / 2 : num_sms();
    const long long tiles = static_cast<long long>(pair ? (kp.num_m_tiles + 1) / 2 : kp.num_m_tiles) * kp.num_n_tiles;
    const long long full = tiles / units * units, rem = tiles - full;
    const long long steps = rem * num_kb;
    if (rem > 0 && steps >= units) {
      const long long per = steps / units;                      // shortest piece
      const long long slots = num_kb / per + 2;                 // pieces that can touch one tile, minus the finisher
      const long long need = rem * (pair ? 2 : 1) * slots * BM * bn;
      if (p->workspace != nullptr && p->tile_counters != nullptr && p->workspace_elems >= need &&
          rem * (pair ? 2 : 1) <= 4096) {
        kp.tail = 1;
        kp.tail_full_tiles = static_cast<int>(full);
        kp.tail_tiles = static_cast<int>(rem);
        kp.tail_slots = static_cast<int>(slots);
        kp.workspace = p->workspace;
        kp.counters = p->tile_counters;
      }
    }
  }
  kp.stats = p->stats;
  kp.rowstats = p->rowstats_out;
  kp.ln_rowstats = p->ln_rowstats;
  kp.ln_colsum = p->ln_colsum;
  kp.ln_inv_c = p->ln_channels > 0 ? 1.f / static_cast<float>(p->ln_channels) : 0.f;
  kp.ln_eps = p->ln_eps;
  kp.debug = g_debug;
  kp.w_tiled = p->weight_tiled;
  {
    static int coop = -1;  // LDMSEG_SPLITK_COOP=0: one warp per owned chunk sums all partials (A/B timing)
    if (coop < 0) {
      const char* e = getenv("LDMSEG_SPLITK_COOP");
      coop = e ? atoi(e) : 1;
    }
    kp.coop_reduce = coop != 0;
  }
  {
    static int mode = -1;  // LDMSEG_IGEMM_PREFETCH=0 disables (A/B timing)
    if (mode < 0) {
      const char* e = getenv("LDMSEG_IGEMM_PREFETCH");
      mode = e ? atoi(e) : 1;
    }
    // only for operands that nothing on the stream writes (real weights): a B operand produced by an earlier
    // launch (the VAE attention's K / V^T) must not be read before the grid-dependency wait
    kp.prefetch_b = mode != 0 && p->weight_static != 0 && !(g_debug & 24);
  }
  {
    const size_t esz = kp.out_f32 ? 4 : 2;
    bool ok = (reinterpret_cast<uintptr_t>(p->out) & 31) == 0 && (static_cast<size_t>(p->out_ld) * esz) % 32 == 0;
    if (p->residual)
      ok = ok && (reinterpret_cast<uintptr_t>(p->residual) & 31) == 0 &&
           (static_cast<size_t>(p->res_ld) * (p->residual_f32 ? 4 : 2)) % 32 == 0;
    if (p->out2)
      ok = ok && (reinterpret_cast<uintptr_t>(p->out2) & 31) == 0 && (static_cast<size_t>(p->out2_ld) * 2) % 32 == 0;
    if (p->bias) ok = ok && (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0;
    if (p->rowbias) ok = ok && (reinterpret_cast<uintptr_t>(p->rowbias) & 15) == 0 && p->rowbias_ld % 4 == 0;
    kp.vec_ok = ok ? 1 : 0;
    if (p->act == LDMSEG_ACT_GEGLU)
      LDM_REQUIRE(ok, "igemm: GEGLU needs 32-byte aligned out (ld %% 16 == 0) and 16-byte aligned biases");
  }
  kp.stats_hw = p->stats_hw > 0 ? p->stats_hw : HW;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (pair) {
    // one CTA pair per work item, at most one pair per TPC
    const long long work = static_cast<long long>((kp.num_m_tiles + 1) / 2) * kp.num_n_tiles * kp.split_k;
    const int pairs = num_sms() / 2;
    const int grid = 2 * static_cast<int>((work < pairs && !kp.tail) ? work : pairs);
    switch (bn) {
      case 128: return launch_igemm<128, true>(kp, grid, st, p->pdl);
      case 160: return launch_igemm<160, true>(kp, grid, st, p->pdl);
      case 320: return launch_igemm<320, true>(kp, grid, st, p->pdl);
      default: return launch_igemm<256, true>(kp, grid, st, p->pdl);
    }
  }
  const long long work = static_cast<long long>(kp.num_m_tiles) * kp.num_n_tiles * kp.split_k;
  const int grid = static_cast<int>(((work < num_sms() && !kp.tail) || kp.csplit) ? work : num_sms());
  switch (bn) {
    case 64: return launch_igemm<64, false>(kp, grid, st, p->pdl);
    case 128: return launch_igemm<128, false>(kp, grid, st, p->pdl);
    case 160: return launch_igemm<160, false>(kp, grid, st, p->pdl);
    default: return launch_igemm<256, false>(kp, grid, st, p->pdl);
  }
}

extern "C" int ldmseg_igemm_max_split_clusters(int block_n, int geglu, int cluster_size) {
  if (block_n != 64 && block_n != 128 && block_n != 160 && block_n != 256) return 0;
  return csplit_max_clusters(block_n, geglu != 0, cluster_size);
}

extern "C" int ldmseg_igemm_simple(const ldmseg_igemm_params* p, void* stream) {
  if (int rc = validate(p)) return rc;
  LDM_REQUIRE(!p->upsample2, "igemm_simple: upsample2 is not defined for the reference-grade kernel");
  const int M = p->nb * p->h * p->w;
  const long long total = static_cast<long long>(M) * p->n;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  launch_kernel(igemm_simple_kernel, dim3(static_cast<unsigned>(blocks)), dim3(threads), 0,
                reinterpret_cast<cudaStream_t>(stream), *p, M, p->h * p->w);
  return check_launch("igemm_simple_kernel");
}

extern "C" int ldmseg_set_debug(int flags) {
  const int old = ldm::g_debug;
  ldm::g_debug = flags;
  return old;
}
