// Implicit GEMM on the 5th-gen tensor cores (tcgen05 + TMEM), fed by TMA.
//
// One kernel covers every contraction of the UNet / VAE hot path that is not attention:
// conv3x3 (9 shifted TMA box loads of the channel-last activation, zero padding by TMA
// out-of-bounds fill), conv1x1 / linear (1 tap), concat-free skip connections and the fused 1x1
// shortcut (several K segments accumulating into one TMEM tile).
//
// Replaces F.conv2d / F.linear inside diffusers' ResnetBlock2D / Transformer2DModel / Attention /
// FeedForward / Downsample2D / Upsample2D as called from /root/reference/ldmseg/models/unet.py:357,
// 361-373, 388-395, 401-425, 431, and the convs of GeneralVAESeg.decode (models/vae.py:133-172).
//
// CTA = 192 threads, persistent over (tile, k-split) work items:
//   warp 0      TMA producer   (one elected lane)
//   warp 1      TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5  epilogue: tcgen05.ld -> bias / row-bias / residual / SiLU / GEGLU -> global
// Pipelines: smem ring (full/empty mbarriers) between TMA and MMA; two TMEM accumulator stages
// (tmem_full/tmem_empty) between MMA and epilogue so tile i's epilogue overlaps tile i+1's MMAs.
//
// Tile: BLOCK_M = 128 output pixels x BN output channels, BLOCK_K = 64 (one 128-byte swizzle row).
#include "common.h"
#include <cstring>
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kIgemmThreads = 64 + 256;  // TMA warp, MMA warp, 8 epilogue warps
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
  const __nv_bfloat16* residual;
  int res_ld;
  void* out;
  int out_ld;
  int out_f32;
  int act;
  int split_k;
  float* workspace;   // split-K partial tiles: [tile][split][128][BN] f32
  int* counters;
  float* stats;       // optional per-(image, channel) {sum, sum of squares} of the stored output
  int stats_hw;       // rows per image for the statistics (the producer may be a plain [M, K] GEMM)
  int w_tiled;        // weights stored as [N/32][K/64][32][64] blocks (4 KB contiguous per block)
  int debug;          // development only: bit0 skip final reduce, bit1 skip sync, bit2 skip partial store
};

template <int BN>
struct IgemmCfg {
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN <= 64) ? 8 : (BN <= 128) ? 6 : (BN <= 160) ? 5 : 4;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kStagingBytes = 8 * 32 * 64;   // per epilogue warp: 32 rows x 16 f32
  static constexpr int kStatBytes = 2 * BN * 2 * 4;    // fused GroupNorm statistics: [2 slots][BN][2] f32
  // stages + epilogue staging + barriers (2*stages + 4) * 8 + tmem ptr + split-K flag, + 1024 slack
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStagingBytes + kStatBytes + (2 * kStages + 4) * 8 + 16 + 1024;
};

// ---- epilogue ------------------------------------------------------------------------------
// Eight epilogue warps (two per TMEM lane quadrant, splitting the tile's 32-column super-chunks
// between them) so that two warps per SM sub-partition hide each other's latencies.
// tcgen05.ld hands every thread ONE accumulator row; storing that directly touches 32 different
// lines per instruction.  Each 16-column half-chunk is therefore transposed through a 2 KB per-warp
// staging buffer (rotated 16-byte slots, conflict-free both ways): afterwards lane l holds columns
// 4*(l&3)..+3 of row 8*it + (l>>2), it = 0..3, so 4 lanes cover 64 contiguous bytes of a row and all
// global accesses (output, residual, split-K partials) move whole sectors.
enum { EPI_DIRECT = 0, EPI_PARTIAL = 1, EPI_FINAL = 2 };
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float fast_silu(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// GELU(x) = x/2 (1 + erf(x/sqrt2)), erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// bf16 rounding of the stored product); 2 MUFU + ~12 FMA instead of the branchy erff
__device__ __forceinline__ float fast_gelu(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(1.f + 0.3275911f * z);
  float poly = 1.061405429f;
  poly = poly * t - 1.453152027f;
  poly = poly * t + 1.421413741f;
  poly = poly * t - 0.284496736f;
  poly = poly * t + 0.254829592f;
  const float erf_abs = 1.f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// Epilogue arguments held in registers (reading them through the parameter block from inside the loops
// made every access a load the compiler had to repeat after each global store).
struct EpiArgs {
  int M, N, HW, out_ld, res_ld, rowbias_ld, act, out_f32, split_k, stats_hw;
  const float* bias;
  const float* rowbias;
  const __nv_bfloat16* residual;
  void* out;
  float* stats;
};

// One warp, one output tile: processes the super-chunks sc = half, half+2, ... (32 columns each, as two
// 16-column halves).  `sstat` = per-CTA shared accumulators [2 image slots][BN][2] for the fused
// GroupNorm statistics.
template <int BN>
__device__ __forceinline__ void epilogue_warp(const EpiArgs p, int mode, uint32_t stg, float* sstat,
                                              uint32_t t_row, float* ws_tile, int split_idx, int m_base,
                                              int n0, int half, int lane) {
  const int jc = lane & 3;     // 16-byte column slot inside the 16-column half-chunk
  const int rsub = lane >> 2;   // row within a group of 8
  constexpr int kTileElems = BM * BN;
  const int m_tile0 = m_base & ~(BM - 1);
  const bool want_stats = p.stats != nullptr && mode != EPI_PARTIAL;
  // cooperative split-K reduction: the tile's 16 row groups (8 rows each; this warp's quadrant owns
  // groups 4q..4q+3) are dealt round-robin to the split CTAs; bit `it` of `mine` = reduce group 4q+it
  uint32_t mine = 0xf;
  if (mode == EPI_FINAL) {
    mine = 0;
    const int q4 = ((m_base >> 5) & 3) * 4;
#pragma unroll
    for (int it = 0; it < 4; ++it) mine |= (((q4 + it) % p.split_k) == split_idx ? 1u : 0u) << it;
  }
  const bool warp_one_image = (p.HW & 31) == 0;
  const int stat_slot = want_stats ? (m_base / p.stats_hw - m_tile0 / p.stats_hw) : 0;
#pragma unroll 1
  for (int sc = half; sc * 32 < BN; sc += 2) {
    if (n0 + sc * 32 >= p.N) break;
    float4 hv[4];  // GEGLU: the h half of the super-chunk
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int c16 = sc * 32 + hh * 16;
      const int col = n0 + c16 + jc * 4;
      const bool col_ok = col < p.N;
      // ---- early, latency-tolerant loads (overlap the TMEM load + staging round trip)
      uint2 resv[4];
      float4 add4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mode != EPI_PARTIAL) {
        if (p.bias != nullptr && col_ok) add4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        if (p.rowbias != nullptr && col_ok && warp_one_image && m_base < p.M)
          add4 = f4_add(add4, __ldg(reinterpret_cast<const float4*>(
                                  p.rowbias + static_cast<size_t>(m_base / p.HW) * p.rowbias_ld + col)));
        if (p.residual != nullptr) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int m = m_base + it * 8 + rsub;
            resv[it] = make_uint2(0u, 0u);
            if (col_ok && m < p.M && ((mine >> it) & 1))
              resv[it] = __ldg(reinterpret_cast<const uint2*>(p.residual + static_cast<size_t>(m) * p.res_ld + col));
          }
        }
      }
      // ---- accumulator values of this lane's 4 (row, 4-column) slots
      if (mode != EPI_FINAL) {
        uint32_t r[16];
        tmem_ld_32x16(t_row + c16, r);
        tmem_wait_ld();
        const uint32_t rowp = stg + lane * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts128(rowp + (((j + (lane >> 1)) & 3) << 4),
                 make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                             __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        __syncwarp();
      }
      float4 v[4];
      if (mode == EPI_FINAL) {
        // sum the split-K partials with many independent loads in flight
#pragma unroll
        for (int it = 0; it < 4; ++it) v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int s0 = 0; s0 < p.split_k; s0 += 2) {
          float4 t[4][2];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int m = m_base + it * 8 + rsub;
            const bool ok = col_ok && m < p.M && ((mine >> it) & 1);
            const float* src = ws_tile + static_cast<size_t>(m - m_tile0) * BN + c16 + jc * 4;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
              t[it][d] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok && s0 + d < p.split_k)
                t[it][d] = __ldcg(reinterpret_cast<const float4*>(src + static_cast<size_t>(s0 + d) * kTileElems));
            }
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) v[it] = f4_add(v[it], f4_add(t[it][0], t[it][1]));
        }
      } else {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = it * 8 + rsub;
          v[it] = lds128(stg + row * 64 + (((jc + (row >> 1)) & 3) << 4));
        }
        __syncwarp();  // staging may be overwritten by the next half-chunk
      }
      // ---- per-slot epilogue
      float4 s_sum = make_float4(0.f, 0.f, 0.f, 0.f), s_sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int m = m_base + it * 8 + rsub;
        const bool ok = col_ok && m < p.M && ((mine >> it) & 1);
        float4 x = v[it];
        if (mode == EPI_PARTIAL) {
          if (ok) *reinterpret_cast<float4*>(ws_tile + static_cast<size_t>(split_idx) * kTileElems +
                                             static_cast<size_t>(m - m_tile0) * BN + c16 + jc * 4) = x;
          continue;
        }
        x = f4_add(x, add4);
        if (p.rowbias != nullptr && !warp_one_image && ok)
          x = f4_add(x, __ldg(reinterpret_cast<const float4*>(
                            p.rowbias + static_cast<size_t>(m / p.HW) * p.rowbias_ld + col)));
        if (p.act == LDMSEG_ACT_GEGLU) {
          // super-chunk columns are [16 x h | 16 x g]
          if (hh == 0) {
            hv[it] = x;
          } else if (ok) {
            uint2 u;
            u.x = pack_bf16x2(hv[it].x * fast_gelu(x.x), hv[it].y * fast_gelu(x.y));
            u.y = pack_bf16x2(hv[it].z * fast_gelu(x.z), hv[it].w * fast_gelu(x.w));
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(m) * p.out_ld +
                                      ((n0 + sc * 32) >> 1) + jc * 4) = u;
          }
          continue;
        }
        if (p.residual != nullptr) {
          const float2 a = unpack_bf16x2(resv[it].x), b = unpack_bf16x2(resv[it].y);
          x.x += a.x; x.y += a.y; x.z += b.x; x.w += b.y;
        }
        if (p.act == LDMSEG_ACT_SILU) {
          x.x = fast_silu(x.x); x.y = fast_silu(x.y); x.z = fast_silu(x.z); x.w = fast_silu(x.w);
        }
        if (p.out_f32) {
          if (ok) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.out_ld + col) = x;
        } else {
          uint2 u;
          u.x = pack_bf16x2(x.x, x.y);
          u.y = pack_bf16x2(x.z, x.w);
          if (ok) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(m) * p.out_ld + col) = u;
          if (want_stats) {  // statistics of what the consumer will read (bf16-rounded)
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
            x = make_float4(a.x, a.y, b.x, b.y);
          }
        }
        if (want_stats && ok) {
          s_sum = f4_add(s_sum, x);
          s_sq = f4_add(s_sq, make_float4(x.x * x.x, x.y * x.y, x.z * x.z, x.w * x.w));
        }
      }
      if (want_stats) {
        // column sums over this warp's rows (lanes with equal jc), then into the CTA's shared accumulators
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) {
          s_sum.x += __shfl_xor_sync(0xffffffffu, s_sum.x, o);
          s_sum.y += __shfl_xor_sync(0xffffffffu, s_sum.y, o);
          s_sum.z += __shfl_xor_sync(0xffffffffu, s_sum.z, o);
          s_sum.w += __shfl_xor_sync(0xffffffffu, s_sum.w, o);
          s_sq.x += __shfl_xor_sync(0xffffffffu, s_sq.x, o);
          s_sq.y += __shfl_xor_sync(0xffffffffu, s_sq.y, o);
          s_sq.z += __shfl_xor_sync(0xffffffffu, s_sq.z, o);
          s_sq.w += __shfl_xor_sync(0xffffffffu, s_sq.w, o);
        }
        if (rsub == 0 && col_ok) {
          float* st = sstat + (static_cast<size_t>(stat_slot) * BN + c16 + jc * 4) * 2;
          atomicAdd(st + 0, s_sum.x); atomicAdd(st + 1, s_sq.x);
          atomicAdd(st + 2, s_sum.y); atomicAdd(st + 3, s_sq.y);
          atomicAdd(st + 4, s_sum.z); atomicAdd(st + 5, s_sq.z);
          atomicAdd(st + 6, s_sum.w); atomicAdd(st + 7, s_sq.w);
        }
      }
    }
  }
}

// After every epilogue warp of the CTA has finished a tile: one global atomic per (image slot, column).
template <int BN>
__device__ __forceinline__ void flush_stats(const EpiArgs p, float* sstat, int m_tile0, int n0, int et) {
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const int img0 = m_tile0 / p.stats_hw;
  for (int idx = et; idx < 2 * BN; idx += kEpiThreads) {
    const int slot = idx / BN, colr = idx - slot * BN;
    const float su = sstat[idx * 2], sq = sstat[idx * 2 + 1];
    if (n0 + colr < p.N && (su != 0.f || sq != 0.f)) {
      float* g = p.stats + (static_cast<size_t>(img0 + slot) * p.N + n0 + colr) * 2;
      atomicAdd(g, su);
      atomicAdd(g + 1, sq);
    }
    sstat[idx * 2] = 0.f;
    sstat[idx * 2 + 1] = 0.f;
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

template <int BN>
__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_kernel(const __grid_constant__ IgemmKParams p) {
  using Cfg = IgemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint8_t* smem_stg = smem + kStages * Cfg::kStageBytes;
  float* sstat = reinterpret_cast<float*>(smem_stg + Cfg::kStagingBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes + Cfg::kStatBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) tma_prefetch_desc(&p.a_map[p.seg_src[s]]);
    tma_prefetch_desc(&p.b_map);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < Cfg::kStatBytes / 4; i += kIgemmThreads) sstat[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here
  // on we touch memory it produced.  (No-ops when the launch carries no PDL attribute.)
  pdl_sync();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int total_work = num_tiles * p.split_k;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x) {
        const int tile = wi / p.split_k;
        const int split = wi - tile * p.split_k;
        const int m_tile = tile / p.num_n_tiles;
        const int n_tile = tile - m_tile * p.num_n_tiles;
        const int m0 = m_tile * BM;
        const int x0 = m0 % p.W;
        const int y0 = (m0 / p.W) % p.H;
        const int b0 = m0 / p.HW;
        const int kb_begin = static_cast<int>(static_cast<long long>(split) * p.num_kb / p.split_k);
        const int kb_end =
            static_cast<int>(static_cast<long long>(split + 1) * p.num_kb / p.split_k);
        // decode kb_begin -> (segment, tap, channel block)
        int seg = 0, rem = kb_begin;
        while (seg < p.nseg - 1 && rem >= p.seg_taps[seg] * p.seg_cblocks[seg]) {
          rem -= p.seg_taps[seg] * p.seg_cblocks[seg];
          ++seg;
        }
        int tap = rem / p.seg_cblocks[seg];
        int cb = rem - tap * p.seg_cblocks[seg];
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          int dx = 0, dy = 0;
          if (p.seg_taps[seg] == 9) {
            dy = tap / 3 - 1;
            dx = tap - (tap / 3) * 3 - 1;
          }
          tma_load_4d(smem_a + stage * kABytes, &p.a_map[p.seg_src[seg]], &full_bar[stage],
                      cb * BK, x0 + dx, y0 + dy, b0);
          if (p.w_tiled)
            tma_load_4d(smem_b + stage * Cfg::kBBytes, &p.b_map, &full_bar[stage], 0, 0, kb,
                        n_tile * (BN / 32));
          else
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &p.b_map, &full_bar[stage], kb * BK,
                        n_tile * BN);
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
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x, ++it) {
        const int tile = wi / p.split_k;
        const int split = wi - tile * p.split_k;
        const int kb_begin = static_cast<int>(static_cast<long long>(split) * p.num_kb / p.split_k);
        const int kb_end =
            static_cast<int>(static_cast<long long>(split + 1) * p.num_kb / p.split_k);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc =
              make_smem_desc_sw128(smem_u32(smem_a + stage * kABytes), 16, 1024);
          const uint64_t bdesc =
              make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::kBBytes), 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the >>4 field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;             // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;    // which super-chunks of the tile this warp takes
    const int et = threadIdx.x - 64;     // 0..255
    const uint32_t stg = smem_u32(smem_stg + (warp - 2) * (32 * 64));
    EpiArgs ea;
    ea.M = p.M; ea.N = p.N; ea.HW = p.HW; ea.out_ld = p.out_ld; ea.res_ld = p.res_ld;
    ea.rowbias_ld = p.rowbias_ld; ea.act = p.act; ea.out_f32 = p.out_f32; ea.split_k = p.split_k;
    ea.stats_hw = p.stats_hw; ea.bias = p.bias; ea.rowbias = p.rowbias; ea.residual = p.residual;
    ea.out = p.out; ea.stats = p.stats;
    int it = 0;
    for (int wi = blockIdx.x; wi < total_work; wi += gridDim.x, ++it) {
      const int tile = wi / p.split_k;
      const int split = wi - tile * p.split_k;
      const int m_tile = tile / p.num_n_tiles;
      const int n_tile = tile - m_tile * p.num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_base = m_tile * BM + q * 32;
      const int n0 = n_tile * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (p.split_k <= 1) {
        epilogue_warp<BN>(ea, EPI_DIRECT, stg, sstat, t_row, nullptr, 0, m_base, n0, half, lane);
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        if (ea.stats != nullptr) flush_stats<BN>(ea, sstat, m_tile * BM, n0, et);
      } else {
        // split-K: every split stores its partial tile (coalesced, no atomics); once all splits of the
        // tile have arrived, each split CTA reduces and finishes its share of the tile's rows.
        float* ws_tile = p.workspace + static_cast<size_t>(tile) * p.split_k * (BM * BN);
        if (!(p.debug & 4))
          epilogue_warp<BN>(ea, EPI_PARTIAL, stg, sstat, t_row, ws_tile, split, m_base, n0, half, lane);
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        // publish + wait for the peers: the CTA barrier orders every thread's partial stores before
        // thread 0's release-atomic (release is cumulative); thread 0 then polls with acquire loads
        // until all splits of this tile have arrived.  Peers are CTAs of the same persistent grid in
        // the same round, so they are running or about to be scheduled (1 CTA per SM, grid <= #SMs).
        if (p.debug & 2) continue;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          int seen;
          asm volatile("atom.release.gpu.global.add.s32 %0, [%1], 1;" : "=r"(seen) : "l"(p.counters + tile) : "memory");
          uint32_t spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.counters + tile) : "memory");
            if (++spins > (1u << 26)) __trap();
          } while (seen < p.split_k);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (!(p.debug & 1))
          epilogue_warp<BN>(ea, EPI_FINAL, stg, sstat, 0, ws_tile, split, m_base, n0, half, lane);
        if (ea.stats != nullptr) flush_stats<BN>(ea, sstat, m_tile * BM, n0, et);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (et == 0) {
          // the last CTA to finish its share re-arms both counters for the next launch
          int* done = p.counters + 4096 + tile;
          int old;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(done) : "memory");
          if (old == p.split_k - 1) {
            p.counters[tile] = 0;
            *done = 0;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
  const __nv_bfloat16* wbase = reinterpret_cast<const __nv_bfloat16*>(p.weight);
  const int kblocks = p.ktot / 64;
  auto wat = [&](int k) -> float {
    const size_t off = p.weight_tiled
                           ? ((static_cast<size_t>(n >> 5) * kblocks + (k >> 6)) * 32 + (n & 31)) * 64 + (k & 63)
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
        dy = tap / 3 - 1;
        dx = tap % 3 - 1;
      }
      const int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
        const __nv_bfloat16* arow =
            a + (static_cast<size_t>(b) * HW + static_cast<size_t>(yy) * p.w + xx) * C;
        for (int c = 0; c < C; ++c)
          acc += __bfloat162float(arow[c]) * wat(koff + c);
      }
      koff += Cpad;
    }
  }
  // epilogue identical to the tcgen05 kernel (GEGLU handled by the pair thread layout below)
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
    acc += __bfloat162float(
        reinterpret_cast<const __nv_bfloat16*>(p.residual)[static_cast<size_t>(m) * p.res_ld + n]);
  if (p.act == LDMSEG_ACT_SILU) acc = silu_f(acc);
  if (p.out_dtype == LDMSEG_OUT_F32) {
    reinterpret_cast<float*>(p.out)[static_cast<size_t>(m) * p.out_ld + n] = acc;
  } else {
    const __nv_bfloat16 o = __float2bfloat16(acc);
    reinterpret_cast<__nv_bfloat16*>(p.out)[static_cast<size_t>(m) * p.out_ld + n] = o;
    acc = __bfloat162float(o);
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
    LDM_REQUIRE(p->seg_taps[s] == 1 || p->seg_taps[s] == 9, "igemm: taps must be 1 or 9");
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
  if (p->rowbias) LDM_REQUIRE(p->rowbias_ld % 4 == 0, "igemm: rowbias_ld must be a multiple of 4");
  if (p->stats) {
    LDM_REQUIRE((p->stats_hw > 0 ? p->stats_hw : hw) % 64 == 0, "igemm: fused statistics need rows per image %% 64 == 0");
    LDM_REQUIRE(p->act != LDMSEG_ACT_GEGLU, "igemm: fused statistics are not defined for GEGLU");
  }
  return 0;
}

template <int BN>
static int launch_igemm(const IgemmKParams& kp, int grid, cudaStream_t stream, int pdl) {
  using Cfg = IgemmCfg<BN>;
  static bool configured = false;
  if (!configured) {
    LDM_CUDA(cudaFuncSetAttribute(igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kIgemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl || g_pdl) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, igemm_kernel<BN>, kp);
  if (e != cudaSuccess) {
    set_error("igemm_kernel launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return check_launch("igemm_kernel");
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
  for (int i = 0; i < p->nsrc; ++i) {
    const uint64_t C = p->src_c[i];
    uint64_t dims[4] = {C, static_cast<uint64_t>(p->w), static_cast<uint64_t>(p->h),
                        static_cast<uint64_t>(p->nb)};
    uint64_t strides[3] = {C * 2, C * 2 * p->w, C * 2 * p->w * p->h};
    uint32_t box[4] = {BK, bw, bh, bb};
    if (int rc = encode_tmap_bf16(&kp.a_map[i], p->src[i], 4, dims, strides, box)) return rc;
  }
  const int m_tiles = (M + BM - 1) / BM;
  int bn = p->block_n;
  if (bn == 0) bn = choose_block_n(m_tiles, p->n, num_sms());
  LDM_REQUIRE(bn == 64 || bn == 128 || bn == 160 || bn == 256, "igemm: unsupported block_n %d", bn);
  {
    if (p->weight_tiled) {
      // [N/32][K/64][32][64]: every 32-row x 64-k block is 4 KB contiguous -> weight streaming reads
      // whole DRAM pages instead of 128-byte pieces at a K-row stride
      const uint64_t kblocks = static_cast<uint64_t>(p->ktot) / BK;
      uint64_t dims[4] = {BK, 32, kblocks, static_cast<uint64_t>((p->n + 31) / 32)};
      uint64_t strides[3] = {128, 4096, kblocks * 4096};
      uint32_t box[4] = {BK, 32, 1, static_cast<uint32_t>(bn / 32)};
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
  kp.residual = reinterpret_cast<const __nv_bfloat16*>(p->residual);
  kp.res_ld = p->res_ld;
  kp.out = p->out;
  kp.out_ld = p->out_ld;
  kp.out_f32 = p->out_dtype == LDMSEG_OUT_F32;
  kp.act = p->act;
  kp.split_k = p->split_k > 1 ? p->split_k : 1;
  if (kp.split_k > num_kb) kp.split_k = num_kb;
  if (kp.split_k > 1) {
    LDM_REQUIRE(p->workspace != nullptr && p->tile_counters != nullptr,
                "igemm: split_k > 1 needs workspace and tile_counters");
    const long long need = static_cast<long long>(kp.num_m_tiles) * kp.num_n_tiles * kp.split_k * BM * bn;
    LDM_REQUIRE(p->workspace_elems >= need, "igemm: split-K workspace too small (%lld f32 needed, %lld given)",
                need, static_cast<long long>(p->workspace_elems));
    LDM_REQUIRE(static_cast<long long>(kp.num_m_tiles) * kp.num_n_tiles <= 4096,
                "igemm: split-K supports at most 4096 output tiles (tile_counters holds 2 x 4096 int32)");
    kp.workspace = p->workspace;
    kp.counters = p->tile_counters;
  }
  kp.stats = p->stats;
  kp.debug = g_debug;
  kp.w_tiled = p->weight_tiled;
  kp.stats_hw = p->stats_hw > 0 ? p->stats_hw : HW;
  const long long work = static_cast<long long>(kp.num_m_tiles) * kp.num_n_tiles * kp.split_k;
  const int grid = static_cast<int>(work < num_sms() ? work : num_sms());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 64: return launch_igemm<64>(kp, grid, st, p->pdl);
    case 128: return launch_igemm<128>(kp, grid, st, p->pdl);
    case 160: return launch_igemm<160>(kp, grid, st, p->pdl);
    default: return launch_igemm<256>(kp, grid, st, p->pdl);
  }
}

extern "C" int ldmseg_igemm_simple(const ldmseg_igemm_params* p, void* stream) {
  if (int rc = validate(p)) return rc;
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
