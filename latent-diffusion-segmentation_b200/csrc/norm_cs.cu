// GroupNorm(+SiLU) apply pass that takes its statistics from the PRODUCER: the tcgen05 implicit-GEMM
// epilogue accumulates per-(image, channel) {sum, sum of squares} of the tensor it stores
// (ldmseg_igemm_params.stats), so the separate statistics pass over the activation (one full HBM read
// + one launch + one memset per GroupNorm, 61 per UNet forward) disappears.  Any grouping -- including
// groups that straddle the two sources of a virtual concat -- is derived here from channel sums.
//
// Replaces nn.GroupNorm + nn.SiLU of diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm and
// UNet.conv_norm_out (/root/reference/ldmseg/models/unet.py:428-430).
#include "common.h"
#include <cstdlib>
#include <type_traits>
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

__device__ __forceinline__ void unpack8c(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}

// grid (chunks, nb); block = C/8 * k threads (k pixel lanes): every thread owns ONE 8-channel octet for the
// whole launch, so its 8 scale / 8 shift values live in registers and the pixel loop is pure
// load -> 8 FMA (+ SiLU) -> store with no index arithmetic beyond an add (the previous version paid a 64-bit
// div/mod per 16 bytes and re-read scale/shift from shared memory: 1.6 TB/s at batch 8).
// dynamic smem: 2 * groups floats.
// SiLU with ONE MUFU op per value: t / (1 + e), e = 2^(-t log2 e) on the MUFU pipe, the reciprocal on the FMA pipe
// (integer-seeded Newton iteration, two steps: relative error 6.6e-6, 600x below the bf16 rounding of the output).  The apply pass at batch 8 was MUFU-bound
// (two MUFU per value: ncu mufu 44 %, 39 % of HBM); the FMA pipe has 8x the MUFU rate.
__device__ __forceinline__ float silu_1mufu(float t) {
  const float e = ex2_approx_f(fminf(-1.4426950408889634f * t, 80.f));
  const float d = 1.f + e;
  float r = __uint_as_float(0x7EF311C7u - __float_as_uint(d));
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  return t * r;
}

// SRC_F32: both sources are f32 (the fp32 residual stream); otherwise bf16.
// HOIST: issue the first trip's loads before the statistics phase (small batches: latency; costs ~20 registers)
template <int UNROLL, bool SRC_F32, bool HOIST>
__global__ void __launch_bounds__(512)
gn_apply_cs_kernel(const void* __restrict__ s0v, int c0, const float* __restrict__ cs0,
                   const void* __restrict__ s1v, int c1, const float* __restrict__ cs1, int hw,
                   int pix_per_cta, int groups, const float* __restrict__ gamma,
                   const float* __restrict__ beta, float eps, int silu, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float gstat[];  // [groups][2] = mean, rstd
  const int C = c0 + c1;
  const int C8 = C >> 3;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int oct = threadIdx.x % C8;      // this thread's channel octet
  const int plane = threadIdx.x / C8;    // pixel lane
  const int lanes = blockDim.x / C8;        // the block is rounded up to whole warps: threads beyond
  const bool idle = plane >= lanes;         // C8 * lanes only help with the group statistics
  const int ch = oct * 8;
  // affine parameters are weights: fetch them before waiting on the producer kernel
  float ga[8], be[8];
  {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + ch + 4));
    ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
    be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
  }
  pdl_sync();
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  const bool from0 = ch < c0;
  using src_t = typename std::conditional<SRC_F32, float, __nv_bfloat16>::type;
  const src_t* s0 = reinterpret_cast<const src_t*>(s0v);
  const src_t* s1 = reinterpret_cast<const src_t*>(s1v);
  const src_t* src = from0 ? s0 + static_cast<size_t>(b) * hw * c0 + ch
                           : s1 + static_cast<size_t>(b) * hw * c1 + (ch - c0);
  const int sld = from0 ? c0 : c1;
  __nv_bfloat16* dst = out + static_cast<size_t>(b) * hw * C + ch;
  constexpr int kVec = SRC_F32 ? 2 : 1;   // 16-byte vectors per octet
  // The first trip's activation loads do not depend on the statistics: issue them now, so that their L2 round trip
  // overlaps the one of the moments below (at batch 1 a thread makes exactly one trip: the whole kernel then costs
  // one round trip instead of two).
  uint4 u[UNROLL][kVec];
  int p0 = idle ? p_end : p_begin + plane;
  auto load_trip = [&](int pbase) {
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int p = pbase + k * lanes;
      if (p < p_end) {
#pragma unroll
        for (int v = 0; v < kVec; ++v)
          u[k][v] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(p) * sld) + v);
      }
    }
  };
  if constexpr (HOIST) load_trip(p0);
  // group moments from the producer's channel moments: one warp per group, several groups per warp.  All loads of
  // a warp's groups are issued before the first reduction, so the phase costs ONE L2 round trip instead of one per
  // group.  (Shared-memory float atomics -- one channel per thread -- were tried and cost 10 us per launch inside
  // the forward graph: profiles/r01_ablate_unet_b1_v12_gn_atomics.log.)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv_cnt = 1.f / (static_cast<float>(cpg) * static_cast<float>(hw));
  constexpr int kGroupsPerTrip = 4, kChanPerLane = 4;   // covers 32 groups on >= 8 warps, <= 128 channels per group
  for (int g0 = warp; g0 < groups; g0 += nw * kGroupsPerTrip) {
    float su[kGroupsPerTrip], sq[kGroupsPerTrip];
#pragma unroll
    for (int i = 0; i < kGroupsPerTrip; ++i) {
      const int g = g0 + i * nw;
      su[i] = 0.f;
      sq[i] = 0.f;
#pragma unroll
      for (int k = 0; k < kChanPerLane; ++k) {
        const int c = g * cpg + lane + 32 * k;
        if (g < groups && lane + 32 * k < cpg) {
          const float2 v = c < c0 ? __ldcg(reinterpret_cast<const float2*>(cs0 + (static_cast<size_t>(b) * c0 + c) * 2))
                                  : __ldcg(reinterpret_cast<const float2*>(cs1 + (static_cast<size_t>(b) * c1 + (c - c0)) * 2));
          su[i] += v.x;
          sq[i] += v.y;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kGroupsPerTrip; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        su[i] += __shfl_xor_sync(0xffffffffu, su[i], o);
        sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
      }
      const int g = g0 + i * nw;
      if (lane == 0 && g < groups) {
        const float mean = su[i] * inv_cnt;
        const float var = fmaxf(sq[i] * inv_cnt - mean * mean, 0.f);
        gstat[2 * g] = mean;
        gstat[2 * g + 1] = rsqrtf(var + eps);
      }
    }
  }
  __syncthreads();
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (ch + i) / cpg;
    sc[i] = ga[i] * gstat[2 * g + 1];
    sh[i] = be[i] - gstat[2 * g] * sc[i];
  }
  while (p0 < p_end) {
    if constexpr (!HOIST) load_trip(p0);
#pragma unroll
    for (int k = 0; k < UNROLL; ++k) {
      const int p = p0 + k * lanes;
      if (p >= p_end) break;
      float f[8];
      if constexpr (SRC_F32) {
        f[0] = __uint_as_float(u[k][0].x); f[1] = __uint_as_float(u[k][0].y);
        f[2] = __uint_as_float(u[k][0].z); f[3] = __uint_as_float(u[k][0].w);
        f[4] = __uint_as_float(u[k][kVec - 1].x); f[5] = __uint_as_float(u[k][kVec - 1].y);
        f[6] = __uint_as_float(u[k][kVec - 1].z); f[7] = __uint_as_float(u[k][kVec - 1].w);
      } else {
        unpack8c(u[k][0], f);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = fmaf(f[i], sc[i], sh[i]);
        f[i] = silu ? silu_1mufu(t) : t;
      }
      uint4 o;
      o.x = pack_bf16x2(f[0], f[1]);
      o.y = pack_bf16x2(f[2], f[3]);
      o.z = pack_bf16x2(f[4], f[5]);
      o.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(dst + static_cast<size_t>(p) * C) = o;
    }
    p0 += lanes * UNROLL;
    if constexpr (HOIST) load_trip(p0);
  }
}

}  // namespace ldm

using namespace ldm;

static int gn_apply_cs_launch(const void* src0, int c0, const float* chan_stats0, const void* src1, int c1,
                              const float* chan_stats1, int nb, int hw, int groups, const float* gamma,
                              const float* beta, float eps, int silu, void* out, int src_f32, void* stream) {
  if (!src1) c1 = 0;
  const int C = c0 + c1;
  LDM_REQUIRE(src0 && chan_stats0 && out && gamma && beta, "groupnorm_apply_cs: null pointer");
  LDM_REQUIRE(!src1 || chan_stats1, "groupnorm_apply_cs: second source needs its channel statistics");
  LDM_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm_apply_cs: C %% groups != 0");
  LDM_REQUIRE(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm_apply_cs: channels must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C8 = C / 8;
  LDM_REQUIRE(C8 <= 512 && C / groups <= 128, "groupnorm_apply_cs: at most 4096 channels, 128 per group");
  // block = C8 * lanes threads, about 384 wide
  int lanes = 384 / C8;
  if (lanes < 1) lanes = 1;
  const int threads = (C8 * lanes + 31) / 32 * 32;
  // enough CTAs to fill the machine twice over, but at least 4 pixels per lane per CTA
  int chunks = (2 * num_sms() + nb - 1) / nb;
  // (small images: one pixel per lane, so that the launch still spreads over tens of SMs -- LDMSEG_GN_MINPX overrides)
  static int min_px_env = -1;
  if (min_px_env < 0) {
    const char* e = getenv("LDMSEG_GN_MINPX");
    min_px_env = e ? atoi(e) : 0;
  }
  static int mid_env = -1;
  if (mid_env < 0) {
    const char* e = getenv("LDMSEG_GN_MINPX_MID");
    mid_env = e ? atoi(e) : 1;
  }
  static int top_env = -1;
  if (top_env < 0) {
    const char* e = getenv("LDMSEG_GN_MINPX_TOP");
    top_env = e ? atoi(e) : 4;
  }
  const int min_px = min_px_env > 0 ? min_px_env : (hw <= 256 ? 1 : hw <= 1024 ? mid_env : top_env);
  int cmax = (hw + min_px * lanes - 1) / (min_px * lanes);
  if (chunks > cmax) chunks = cmax;
  if (chunks < 1) chunks = 1;
  int ppc = (hw + chunks - 1) / chunks;
  chunks = (hw + ppc - 1) / ppc;
  const size_t smem = 2 * static_cast<size_t>(groups) * sizeof(float);
  // Three shapes of the same kernel.  One trip per thread (batch 1: latency-bound): the loads are hoisted above the
  // statistics phase (82 registers).  Several trips (batch >= 4: bandwidth-bound): the plain loop at 63 registers keeps
  // two CTAs per SM (hoisting there cost 160 us per batch-8 forward, profiles/r02_ablate_unet_b8*.log); many trips:
  // eight loads in flight per thread.
  const bool one_trip = ppc <= 4 * lanes;
  const bool deep = ppc >= 16 * lanes;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (src_f32)
    launch_kernel(gn_apply_cs_kernel<4, true, true>, dim3(chunks, nb), dim3(threads), smem, st, src0, c0, chan_stats0,
                  src1, c1, chan_stats1, hw, ppc, groups, gamma, beta, eps, silu, o);
  else if (deep)
    launch_kernel(gn_apply_cs_kernel<8, false, false>, dim3(chunks, nb), dim3(threads), smem, st, src0, c0, chan_stats0,
                  src1, c1, chan_stats1, hw, ppc, groups, gamma, beta, eps, silu, o);
  else if (one_trip)
    launch_kernel(gn_apply_cs_kernel<4, false, true>, dim3(chunks, nb), dim3(threads), smem, st, src0, c0, chan_stats0,
                  src1, c1, chan_stats1, hw, ppc, groups, gamma, beta, eps, silu, o);
  else
    launch_kernel(gn_apply_cs_kernel<4, false, false>, dim3(chunks, nb), dim3(threads), smem, st, src0, c0, chan_stats0,
                  src1, c1, chan_stats1, hw, ppc, groups, gamma, beta, eps, silu, o);
  return check_launch("gn_apply_cs_kernel");
}

extern "C" int ldmseg_groupnorm_apply_cs(const void* src0, int c0, const float* chan_stats0,
                                         const void* src1, int c1, const float* chan_stats1, int nb,
                                         int hw, int groups, const float* gamma, const float* beta,
                                         float eps, int silu, void* out, void* stream) {
  return gn_apply_cs_launch(src0, c0, chan_stats0, src1, c1, chan_stats1, nb, hw, groups, gamma, beta, eps, silu,
                            out, 0, stream);
}

extern "C" int ldmseg_groupnorm_apply_cs_f32(const void* src0, int c0, const float* chan_stats0,
                                             const void* src1, int c1, const float* chan_stats1, int nb,
                                             int hw, int groups, const float* gamma, const float* beta,
                                             float eps, int silu, void* out, void* stream) {
  return gn_apply_cs_launch(src0, c0, chan_stats0, src1, c1, chan_stats1, nb, hw, groups, gamma, beta, eps, silu,
                            out, 1, stream);
}
