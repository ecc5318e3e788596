// GroupNorm(+SiLU) apply pass that takes its statistics from the PRODUCER: the tcgen05 implicit-GEMM
// epilogue accumulates per-(image, channel) {sum, sum of squares} of the tensor it stores
// (ldmseg_igemm_params.stats), so the separate statistics pass over the activation (one full HBM read
// + one launch + one memset per GroupNorm, 61 per UNet forward) disappears.  Any grouping -- including
// groups that straddle the two sources of a virtual concat -- is derived here from channel sums.
//
// Replaces nn.GroupNorm + nn.SiLU of diffusers ResnetBlock2D.norm1/norm2, Transformer2DModel.norm and
// UNet.conv_norm_out (/root/reference/ldmseg/models/unet.py:428-430).
#include "common.h"
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

// grid (chunks, nb), block 256, dynamic smem: (2*C + 2*groups) floats
__global__ void gn_apply_cs_kernel(const __nv_bfloat16* __restrict__ s0, int c0,
                                   const float* __restrict__ cs0,
                                   const __nv_bfloat16* __restrict__ s1, int c1,
                                   const float* __restrict__ cs1, int hw, int pix_per_cta, int groups,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, int silu, __nv_bfloat16* __restrict__ out) {
  extern __shared__ float sm[];
  const int C = c0 + c1;
  float* scale = sm;
  float* shift = sm + C;
  float* gstat = sm + 2 * C;  // [groups][2] = mean, rstd
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  pdl_sync();
  // group moments from channel moments: one warp per group
  const float inv_cnt = 1.f / (static_cast<float>(cpg) * static_cast<float>(hw));
  for (int g = warp; g < groups; g += nw) {
    float su = 0.f, sq = 0.f;
    for (int c = g * cpg + lane; c < (g + 1) * cpg; c += 32) {
      const float* cs = c < c0 ? cs0 + (static_cast<size_t>(b) * c0 + c) * 2
                               : cs1 + (static_cast<size_t>(b) * c1 + (c - c0)) * 2;
      su += cs[0];
      sq += cs[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      su += __shfl_xor_sync(0xffffffffu, su, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
      const float mean = su * inv_cnt;
      const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
      gstat[2 * g] = mean;
      gstat[2 * g + 1] = rsqrtf(var + eps);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float ga = gamma[c] * gstat[2 * g + 1];
    scale[c] = ga;
    shift[c] = beta[c] - gstat[2 * g] * ga;
  }
  __syncthreads();
  const int C8 = C >> 3;
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  const long long total = static_cast<long long>(p_end - p_begin) * C8;
  for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int p = p_begin + static_cast<int>(idx / C8);
    const int ch = static_cast<int>(idx % C8) * 8;
    const size_t pix = static_cast<size_t>(b) * hw + p;
    const __nv_bfloat16* src = ch < c0 ? s0 + pix * c0 + ch : s1 + pix * c1 + (ch - c0);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src));
    float f[8];
    unpack8c(u, f);
    uint4 o;
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = f[i] * scale[ch + i] + shift[ch + i];
      y[i] = silu ? silu_f(t) : t;
    }
    o.x = pack_bf16x2(y[0], y[1]);
    o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]);
    o.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(out + pix * C + ch) = o;
  }
}

}  // namespace ldm

using namespace ldm;

extern "C" int ldmseg_groupnorm_apply_cs(const void* src0, int c0, const float* chan_stats0,
                                         const void* src1, int c1, const float* chan_stats1, int nb,
                                         int hw, int groups, const float* gamma, const float* beta,
                                         float eps, int silu, void* out, void* stream) {
  if (!src1) c1 = 0;
  const int C = c0 + c1;
  LDM_REQUIRE(src0 && chan_stats0 && out && gamma && beta, "groupnorm_apply_cs: null pointer");
  LDM_REQUIRE(!src1 || chan_stats1, "groupnorm_apply_cs: second source needs its channel statistics");
  LDM_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm_apply_cs: C %% groups != 0");
  LDM_REQUIRE(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm_apply_cs: channels must be multiples of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int C8 = C / 8;
  int chunks = (4 * num_sms() + nb - 1) / nb;
  int cmax = static_cast<int>((static_cast<long long>(hw) * C8 + 1023) / 1024);
  if (chunks > cmax) chunks = cmax;
  if (chunks < 1) chunks = 1;
  int ppc = (hw + chunks - 1) / chunks;
  chunks = (hw + ppc - 1) / ppc;
  const size_t smem = (2 * static_cast<size_t>(C) + 2 * groups) * sizeof(float);
  launch_kernel(gn_apply_cs_kernel, dim3(chunks, nb), dim3(256), smem, st,
                reinterpret_cast<const __nv_bfloat16*>(src0), c0, chan_stats0,
                reinterpret_cast<const __nv_bfloat16*>(src1), c1, chan_stats1, hw, ppc, groups, gamma,
                beta, eps, silu, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("gn_apply_cs_kernel");
}
