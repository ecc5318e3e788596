// GroupNorm(+SiLU) and LayerNorm(+SiLU) for channel-last bf16 activations.  HBM-bound kernels:
// 16-byte vector loads, fp32 statistics, one read for statistics + one read/write for apply.
//
// GroupNorm replaces nn.GroupNorm + SiLU of diffusers ResnetBlock2D.norm1/norm2,
// Transformer2DModel.norm and UNet.conv_norm_out (/root/reference/ldmseg/models/unet.py:428-430),
// and the decoder GroupNorm of GeneralVAESeg (/root/reference/ldmseg/models/vae.py:162).  It reads
// a *virtual concatenation* of two sources so torch.cat([hidden, skip], 1) of the up blocks is
// never materialised un-normalised.
// LayerNorm replaces nn.LayerNorm (BasicTransformerBlock.norm1/norm3) and LayerNorm2d
// (/root/reference/ldmseg/models/vae.py:309-322: biased variance, eps inside the sqrt).
#include "common.h"
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ---- GroupNorm statistics: grid (chunks, nb), block (C/8, PY) ------------------------------
// Each thread owns 8 consecutive channels of the virtual concat and strides over the pixels of
// its chunk; per-channel sums are folded into groups through shared-memory atomics, then one
// global atomicAdd per (group, moment) per CTA.
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ s0, int c0,
                                const __nv_bfloat16* __restrict__ s1, int c1, int hw,
                                int pix_per_cta, int groups, float* __restrict__ stats) {
  pdl_sync();
  __shared__ float sg[2 * 64];
  const int C = c0 + c1;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tid; i < 2 * groups; i += blockDim.x * blockDim.y) sg[i] = 0.f;
  __syncthreads();
  const int ch = threadIdx.x * 8;
  const __nv_bfloat16* base;
  int cs, cc;
  if (ch < c0) {
    base = s0; cs = c0; cc = ch;
  } else {
    base = s1; cs = c1; cc = ch - c0;
  }
  const int p_begin = blockIdx.x * pix_per_cta;
  const int p_end = min(hw, p_begin + pix_per_cta);
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  for (int p = p_begin + threadIdx.y; p < p_end; p += blockDim.y) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(
        base + (static_cast<size_t>(b) * hw + p) * cs + cc));
    float f[8];
    unpack8(u, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i] += f[i];
      ss[i] += f[i] * f[i];
    }
  }
  // fold channels into groups (8 consecutive channels touch at most two groups when cpg >= 8,
  // more when cpg < 8; handle generally)
  int g_prev = ch / cpg;
  float as = 0.f, ass = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (ch + i) / cpg;
    if (g != g_prev) {
      atomicAdd(&sg[2 * g_prev], as);
      atomicAdd(&sg[2 * g_prev + 1], ass);
      as = ass = 0.f;
      g_prev = g;
    }
    as += s[i];
    ass += ss[i];
  }
  atomicAdd(&sg[2 * g_prev], as);
  atomicAdd(&sg[2 * g_prev + 1], ass);
  __syncthreads();
  for (int i = tid; i < 2 * groups; i += blockDim.x * blockDim.y)
    atomicAdd(&stats[static_cast<size_t>(b) * 2 * groups + i], sg[i]);
}

// ---- GroupNorm apply (+SiLU): grid (chunks, nb), block 256, smem 2*C floats -----------------
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ s0, int c0,
                                const __nv_bfloat16* __restrict__ s1, int c1, int hw,
                                int pix_per_cta, int groups, const float* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                float eps, int silu, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  extern __shared__ float sm[];
  const int C = c0 + c1;
  float* scale = sm;
  float* shift = sm + C;
  const int cpg = C / groups;
  const int b = blockIdx.y;
  const float inv_cnt = 1.f / (static_cast<float>(cpg) * static_cast<float>(hw));
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float su = stats[(static_cast<size_t>(b) * groups + g) * 2];
    const float sq = stats[(static_cast<size_t>(b) * groups + g) * 2 + 1];
    const float mean = su * inv_cnt;
    const float var = fmaxf(sq * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float ga = gamma[c] * rstd;
    scale[c] = ga;
    shift[c] = beta[c] - mean * ga;
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
    unpack8(u, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = f[i] * scale[ch + i] + shift[ch + i];
      f[i] = silu ? silu_f(y) : y;
    }
    *reinterpret_cast<uint4*>(out + pix * C + ch) = pack8(f);
  }
}

// ---- LayerNorm over channels, one warp per row ---------------------------------------------
// SRC_F32: the row is f32 (fp32 residual stream) instead of bf16.
template <int MAXV, bool SRC_F32>
__global__ void layernorm_kernel(const void* __restrict__ srcv, int rows, int c,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, int silu, __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = c >> 3;
  float f[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      if constexpr (SRC_F32) {
        const float4* s = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(srcv) +
                                                          static_cast<size_t>(row) * c);
        const float4 a = __ldg(s + 2 * v), b = __ldg(s + 2 * v + 1);
        f[i][0] = a.x; f[i][1] = a.y; f[i][2] = a.z; f[i][3] = a.w;
        f[i][4] = b.x; f[i][5] = b.y; f[i][6] = b.z; f[i][7] = b.w;
      } else {
        const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(srcv) +
                                                        static_cast<size_t>(row) * c);
        unpack8(__ldg(s + v), f[i]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(c);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(c) + eps);
  uint4* d = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * c);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * v + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v);
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * v + 1);
      const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = (f[i][j] - mean) * rstd * ga[j] + be[j];
        y[j] = silu ? silu_f(t) : t;
      }
      d[v] = pack8(y);
    }
  }
}

// ---- ConvTranspose2d(k2,s2) pixel shuffle + LayerNorm2d + SiLU -------------------------------
// src [nb*h*w, 4*c] with column block t = ky*2+kx; one warp per (pixel, tap).
__global__ void convt_shuffle_ln_kernel(const __nv_bfloat16* __restrict__ src, int nb, int h, int w,
                                        int c, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float eps, int silu,
                                        __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const long long wid = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(nb) * h * w * 4;
  if (wid >= total) return;
  const int tap = static_cast<int>(wid & 3);
  const long long m = wid >> 2;
  const int x = static_cast<int>(m % w);
  const int y = static_cast<int>((m / w) % h);
  const int b = static_cast<int>(m / (static_cast<long long>(w) * h));
  const int ky = tap >> 1, kx = tap & 1;
  const int nv = c >> 3;  // <= 32*4
  const uint4* s = reinterpret_cast<const uint4*>(src + (m * 4 + tap) * c);
  float f[4][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      unpack8(__ldg(s + v), f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += f[i][j];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(c);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(c) + eps);
  const size_t opix = (static_cast<size_t>(b) * (2 * h) + (2 * y + ky)) * (2 * w) + (2 * x + kx);
  uint4* d = reinterpret_cast<uint4*>(out + opix * c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = lane + i * 32;
    if (v < nv) {
      float yv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = v * 8 + j;
        const float t = (f[i][j] - mean) * rstd * __ldg(gamma + ch) + __ldg(beta + ch);
        yv[j] = silu ? silu_f(t) : t;
      }
      d[v] = pack8(yv);
    }
  }
}

// ---- row softmax: out[r, :] = softmax(scale * s[r, :]) (f32 in, bf16 out), one CTA per row ------
// Used by the single-head (d = 512) spatial attention of the AutoencoderKL mid block
// (diffusers 0.16.1 AttentionBlock: softmax in fp32 over baddbmm scores).
__global__ void softmax_rows_kernel(const float* __restrict__ s, int cols, float scale,
                                    __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  __shared__ float red[32];
  const float* row = s + static_cast<size_t>(blockIdx.x) * cols;
  __nv_bfloat16* orow = out + static_cast<size_t>(blockIdx.x) * cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x * 4; i < cols; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    mx = fmaxf(mx, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < nw; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x * 4; i < cols; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    sum += __expf((v.x - mx) * scale) + __expf((v.y - mx) * scale) + __expf((v.z - mx) * scale) +
           __expf((v.w - mx) * scale);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < nw; ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x * 4; i < cols; i += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    uint2 u;
    u.x = pack_bf16x2(__expf((v.x - mx) * scale) * inv, __expf((v.y - mx) * scale) * inv);
    u.y = pack_bf16x2(__expf((v.z - mx) * scale) * inv, __expf((v.w - mx) * scale) * inv);
    *reinterpret_cast<uint2*>(orow + i) = u;
  }
}

}  // namespace ldm

using namespace ldm;

extern "C" int ldmseg_softmax_rows(const float* s, int rows, int cols, float scale, void* out,
                                   void* stream) {
  LDM_REQUIRE(s && out && cols % 4 == 0, "softmax_rows: bad arguments");
  launch_kernel(softmax_rows_kernel, dim3(rows), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      s, cols, scale, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("softmax_rows_kernel");
}

extern "C" int ldmseg_groupnorm(const void* src0, int c0, const void* src1, int c1, int nb, int hw,
                                int groups, const float* gamma, const float* beta, float eps,
                                int silu, void* out, float* stats, void* stream) {
  const int C = c0 + (src1 ? c1 : 0);
  if (!src1) c1 = 0;
  LDM_REQUIRE(src0 && out && stats && gamma && beta, "groupnorm: null pointer");
  LDM_REQUIRE(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: C %% groups != 0");
  LDM_REQUIRE(c0 % 8 == 0 && c1 % 8 == 0, "groupnorm: channels must be multiples of 8");
  LDM_REQUIRE(C / 8 <= 1024, "groupnorm: too many channels");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  LDM_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * nb, st));
  const int C8 = C / 8;
  int py = 256 / C8;
  if (py < 1) py = 1;
  // enough CTAs to cover the machine ~2x, at least 8 pixel-iterations per thread when possible
  int chunks = (2 * num_sms() + nb - 1) / nb;
  int max_chunks = (hw + py * 4 - 1) / (py * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int ppc = (hw + chunks - 1) / chunks;
  chunks = (hw + ppc - 1) / ppc;
  dim3 grid(chunks, nb), block(C8, py);
  launch_kernel(gn_stats_kernel, dim3(grid), dim3(block), 0, st, reinterpret_cast<const __nv_bfloat16*>(src0), c0,
                                          reinterpret_cast<const __nv_bfloat16*>(src1), c1, hw, ppc,
                                          groups, stats);
  if (int rc = check_launch("gn_stats_kernel")) return rc;
  // apply: ~4 vectors per thread per CTA at least
  int achunks = (4 * num_sms() + nb - 1) / nb;
  int amax = static_cast<int>((static_cast<long long>(hw) * C8 + 1023) / 1024);
  if (achunks > amax) achunks = amax;
  if (achunks < 1) achunks = 1;
  int appc = (hw + achunks - 1) / achunks;
  achunks = (hw + appc - 1) / appc;
  dim3 agrid(achunks, nb);
  launch_kernel(gn_apply_kernel, dim3(agrid), dim3(256), 2 * C * sizeof(float), st, 
      reinterpret_cast<const __nv_bfloat16*>(src0), c0, reinterpret_cast<const __nv_bfloat16*>(src1),
      c1, hw, appc, groups, stats, gamma, beta, eps, silu, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("gn_apply_kernel");
}

template <bool SRC_F32>
static int layernorm_launch(const void* src, int rows, int c, const float* gamma, const float* beta, float eps,
                            int silu, void* out, void* stream) {
  LDM_REQUIRE(src && out && gamma && beta, "layernorm: null pointer");
  LDM_REQUIRE(c % 8 == 0 && c <= 2048, "layernorm: c must be a multiple of 8 and <= 2048");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int wpb = 8;
  const int grid = (rows + wpb - 1) / wpb;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (c <= 512)
    launch_kernel(layernorm_kernel<2, SRC_F32>, dim3(grid), dim3(wpb * 32), 0, st, src, rows, c, gamma, beta, eps, silu, o);
  else if (c <= 1280)
    launch_kernel(layernorm_kernel<5, SRC_F32>, dim3(grid), dim3(wpb * 32), 0, st, src, rows, c, gamma, beta, eps, silu, o);
  else
    launch_kernel(layernorm_kernel<8, SRC_F32>, dim3(grid), dim3(wpb * 32), 0, st, src, rows, c, gamma, beta, eps, silu, o);
  return check_launch("layernorm_kernel");
}

extern "C" int ldmseg_layernorm(const void* src, int rows, int c, const float* gamma,
                                const float* beta, float eps, int silu, void* out, void* stream) {
  return layernorm_launch<false>(src, rows, c, gamma, beta, eps, silu, out, stream);
}

extern "C" int ldmseg_layernorm_f32(const void* src, int rows, int c, const float* gamma,
                                    const float* beta, float eps, int silu, void* out, void* stream) {
  return layernorm_launch<true>(src, rows, c, gamma, beta, eps, silu, out, stream);
}

extern "C" int ldmseg_convt_shuffle_ln(const void* src, int nb, int h, int w, int c,
                                       const float* gamma, const float* beta, float eps, int silu,
                                       void* out, void* stream) {
  LDM_REQUIRE(src && out && gamma && beta, "convt_shuffle_ln: null pointer");
  LDM_REQUIRE(c % 8 == 0 && c <= 1024, "convt_shuffle_ln: c must be a multiple of 8 and <= 1024");
  const long long warps = static_cast<long long>(nb) * h * w * 4;
  const int wpb = 8;
  const long long grid = (warps + wpb - 1) / wpb;
  launch_kernel(convt_shuffle_ln_kernel, dim3(static_cast<unsigned>(grid)), dim3(wpb * 32), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(src), nb, h, w, c, gamma, beta, eps, silu,
      reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("convt_shuffle_ln_kernel");
}
