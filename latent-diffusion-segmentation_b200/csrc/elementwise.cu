// Element-wise, layout and scheduler kernels of the sampling loop.  All are HBM/latency-bound:
// coalesced 16-byte accesses, no shared-memory staging unless a transpose needs it.
//
// Reference anchors (under /root/reference):
//   ddim step / sampler step   ldmseg/schedulers/ddim_scheduler.py:218-269,
//                              ldmseg/trainers/trainers_ldm_cond.py:1127-1159
//   time embedding             ldmseg/models/unet.py:303-307 (diffusers Timesteps/TimestepEmbedding)
//   GEGLU / Upsample2D / Downsample2D   diffusers blocks behind ldmseg/models/unet.py:361-425
//   bilinear x2 + argmax       ldmseg/models/vae.py:270, ldmseg/trainers/trainers_ldm_cond.py:428-433
#include "common.h"
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

__device__ __forceinline__ void unpack8e(const uint4& u, float (&f)[8]) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8e(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ---- GEGLU ------------------------------------------------------------------------------
__global__ void geglu_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int c,
                             __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const int c8 = c >> 3;
  const long long total = rows * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c8;
    const int v = static_cast<int>(i - r * c8);
    const uint4 uh = __ldg(reinterpret_cast<const uint4*>(x + r * 2 * c) + v);
    const uint4 ug = __ldg(reinterpret_cast<const uint4*>(x + r * 2 * c + c) + v);
    float h[8], g[8], o[8];
    unpack8e(uh, h);
    unpack8e(ug, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = h[j] * gelu_erf_f(g[j]);
    reinterpret_cast<uint4*>(out + r * c)[v] = pack8e(o);
  }
}

// ---- nearest x2 upsample ------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ src, int nb, int h, int w, int c8,
                                  uint4* __restrict__ out) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * 4 * h * w * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % c8);
    long long p = i / c8;
    const int ox = static_cast<int>(p % (2 * w));
    p /= 2 * w;
    const int oy = static_cast<int>(p % (2 * h));
    const int b = static_cast<int>(p / (2 * h));
    out[i] = __ldg(src + ((static_cast<long long>(b) * h + (oy >> 1)) * w + (ox >> 1)) * c8 + v);
  }
}

// ---- im2col for 3x3 stride-2 -----------------------------------------------------------------
__global__ void im2col_s2_kernel(const uint4* __restrict__ src, int nb, int h, int w, int c8,
                                 int pad_lo, int ho, int wo, uint4* __restrict__ out) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * ho * wo * 9 * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % c8);
    long long p = i / c8;
    const int tap = static_cast<int>(p % 9);
    p /= 9;
    const int ox = static_cast<int>(p % wo);
    p /= wo;
    const int oy = static_cast<int>(p % ho);
    const int b = static_cast<int>(p / ho);
    const int iy = 2 * oy + tap / 3 - pad_lo;
    const int ix = 2 * ox + tap % 3 - pad_lo;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w)
      val = __ldg(src + ((static_cast<long long>(b) * h + iy) * w + ix) * c8 + v);
    out[i] = val;
  }
}

// ---- layout conversions -----------------------------------------------------------------------
// NCHW f32 -> channel-last bf16 rows of cpad channels; writes channels [coff, coff+c), leaves the
// rest of the row untouched (caller zero-fills padding once).
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int nb, int c, int hw, int cpad,
                                    int coff, float scale, float shift,
                                    __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / hw;
    const long long p = i - b * hw;
    for (int ch = 0; ch < c; ++ch)
      out[i * cpad + coff + ch] =
          __float2bfloat16(__ldg(src + (b * c + ch) * hw + p) * scale + shift);
  }
}
__global__ void nchw_f32_to_nhwc_kernel(const float* __restrict__ src, int nb, int c, int hw, int ld,
                                        float scale, float* __restrict__ out) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * hw * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const long long pix = i / c;
    const long long b = pix / hw, p = pix - b * hw;
    out[pix * ld + ch] = __ldg(src + (b * c + ch) * hw + p) * scale;
  }
}
__global__ void nhwc_f32_to_nchw_kernel(const float* __restrict__ src, int nb, int c, int hw, int ld,
                                        float scale, float* __restrict__ out) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * c * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % hw;
    const long long t = i / hw;
    const int ch = static_cast<int>(t % c);
    const long long b = t / c;
    out[i] = __ldg(src + (b * hw + p) * ld + ch) * scale;
  }
}
__global__ void nhwc_bf16_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, int nb, int c, int hw,
                                         int ld, float scale, float* __restrict__ out) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * c * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long p = i % hw;
    const long long t = i / hw;
    const int ch = static_cast<int>(t % c);
    const long long b = t / c;
    out[i] = __bfloat162float(src[(b * hw + p) * ld + ch]) * scale;
  }
}

// ---- DDIM step (element-wise, any layout) -------------------------------------------------------
// Mirrors ddim_scheduler.py:238-267 term by term, fp32.
__device__ __forceinline__ void ddim_update(float eps_in, float x, float sa_t, float sb_t,
                                            float sa_p, float sb_p, int ptype, int clip,
                                            float clip_range, int use_clipped, float& prev,
                                            float& x0) {
  float eps;
  if (ptype == 0) {
    x0 = (x - sb_t * eps_in) / sa_t;
    eps = eps_in;
  } else if (ptype == 1) {
    x0 = eps_in;
    eps = (x - sa_t * x0) / sb_t;
  } else {
    x0 = sa_t * x - sb_t * eps_in;
    eps = sa_t * eps_in + sb_t * x;
  }
  if (clip) x0 = fminf(fmaxf(x0, -clip_range), clip_range);
  if (use_clipped) eps = (x - sa_t * x0) / sb_t;
  prev = sa_p * x0 + sb_p * eps;
}

__global__ void ddim_step_kernel(const float* __restrict__ mo, const float* __restrict__ x,
                                 long long n, float sa_t, float sb_t, float sa_p, float sb_p,
                                 int ptype, int clip, float clip_range, int use_clipped, float sigma,
                                 const float* __restrict__ noise, float* __restrict__ prev,
                                 float* __restrict__ x0o) {
  pdl_sync();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float pv, x0;
    ddim_update(mo[i], x[i], sa_t, sb_t, sa_p, sb_p, ptype, clip, clip_range, use_clipped, pv, x0);
    if (noise != nullptr) pv += sigma * noise[i];
    if (prev) prev[i] = pv;
    if (x0o) x0o[i] = x0;
  }
}

// Fused sampler step over pixels (one thread per pixel, 4 latent channels as float4).
// cfg: classifier-free guidance (trainers_ldm_cond.py:1143-1146): eps holds 2m rows (uncond | cond) and the next
// UNet input is written twice (rows i and m + i), as `torch.cat([latents] * 2)` does.
__global__ void sampler_step_kernel(const float4* __restrict__ eps, float4* __restrict__ lat,
                                    float4* __restrict__ x0o, const float4* __restrict__ rgb,
                                    uint4* __restrict__ unet_in, long long m,
                                    const float* __restrict__ coef, const int* __restrict__ step_ptr,
                                    int nsteps, int self_cond, const float* __restrict__ mask,
                                    const float4* __restrict__ known, const float4* __restrict__ noise,
                                    const float* __restrict__ sigma, int ptype, int clip, float clip_range,
                                    int cfg, float guidance) {
  pdl_sync();
  const int step = *step_ptr;
  const float sa_t = coef[step * 4 + 0], sb_t = coef[step * 4 + 1];
  const float sa_p = coef[step * 4 + 2], sb_p = coef[step * 4 + 3];
  const bool last = step == nsteps - 1;
  const float sg = (sigma != nullptr) ? sigma[step] : 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < m;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 e = eps[i];
    if (cfg) {
      const float4 et = eps[m + i];
      e.x += guidance * (et.x - e.x); e.y += guidance * (et.y - e.y);
      e.z += guidance * (et.z - e.z); e.w += guidance * (et.w - e.w);
    }
    const float4 x = lat[i];
    float4 pv, x0;
    ddim_update(e.x, x.x, sa_t, sb_t, sa_p, sb_p, ptype, clip, clip_range, 0, pv.x, x0.x);
    ddim_update(e.y, x.y, sa_t, sb_t, sa_p, sb_p, ptype, clip, clip_range, 0, pv.y, x0.y);
    ddim_update(e.z, x.z, sa_t, sb_t, sa_p, sb_p, ptype, clip, clip_range, 0, pv.z, x0.z);
    ddim_update(e.w, x.w, sa_t, sb_t, sa_p, sb_p, ptype, clip, clip_range, 0, pv.w, x0.w);
    if (noise != nullptr && !last) {
      const float4 z = noise[static_cast<long long>(step) * m + i];
      pv.x += sg * z.x; pv.y += sg * z.y; pv.z += sg * z.z; pv.w += sg * z.w;
    }
    float4 nxt = last ? x0 : pv;
    if (mask != nullptr) {
      const float mk = mask[i];
      const float4 kn = known[static_cast<long long>(step) * m + i];
      nxt.x = mk * kn.x + (1.f - mk) * nxt.x;
      nxt.y = mk * kn.y + (1.f - mk) * nxt.y;
      nxt.z = mk * kn.z + (1.f - mk) * nxt.z;
      nxt.w = mk * kn.w + (1.f - mk) * nxt.w;
    }
    lat[i] = nxt;
    if (x0o) x0o[i] = x0;
    if (unet_in) {
      const float4 r = rgb[i];
      uint4 a, b;
      a.x = pack_bf16x2(nxt.x, nxt.y);
      a.y = pack_bf16x2(nxt.z, nxt.w);
      a.z = pack_bf16x2(r.x, r.y);
      a.w = pack_bf16x2(r.z, r.w);
      if (self_cond) {
        b.x = pack_bf16x2(x0.x, x0.y);
        b.y = pack_bf16x2(x0.z, x0.w);
      } else {
        b.x = b.y = 0;
      }
      b.z = b.w = 0;
      unet_in[2 * i] = a;
      unet_in[2 * i + 1] = b;
      if (cfg) {
        unet_in[2 * (m + i)] = a;
        unet_in[2 * (m + i) + 1] = b;
      }
    }
  }
}

// ---- add_noise / remove_noise with per-sample timesteps (ddim_scheduler.py:155-216) ----------------------------
// x [nb, per] f32 (any layout with the sample index outermost), t int64 [nb] on the device, acp f32 table.
// mode 0: out = sqrt(a) * scale * x + sqrt(1 - a) * noise          (add_noise)
// mode 1: out = (x - sqrt(1 - a) * noise) / (sqrt(a) * scale)      (remove_noise)
// unet_in (optional, mode 0, NCHW input with 4 channels): also writes channels [0,4) of the channel-last bf16
// UNet input rows (the training-step no-grad forward, trainers_ldm_cond.py:824-831).
__global__ void noise_mix_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                 const long long* __restrict__ t, const float* __restrict__ acp, int nb,
                                 long long per, float scale, int mode, float* __restrict__ out,
                                 __nv_bfloat16* __restrict__ unet_in, int hw, int cpad) {
  pdl_sync();
  const long long total = static_cast<long long>(nb) * per;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per);
    const float a = acp[t[b]];
    // the reference raises the gathered fp32 alphas to the power 0.5 (ddim_scheduler.py:171-176)
    const float sa = sqrtf(a), sb = sqrtf(1.f - a);
    float o;
    if (mode == 0) o = sa * scale * x[i] + sb * noise[i];
    else o = (x[i] - sb * noise[i]) / (sa * scale);
    out[i] = o;
    if (unet_in != nullptr) {
      const long long r = i - static_cast<long long>(b) * per;   // c * hw + p
      const int c = static_cast<int>(r / hw);
      const long long pix = static_cast<long long>(b) * hw + (r - static_cast<long long>(c) * hw);
      unet_in[pix * cpad + c] = __float2bfloat16(o);
    }
  }
}
__global__ void advance_step_kernel(int* p) { *p = *p + 1; }

// ---- time embedding -------------------------------------------------------------------------------
__global__ void sinusoid_kernel(const float* __restrict__ t, int rows, int dim, int flip,
                                float freq_shift, float* __restrict__ out) {
  pdl_sync();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * half) return;
  const int r = i / half, k = i % half;
  const float expo = -logf(10000.f) * static_cast<float>(k) / (static_cast<float>(half) - freq_shift);
  const float a = t[r] * expf(expo);
  const float s = sinf(a), c = cosf(a);
  float* o = out + static_cast<size_t>(r) * dim;
  if (flip) {
    o[k] = c;
    o[half + k] = s;
  } else {
    o[k] = s;
    o[half + k] = c;
  }
}
// y[r, j] = act_out(b[j] + sum_k w[j,k] * act_in(x[r,k])); one warp per output column j, looping
// over rows in chunks of 8 so each weight row is streamed from HBM once.
__global__ void small_linear_kernel(const float* __restrict__ x, int rows, int k,
                                    const float* __restrict__ w, const float* __restrict__ b, int n,
                                    int silu_in, int silu_out, float* __restrict__ out, int out_ld) {
  pdl_sync();
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= n) return;
  const float* wr = w + static_cast<size_t>(j) * k;
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int c = lane; c < k; c += 32) {
      const float wv = __ldg(wr + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (r0 + i < rows) {
          float xv = x[static_cast<size_t>(r0 + i) * k + c];
          if (silu_in) xv = xv / (1.f + expf(-xv));
          acc[i] += xv * wv;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (lane == 0 && r0 + i < rows) {
        a += b ? b[j] : 0.f;
        if (silu_out) a = a / (1.f + expf(-a));
        out[static_cast<size_t>(r0 + i) * out_ld + j] = a;
      }
    }
  }
}

// dst[b, :] = table[*step_ptr, :] for every image b (per-step time-embedding bias selection)
__global__ void select_row_kernel(const float* __restrict__ table, int ncols,
                                  const int* __restrict__ step_ptr, int nb, float* __restrict__ dst) {
  pdl_sync();
  const int step = *step_ptr;
  const long long total = static_cast<long long>(nb) * ncols;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    dst[i] = table[static_cast<size_t>(step) * ncols + (i % ncols)];
}

// DDIM step with the timestep read from device memory (no host synchronisation).
__global__ void ddim_step_indexed_kernel(const float* __restrict__ mo, const float* __restrict__ x,
                                         long long n, const long long* __restrict__ t_ptr,
                                         const float* __restrict__ acp, int step_ratio,
                                         float final_alpha, int ptype, int clip, float clip_range,
                                         int use_clipped, float* __restrict__ prev,
                                         float* __restrict__ x0o) {
  pdl_sync();
  const long long t = *t_ptr;
  const long long tp = t - step_ratio;
  const float a_t = acp[t];
  const float a_p = tp >= 0 ? acp[tp] : final_alpha;
  const float sa_t = sqrtf(a_t), sb_t = sqrtf(1.f - a_t), sa_p = sqrtf(a_p), sb_p = sqrtf(1.f - a_p);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float pv, x0;
    ddim_update(mo[i], x[i], sa_t, sb_t, sa_p, sb_p, ptype, clip, clip_range, use_clipped, pv, x0);
    if (prev) prev[i] = pv;
    if (x0o) x0o[i] = x0;
  }
}

// ---- bilinear x2 (align_corners=False) ---------------------------------------------------------------
// CTA = (image b, output row pair {2k+1, 2k+2} -> input rows k, k+1, 32 input columns -> 64 output
// columns).  The two input rows are staged in shared memory as f32 [row][col][c+1].
template <bool kArgmax>
__global__ void bilinear2x_kernel(const void* __restrict__ src, int src_is_f32, int nb, int h, int w,
                                  int c, int ld, float* __restrict__ out, uint8_t* __restrict__ ids,
                                  float* __restrict__ maxprob) {
  pdl_sync();
  extern __shared__ float tile[];  // [2][34][c+1]
  const int cp = c + 1;
  const int xb = blockIdx.x * 32;          // first input column of this CTA's span
  const int kr = static_cast<int>(blockIdx.y) - 1;  // input row pair (kr, kr+1); kr = -1..h-1
  const int b = blockIdx.z;
  const int H2 = 2 * h, W2 = 2 * w;
  // stage input rows clamp(kr), clamp(kr+1), columns xb-1 .. xb+32 (clamped)
  for (int i = threadIdx.x; i < 2 * 34 * c; i += blockDim.x) {
    const int ch = i % c;
    const int col = (i / c) % 34;
    const int rr = i / (34 * c);
    int iy = min(max(kr + rr, 0), h - 1);
    int ix = min(max(xb - 1 + col, 0), w - 1);
    const size_t off = ((static_cast<size_t>(b) * h + iy) * w + ix) * ld + ch;
    tile[(rr * 34 + col) * cp + ch] =
        src_is_f32 ? reinterpret_cast<const float*>(src)[off]
                   : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[off]);
  }
  __syncthreads();
  // output rows: oy = 2*kr+1 (ly = .25 toward row kr+1) and oy = 2*kr+2 (ly = .75)
  for (int rsel = 0; rsel < 2; ++rsel) {
    const int oy = 2 * kr + 1 + rsel;
    if (oy < 0 || oy >= H2) continue;
    const float ly = rsel == 0 ? 0.25f : 0.75f;
    if (!kArgmax) {
      // thread -> (channel, ox) with ox fastest for coalesced NCHW stores
      for (int i = threadIdx.x; i < c * 64; i += blockDim.x) {
        const int oxl = i & 63, ch = i >> 6;
        const int ox = 2 * xb + oxl;
        if (ox >= W2) continue;
        // source x = ox/2 - 0.25: even ox -> cols (k-1,k) lx=.75 ; odd ox -> cols (k,k+1) lx=.25
        const int k = oxl >> 1;
        const int cl = (oxl & 1) ? k + 1 : k;  // tile column of the left sample (tile col 0 = xb-1)
        const float lx = (oxl & 1) ? 0.25f : 0.75f;
        const float v00 = tile[(0 * 34 + cl) * cp + ch], v01 = tile[(0 * 34 + cl + 1) * cp + ch];
        const float v10 = tile[(1 * 34 + cl) * cp + ch], v11 = tile[(1 * 34 + cl + 1) * cp + ch];
        const float top = v00 + (v01 - v00) * lx, bot = v10 + (v11 - v10) * lx;
        out[((static_cast<size_t>(b) * c + ch) * H2 + oy) * W2 + ox] = top + (bot - top) * ly;
      }
    } else {
      for (int oxl = threadIdx.x; oxl < 64; oxl += blockDim.x) {
        const int ox = 2 * xb + oxl;
        if (ox >= W2) continue;
        const int k = oxl >> 1;
        const int cl = (oxl & 1) ? k + 1 : k;
        const float lx = (oxl & 1) ? 0.25f : 0.75f;
        float best = -INFINITY, sum = 0.f;
        int bi = 0;
        for (int ch = 0; ch < c; ++ch) {
          const float v00 = tile[(0 * 34 + cl) * cp + ch], v01 = tile[(0 * 34 + cl + 1) * cp + ch];
          const float v10 = tile[(1 * 34 + cl) * cp + ch], v11 = tile[(1 * 34 + cl + 1) * cp + ch];
          const float top = v00 + (v01 - v00) * lx, bot = v10 + (v11 - v10) * lx;
          const float v = top + (bot - top) * ly;
          if (v > best) {
            sum = sum * __expf(best - v) + 1.f;
            best = v;
            bi = ch;
          } else {
            sum += __expf(v - best);
          }
        }
        const size_t o = (static_cast<size_t>(b) * H2 + oy) * W2 + ox;
        ids[o] = static_cast<uint8_t>(bi);
        if (maxprob) maxprob[o] = 1.f / sum;
      }
    }
  }
}

static inline int ew_grid(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ldm

using namespace ldm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int ldmseg_geglu(const void* x, int rows, int c, void* out, void* stream) {
  LDM_REQUIRE(x && out && c % 8 == 0, "geglu: bad arguments");
  const long long total = static_cast<long long>(rows) * (c / 8);
  launch_kernel(geglu_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), rows, c, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("geglu_kernel");
}

extern "C" int ldmseg_upsample2x(const void* src, int nb, int h, int w, int c, void* out,
                                 void* stream) {
  LDM_REQUIRE(src && out && c % 8 == 0, "upsample2x: bad arguments");
  const long long total = static_cast<long long>(nb) * 4 * h * w * (c / 8);
  launch_kernel(upsample2x_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), 
      reinterpret_cast<const uint4*>(src), nb, h, w, c / 8, reinterpret_cast<uint4*>(out));
  return check_launch("upsample2x_kernel");
}

extern "C" int ldmseg_im2col_s2(const void* src, int nb, int h, int w, int c, int pad_lo, void* out,
                                void* stream) {
  LDM_REQUIRE(src && out && c % 8 == 0 && h % 2 == 0 && w % 2 == 0, "im2col_s2: bad arguments");
  const int ho = h / 2, wo = w / 2;
  const long long total = static_cast<long long>(nb) * ho * wo * 9 * (c / 8);
  launch_kernel(im2col_s2_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), 
      reinterpret_cast<const uint4*>(src), nb, h, w, c / 8, pad_lo, ho, wo,
      reinterpret_cast<uint4*>(out));
  return check_launch("im2col_s2_kernel");
}

extern "C" int ldmseg_nchw_to_nhwc_bf16(const float* src, int nb, int c, int hw, int cpad, int coff,
                                        float scale, float shift, void* out, void* stream) {
  LDM_REQUIRE(src && out && coff + c <= cpad, "nchw_to_nhwc: bad arguments");
  const long long total = static_cast<long long>(nb) * hw;
  launch_kernel(nchw_to_nhwc_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), 
      src, nb, c, hw, cpad, coff, scale, shift, reinterpret_cast<__nv_bfloat16*>(out));
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int ldmseg_nhwc_f32_to_nchw(const float* src, int nb, int c, int hw, int ld, float scale,
                                       float* out, void* stream) {
  LDM_REQUIRE(src && out, "nhwc_f32_to_nchw: null pointer");
  const long long total = static_cast<long long>(nb) * c * hw;
  launch_kernel(nhwc_f32_to_nchw_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), src, nb, c, hw, ld, scale,
                                                                      out);
  return check_launch("nhwc_f32_to_nchw_kernel");
}

extern "C" int ldmseg_nchw_f32_to_nhwc(const float* src, int nb, int c, int hw, int ld, float scale,
                                       float* out, void* stream) {
  LDM_REQUIRE(src && out && ld >= c, "nchw_f32_to_nhwc: bad arguments");
  const long long total = static_cast<long long>(nb) * c * hw;
  launch_kernel(nchw_f32_to_nhwc_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), src, nb, c, hw, ld, scale, out);
  return check_launch("nchw_f32_to_nhwc_kernel");
}

extern "C" int ldmseg_nhwc_bf16_to_nchw(const void* src, int nb, int c, int hw, int ld, float scale,
                                        float* out, void* stream) {
  LDM_REQUIRE(src && out, "nhwc_bf16_to_nchw: null pointer");
  const long long total = static_cast<long long>(nb) * c * hw;
  launch_kernel(nhwc_bf16_to_nchw_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), 
      reinterpret_cast<const __nv_bfloat16*>(src), nb, c, hw, ld, scale, out);
  return check_launch("nhwc_bf16_to_nchw_kernel");
}

extern "C" int ldmseg_ddim_step(const float* model_out, const float* sample, int64_t n, float alpha_t,
                                float alpha_prev, int prediction_type, int clip, float clip_range,
                                int use_clipped, float sigma, const float* noise, float* prev_sample,
                                float* pred_x0, void* stream) {
  LDM_REQUIRE(model_out && sample && n >= 0, "ddim_step: bad arguments");
  LDM_REQUIRE(prediction_type >= 0 && prediction_type <= 2, "ddim_step: bad prediction_type");
  if (n == 0) return 0;
  // the reference raises fp32 0-dim tensors to the power 0.5 (ddim_scheduler.py:240,264,267)
  const float sa_t = sqrtf(alpha_t), sb_t = sqrtf(1.f - alpha_t);
  const float sa_p = sqrtf(alpha_prev), sb_p = sqrtf(1.f - alpha_prev - sigma * sigma);
  launch_kernel(ddim_step_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, ST(stream), 
      model_out, sample, n, sa_t, sb_t, sa_p, sb_p, prediction_type, clip, clip_range, use_clipped,
      sigma, sigma > 0.f ? noise : nullptr, prev_sample, pred_x0);
  return check_launch("ddim_step_kernel");
}

extern "C" int ldmseg_sampler_step(const float* eps, float* latents, float* x0,
                                   const float* rgb_latents, void* unet_in, int64_t m,
                                   const float* coef, const int* step_ptr, int nsteps, int self_cond,
                                   const float* mask, const float* known, const float* noise,
                                   const float* sigma, int prediction_type, int clip, float clip_range,
                                   int cfg, float guidance, void* stream) {
  LDM_REQUIRE(eps && latents && coef && step_ptr, "sampler_step: null pointer");
  LDM_REQUIRE(!unet_in || rgb_latents, "sampler_step: unet_in needs rgb_latents");
  LDM_REQUIRE(!mask || known, "sampler_step: mask needs known latents");
  LDM_REQUIRE(prediction_type >= 0 && prediction_type <= 2, "sampler_step: bad prediction_type");
  LDM_REQUIRE(!(cfg && self_cond), "sampler_step: guidance with self-conditioning is undefined in the reference "
                                   "(trainers_ldm_cond.py:1126-1150 concatenates a 2B batch with a B condition)");
  launch_kernel(sampler_step_kernel, dim3(ew_grid(m, 128)), dim3(128), 0, ST(stream), 
      reinterpret_cast<const float4*>(eps), reinterpret_cast<float4*>(latents),
      reinterpret_cast<float4*>(x0), reinterpret_cast<const float4*>(rgb_latents),
      reinterpret_cast<uint4*>(unet_in), m, coef, step_ptr, nsteps, self_cond, mask,
      reinterpret_cast<const float4*>(known), reinterpret_cast<const float4*>(noise), sigma, prediction_type, clip,
      clip_range, cfg, guidance);
  return check_launch("sampler_step_kernel");
}

extern "C" int ldmseg_noise_mix(const float* x, const float* noise, const int64_t* timesteps_dev,
                                const float* alphas_cumprod_dev, int nb, int64_t per_sample, float scale, int mode,
                                float* out, void* unet_in, int hw, int cpad, void* stream) {
  LDM_REQUIRE(x && noise && timesteps_dev && alphas_cumprod_dev && out, "noise_mix: null pointer");
  LDM_REQUIRE(mode == 0 || mode == 1, "noise_mix: mode must be 0 (add_noise) or 1 (remove_noise)");
  LDM_REQUIRE(!unet_in || (mode == 0 && hw > 0 && per_sample % hw == 0 && per_sample / hw <= cpad),
              "noise_mix: unet_in needs mode 0 and per_sample = channels * hw with channels <= cpad");
  const long long total = static_cast<long long>(nb) * per_sample;
  if (total == 0) return 0;
  launch_kernel(noise_mix_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), x, noise,
                reinterpret_cast<const long long*>(timesteps_dev), alphas_cumprod_dev, nb,
                static_cast<long long>(per_sample), scale, mode, out, reinterpret_cast<__nv_bfloat16*>(unet_in), hw, cpad);
  return check_launch("noise_mix_kernel");
}

extern "C" int ldmseg_advance_step(int* step_ptr, void* stream) {
  LDM_REQUIRE(step_ptr, "advance_step: null pointer");
  launch_kernel(advance_step_kernel, dim3(1), dim3(1), 0, ST(stream), step_ptr);
  return check_launch("advance_step_kernel");
}

extern "C" int ldmseg_timestep_sinusoid(const float* t, int rows, int dim, int flip_sin_to_cos,
                                        float freq_shift, float* out, void* stream) {
  LDM_REQUIRE(t && out && dim % 2 == 0, "timestep_sinusoid: bad arguments");
  const int total = rows * (dim / 2);
  launch_kernel(sinusoid_kernel, dim3((total + 127) / 128), dim3(128), 0, ST(stream), t, rows, dim, flip_sin_to_cos,
                                                              freq_shift, out);
  return check_launch("sinusoid_kernel");
}

extern "C" int ldmseg_small_linear(const float* x, int rows, int k, const float* w, const float* b,
                                   int n, int silu_in, int silu_out, float* out, int out_ld,
                                   void* stream) {
  LDM_REQUIRE(x && w && out, "small_linear: null pointer");
  const int wpb = 4;
  launch_kernel(small_linear_kernel, dim3((n + wpb - 1) / wpb), dim3(wpb * 32), 0, ST(stream), x, rows, k, w, b, n, silu_in,
                                                                        silu_out, out, out_ld);
  return check_launch("small_linear_kernel");
}

extern "C" int ldmseg_select_row(const float* table, int ncols, const int* step_ptr, int nb,
                                 float* dst, void* stream) {
  LDM_REQUIRE(table && step_ptr && dst, "select_row: null pointer");
  const long long total = static_cast<long long>(nb) * ncols;
  launch_kernel(select_row_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, ST(stream), table, ncols, step_ptr, nb, dst);
  return check_launch("select_row_kernel");
}

extern "C" int ldmseg_ddim_step_indexed(const float* model_out, const float* sample, int64_t n,
                                        const int64_t* timestep_dev, const float* alphas_cumprod_dev,
                                        int step_ratio, float final_alpha, int prediction_type,
                                        int clip, float clip_range, int use_clipped,
                                        float* prev_sample, float* pred_x0, void* stream) {
  LDM_REQUIRE(model_out && sample && timestep_dev && alphas_cumprod_dev, "ddim_step_indexed: null pointer");
  LDM_REQUIRE(prediction_type >= 0 && prediction_type <= 2, "ddim_step_indexed: bad prediction_type");
  if (n == 0) return 0;
  launch_kernel(ddim_step_indexed_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, ST(stream), 
      model_out, sample, n, reinterpret_cast<const long long*>(timestep_dev), alphas_cumprod_dev,
      step_ratio, final_alpha, prediction_type, clip, clip_range, use_clipped, prev_sample, pred_x0);
  return check_launch("ddim_step_indexed_kernel");
}

static int launch_bilinear(bool argmax, const void* src, int src_is_f32, int nb, int h, int w, int c,
                           int ld, float* out, uint8_t* ids, float* maxprob, void* stream) {
  LDM_REQUIRE(src && c >= 1 && c <= 256, "bilinear2x: bad arguments");
  const size_t smem = static_cast<size_t>(2) * 34 * (c + 1) * sizeof(float);
  dim3 grid((w + 31) / 32, h + 1, nb);
  if (argmax) {
    static bool cfg = false;
    if (!cfg) {
      LDM_CUDA(cudaFuncSetAttribute(bilinear2x_kernel<true>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 34 * 257 * 4));
      cfg = true;
    }
    launch_kernel(bilinear2x_kernel<true>, dim3(grid), dim3(64), smem, ST(stream), src, src_is_f32, nb, h, w, c, ld, nullptr,
                                                            ids, maxprob);
  } else {
    static bool cfg = false;
    if (!cfg) {
      LDM_CUDA(cudaFuncSetAttribute(bilinear2x_kernel<false>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 34 * 257 * 4));
      cfg = true;
    }
    launch_kernel(bilinear2x_kernel<false>, dim3(grid), dim3(256), smem, ST(stream), src, src_is_f32, nb, h, w, c, ld, out,
                                                              nullptr, nullptr);
  }
  return check_launch("bilinear2x_kernel");
}

extern "C" int ldmseg_bilinear2x_to_nchw(const void* src, int src_is_f32, int nb, int h, int w, int c,
                                         int ld, float* out, void* stream) {
  LDM_REQUIRE(out, "bilinear2x_to_nchw: null out");
  return launch_bilinear(false, src, src_is_f32, nb, h, w, c, ld, out, nullptr, nullptr, stream);
}

extern "C" int ldmseg_bilinear2x_argmax(const void* src, int src_is_f32, int nb, int h, int w, int c,
                                        int ld, uint8_t* ids, float* maxprob, void* stream) {
  LDM_REQUIRE(ids, "bilinear2x_argmax: null ids");
  return launch_bilinear(true, src, src_is_f32, nb, h, w, c, ld, nullptr, ids, maxprob, stream);
}
