// Self-attention forward (non-causal) on tcgen05: S = Q K^T into TMEM, online softmax in
// registers (exp2, fp32 statistics), P (bf16) through shared memory, O += P V accumulated in TMEM.
//
// Replaces F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 for the 16
// Transformer2DModel blocks reached from /root/reference/ldmseg/models/unet.py:361-425
// (cross-attention is stripped by unet.py:83-105, so only self-attention remains).
// Shapes at a 64x64 latent: (tokens, head_dim) = (4096,40) (1024,80) (256,160) (64,160), 8 heads.
//
// Input is the fused QKV projection output, bf16 [nb*ntok, 3*heads*d] (q | k | v); a 5-D TMA map
// (d, head, which, token, image) loads 64-column panels and zero-fills columns >= d, so head
// dims that are not multiples of 64 (40, 80, 160) need no padding in HBM.
//
// CTA = 192 threads, one (image, head, 128-query tile):
//   warp 0  TMA producer   warp 1  MMA issuer + TMEM allocator   warps 2..5  softmax / correction
// TMEM: S double-buffered at columns [0,128) and [128,256), O at [256, 256+d).
#include "common.h"
#include <cstring>
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

constexpr int kAttnThreads = 192;

template <int D>
struct AttnCfg {
  static constexpr int BKV = (D <= 80) ? 128 : 64;
  static constexpr int kPanels = (D + 63) / 64;
  static constexpr int kDK = (D + 15) / 16 * 16;
  static constexpr int kQBytes = kPanels * 128 * 128;
  static constexpr int kKBytes = kPanels * BKV * 128;
  static constexpr int kPBytes = (BKV / 64) * 128 * 128;
  static constexpr int kSmemBytes = kQBytes + 2 * 2 * kKBytes + kPBytes + 16 * 8 + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kOCol = 256;
};

struct alignas(64) AttnKParams {
  CUtensorMap map_q;
  CUtensorMap map_kv;
  __nv_bfloat16* out;
  int nb, ntok, heads;
  float scale_log2;  // (1/sqrt(d)) * log2(e)
};

template <int D>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int BKV = Cfg::BKV;
  constexpr int kPanels = Cfg::kPanels;
  constexpr int kDK = Cfg::kDK;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sm_q = smem;
  uint8_t* sm_k = sm_q + Cfg::kQBytes;               // [2 stages][K]
  uint8_t* sm_v = sm_k + 2 * Cfg::kKBytes;           // [2 stages][V]
  uint8_t* sm_p = sm_v + 2 * Cfg::kKBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + Cfg::kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int q_tiles = (p.ntok + 127) / 128;
  const int qt = blockIdx.x % q_tiles;
  const int head = (blockIdx.x / q_tiles) % p.heads;
  const int b = blockIdx.x / (q_tiles * p.heads);
  const int q0 = qt * 128;
  const int T = (p.ntok + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();  // the prologue above overlapped the previous kernel; inputs are read from here on

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, Cfg::kQBytes);
      for (int pn = 0; pn < kPanels; ++pn)
        tma_load_5d(sm_q + pn * 128 * 128, &p.map_q, q_full, pn * 64, head, 0, q0, b);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t n = static_cast<uint32_t>(j >> 1);
        mbar_wait(&kv_empty[st], (n & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * Cfg::kKBytes);
        for (int pn = 0; pn < kPanels; ++pn) {
          tma_load_5d(sm_k + st * Cfg::kKBytes + pn * BKV * 128, &p.map_kv, &kv_full[st], pn * 64,
                      head, 1, j * BKV, b);
          tma_load_5d(sm_v + st * Cfg::kKBytes + pn * BKV * 128, &p.map_kv, &kv_full[st], pn * 64,
                      head, 2, j * BKV, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kDK, 0, 1);
      auto issue_s = [&](int j) {
        const int st = j & 1;
        const uint32_t d_tmem = tmem_base + (j & 1) * 128;
#pragma unroll
        for (int ks = 0; ks < kDK / 16; ++ks) {
          const int pn = ks >> 2, kk = ks & 3;
          const uint64_t a = make_smem_desc_sw128(smem_u32(sm_q + pn * 128 * 128), 16, 1024) + 2 * kk;
          const uint64_t bd =
              make_smem_desc_sw128(smem_u32(sm_k + st * Cfg::kKBytes + pn * BKV * 128), 16, 1024) +
              2 * kk;
          umma_bf16(d_tmem, a, bd, idesc_s, ks > 0 ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      umma_commit(&s_full[0]);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        if (j + 1 < T) {
          const int ns = (j + 1) & 1;
          mbar_wait(&kv_full[ns], static_cast<uint32_t>((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(j + 1);
          umma_commit(&s_full[ns]);
        }
        mbar_wait(p_full, static_cast<uint32_t>(j) & 1);
        tc_fence_after();
        const uint32_t o_tmem = tmem_base + Cfg::kOCol;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const int pn = ks >> 2, kk = ks & 3;
          const uint64_t a = make_smem_desc_sw128(smem_u32(sm_p + pn * 128 * 128), 16, 1024) + 2 * kk;
          // V tile: rows = kv (128 B each), MN(d)-major; 16 kv rows per k-step = 2048 B
          const uint64_t bd = make_smem_desc_sw128(
              smem_u32(sm_v + st * Cfg::kKBytes + ks * 2048), BKV * 128, 1024);
          umma_bf16(o_tmem, a, bd, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax / correction
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < T; ++j) {
      const int buf = j & 1;
      mbar_wait(&s_full[buf], static_cast<uint32_t>(j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + lane_off + buf * 128;
      const int kv_valid = min(BKV, p.ntok - j * BKV);  // columns < kv_valid are real tokens
      // pass 1: row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(s_addr + c, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, mx * p.scale_log2);
      const float alpha = exp2f(m_run - m_new);
      // P buffer and O are free once PV of the previous tile has completed
      if (j > 0) {
        mbar_wait(pv_done, static_cast<uint32_t>(j - 1) & 1);
        tc_fence_after();
      }
      // pass 2: p = exp2(s*scale - m), row sum, bf16 P tile into shared memory (K-major, SW128)
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < BKV; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(s_addr + c, r);
        tmem_wait_ld();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = exp2f(__uint_as_float(r[i]) * p.scale_log2 - m_new);
          pv[i] = (c + i < kv_valid) ? e : 0.f;
          rowsum += pv[i];
        }
        uint8_t* prow = sm_p + (c >> 6) * (128 * 128) + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c & 63) >> 3) + g;  // 16-byte chunk index within the 128-byte row
          uint4 u;
          u.x = pack_bf16x2(pv[g * 8 + 0], pv[g * 8 + 1]);
          u.y = pack_bf16x2(pv[g * 8 + 2], pv[g * 8 + 3]);
          u.z = pack_bf16x2(pv[g * 8 + 4], pv[g * 8 + 5]);
          u.w = pack_bf16x2(pv[g * 8 + 6], pv[g * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = u;
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      // correction: O *= alpha (skipped warp-uniformly when no row of this warp changed its max)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        const uint32_t o_addr = tmem_base + lane_off + Cfg::kOCol;
#pragma unroll 1
        for (int c = 0; c < kDK; c += 8) {
          uint32_t r[8];
          tmem_ld_32x8(o_addr + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x8(o_addr + c, r);
        }
        tmem_wait_st();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---------------------------------------------------------------- final normalisation
    mbar_wait(pv_done, static_cast<uint32_t>(T - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const uint32_t o_addr = tmem_base + lane_off + Cfg::kOCol;
    const int tok = q0 + row;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.ntok + tok) * (p.heads * D) + head * D;
#pragma unroll 1
    for (int c = 0; c < kDK; c += 8) {
      uint32_t r[8];
      tmem_ld_32x8(o_addr + c, r);
      tmem_wait_ld();
      if (tok < p.ntok && c < D) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
        u.y = pack_bf16x2(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
        u.z = pack_bf16x2(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
        u.w = pack_bf16x2(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        *reinterpret_cast<uint4*>(dst + c) = u;
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
// Reference-grade attention: one thread per query row, K/V streamed through shared memory.
template <int D>
__global__ void attn_simple_kernel(const __nv_bfloat16* __restrict__ qkv, int nb, int ntok, int heads,
                                   __nv_bfloat16* __restrict__ out) {
  pdl_sync();
  constexpr int TK = 32;
  __shared__ float sk[TK][D];
  __shared__ float sv[TK][D];
  const int q_tiles = (ntok + blockDim.x - 1) / blockDim.x;
  const int qt = blockIdx.x % q_tiles;
  const int head = (blockIdx.x / q_tiles) % heads;
  const int b = blockIdx.x / (q_tiles * heads);
  const int C = heads * D;
  const int tok = qt * blockDim.x + threadIdx.x;
  const bool valid = tok < ntok;
  float qv[D], acc[D];
  const float scale = rsqrtf(static_cast<float>(D));
  for (int i = 0; i < D; ++i) {
    qv[i] = valid ? __bfloat162float(qkv[(static_cast<size_t>(b) * ntok + tok) * 3 * C + head * D + i]) * scale
                  : 0.f;
    acc[i] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < ntok; k0 += TK) {
    __syncthreads();
    for (int i = threadIdx.x; i < TK * D; i += blockDim.x) {
      const int r = i / D, c = i % D;
      const int kt = k0 + r;
      float kvv = 0.f, vvv = 0.f;
      if (kt < ntok) {
        const size_t base = (static_cast<size_t>(b) * ntok + kt) * 3 * C + head * D + c;
        kvv = __bfloat162float(qkv[base + C]);
        vvv = __bfloat162float(qkv[base + 2 * C]);
      }
      sk[r][c] = kvv;
      sv[r][c] = vvv;
    }
    __syncthreads();
    const int lim = min(TK, ntok - k0);
    for (int r = 0; r < lim; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < D; ++i) s += qv[i] * sk[r][i];
      const float mn = fmaxf(m, s);
      const float a = __expf(m - mn);
      const float pe = __expf(s - mn);
      l = l * a + pe;
#pragma unroll
      for (int i = 0; i < D; ++i) acc[i] = acc[i] * a + pe * sv[r][i];
      m = mn;
    }
  }
  if (valid) {
    const float inv = 1.f / l;
    for (int i = 0; i < D; ++i)
      out[(static_cast<size_t>(b) * ntok + tok) * C + head * D + i] = __float2bfloat16(acc[i] * inv);
  }
}

template <int D>
static int launch_attn(const void* qkv, int nb, int ntok, int heads, void* out, cudaStream_t st) {
  using Cfg = AttnCfg<D>;
  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  const uint64_t C = static_cast<uint64_t>(heads) * D;
  uint64_t dims[5] = {static_cast<uint64_t>(D), static_cast<uint64_t>(heads), 3,
                      static_cast<uint64_t>(ntok), static_cast<uint64_t>(nb)};
  uint64_t strides[4] = {static_cast<uint64_t>(D) * 2, C * 2, 3 * C * 2, 3 * C * 2 * ntok};
  uint32_t box_q[5] = {64, 1, 1, 128, 1};
  uint32_t box_kv[5] = {64, 1, 1, static_cast<uint32_t>(Cfg::BKV), 1};
  if (int rc = encode_tmap_bf16(&kp.map_q, qkv, 5, dims, strides, box_q)) return rc;
  if (int rc = encode_tmap_bf16(&kp.map_kv, qkv, 5, dims, strides, box_kv)) return rc;
  kp.out = reinterpret_cast<__nv_bfloat16*>(out);
  kp.nb = nb;
  kp.ntok = ntok;
  kp.heads = heads;
  kp.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(D));
  static bool configured = false;
  if (!configured) {
    LDM_CUDA(cudaFuncSetAttribute(attn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = nb * heads * ((ntok + 127) / 128);
  launch_kernel(attn_kernel<D>, dim3(grid), dim3(kAttnThreads), Cfg::kSmemBytes, st, kp);
  return check_launch("attn_kernel");
}

}  // namespace ldm

using namespace ldm;

extern "C" int ldmseg_attention(const void* qkv, int nb, int ntok, int heads, int d, void* out,
                                void* stream) {
  LDM_REQUIRE(qkv && out && nb > 0 && ntok > 0 && heads > 0, "attention: bad arguments");
  LDM_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "attention: qkv not 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (d) {
    case 40: return launch_attn<40>(qkv, nb, ntok, heads, out, st);
    case 80: return launch_attn<80>(qkv, nb, ntok, heads, out, st);
    case 160: return launch_attn<160>(qkv, nb, ntok, heads, out, st);
    default: set_error("attention: unsupported head dim %d (40, 80, 160)", d); return -2;
  }
}

extern "C" int ldmseg_attention_simple(const void* qkv, int nb, int ntok, int heads, int d, void* out,
                                       void* stream) {
  LDM_REQUIRE(qkv && out && nb > 0 && ntok > 0 && heads > 0, "attention_simple: bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int threads = 128;
  const int grid = nb * heads * ((ntok + threads - 1) / threads);
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  switch (d) {
    case 40: launch_kernel(attn_simple_kernel<40>, dim3(grid), dim3(threads), 0, st, q, nb, ntok, heads, o); break;
    case 80: launch_kernel(attn_simple_kernel<80>, dim3(grid), dim3(threads), 0, st, q, nb, ntok, heads, o); break;
    case 160: launch_kernel(attn_simple_kernel<160>, dim3(grid), dim3(threads), 0, st, q, nb, ntok, heads, o); break;
    default: set_error("attention_simple: unsupported head dim %d", d); return -2;
  }
  return check_launch("attn_simple_kernel");
}
