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
// The softmax (one MUFU.EX2 per score) is the bottleneck at these small head dims, not the MMAs, so
// the CTA is built to keep the MUFU pipe busy: it owns TWO 128-query tiles of one (image, head), each
// with its own softmax warpgroup; while one warpgroup exponentiates S_j the tensor core computes the
// other tile's S and P.V (ping-pong), and K/V tiles are loaded once for both.
//
// CTA = 320 threads:  warp 0 TMA producer | warp 1 MMA issuer + TMEM allocator |
//                     warps 2..5 softmax/correction of query tile 0 | warps 6..9 of query tile 1
// TMEM: S0 at [0,BKV), S1 at [BKV,2BKV), O0 at [2BKV, 2BKV+dk), O1 at [2BKV+dk, 2BKV+2dk)  (<= 448 columns).
#include "common.h"
#include <cstdlib>
#include <cstring>
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

constexpr int kAttnThreads = 320;

template <int D>
struct AttnCfg {
  static constexpr int BKV = (D <= 40) ? 128 : 64;
  static constexpr int kPanels = (D + 63) / 64;
  static constexpr int kDK = (D + 15) / 16 * 16;
  static constexpr int kQBytes = kPanels * 128 * 128;       // one query tile
  static constexpr int kKBytes = kPanels * BKV * 128;       // one K (or V) tile
  static constexpr int kPBytes = (BKV / 64) * 128 * 128;    // one P tile
  // K/V ring depth: a K/V tile is only reloaded after both P.V products that read it have completed, and the
  // next S product needs the tile after that one, so two stages expose the whole TMA latency every iteration
  static constexpr int kKVStages = (D <= 80) ? 4 : 2;
  static constexpr int kSmemBytes = 2 * kQBytes + kKVStages * 2 * kKBytes + 2 * kPBytes + 24 * 8 + 1024;
  static constexpr int kTmemCols = 512;
  static constexpr int kOCol = 2 * BKV;   // S0 at [0,BKV), S1 at [BKV,2BKV), then O0, O1 (dk columns each)
};

struct alignas(64) AttnKParams {
  CUtensorMap map_q;
  CUtensorMap map_kv;
  __nv_bfloat16* out;
  int nb, ntok, heads;
  int ntok_kv;       // number of key / value tokens (= ntok for self-attention)
  int k_which, v_which;  // index of K / V along the "which" dimension of map_kv (1, 2 in a fused QKV; 0, 1 in a KV tensor)
  float scale_log2;  // (1/sqrt(d)) * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// packed fp32x2 arithmetic (FFMA2 / FADD2): half the issue slots of the scalar forms
__device__ __forceinline__ float2 fma2(float2 a, float s, float c) {
  uint64_t x, y, z, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(s));
  asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(c));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t x, y, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(x), "l"(y));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
__device__ __forceinline__ float2 fma2v(float2 a, float2 b, float2 c) {
  uint64_t x, y, z, r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
  float2 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(r));
  return o;
}
// exp2 on the FMA pipe for a share of the scores (the MUFU pipe, 4 lanes per clock per sub-partition, is what
// bounds the softmax): round-to-nearest split x = n + f, |f| <= 0.5, degree-3 minimax polynomial for 2^f
// (max relative error 7.5e-5, 25x below the bf16 rounding of P), exponent patched in with one integer add.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 t = add2(x, magic);
  const float2 n = add2(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = add2(x, make_float2(-n.x, -n.y));
  float2 q = fma2v(f, make_float2(0.0551716648f, 0.0551716648f), make_float2(0.2426111251f, 0.2426111251f));
  q = fma2v(q, f, make_float2(0.6932609677f, 0.6932609677f));
  q = fma2v(q, f, make_float2(0.9999280572f, 0.9999280572f));
  return make_float2(__uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t.x) << 23)),
                     __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t.y) << 23)));
}
constexpr float kRescaleLog2 = 8.f;

// PM: share of the exponentials evaluated on the FMA pipe (0 = none, 1 = half, 2 = a quarter)
template <int D, int PM>
__global__ void __launch_bounds__(kAttnThreads, 1) attn_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<D>;
  constexpr int BKV = Cfg::BKV;
  constexpr int kPanels = Cfg::kPanels;
  constexpr int kDK = Cfg::kDK;
  constexpr int KS = Cfg::kKVStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sm_q = smem;                              // [2 query tiles][Q]
  uint8_t* sm_k = sm_q + 2 * Cfg::kQBytes;           // [2 stages][K]
  uint8_t* sm_v = sm_k + KS * Cfg::kKBytes;          // [stages][V]
  uint8_t* sm_p = sm_v + KS * Cfg::kKBytes;          // [2 query tiles][P]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + 2 * Cfg::kPBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // [2] per query tile
  uint64_t* p_full = bars + 3;    // [2]
  uint64_t* pv_done = bars + 5;   // [2]
  uint64_t* s_free = bars + 7;    // [2] S_t has been copied to registers
  uint64_t* kv_full = bars + 9;   // [stages]
  uint64_t* kv_empty = bars + 9 + KS;  // [stages]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9 + 2 * KS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int q_blocks = (p.ntok + 255) / 256;
  const int qb = blockIdx.x % q_blocks;
  const int head = (blockIdx.x / q_blocks) % p.heads;
  const int b = blockIdx.x / (q_blocks * p.heads);
  const int q0 = qb * 256;
  const int T = (p.ntok_kv + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 128);
    }
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
  // Programmatic dependent launch.  Trigger only now that this CTA owns its TMEM columns: a dependent CTA that
  // became co-resident earlier could otherwise take them and starve this (prerequisite) grid forever.
  pdl_trigger();
  pdl_wait();  // the prologue above overlapped the previous kernel; inputs are read from here on

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, 2 * Cfg::kQBytes);
      for (int t = 0; t < 2; ++t)
        for (int pn = 0; pn < kPanels; ++pn)
          tma_load_5d(sm_q + t * Cfg::kQBytes + pn * 128 * 128, &p.map_q, q_full, pn * 64, head, 0,
                      q0 + t * 128, b);
      for (int j = 0; j < T; ++j) {
        const int st = j % KS;
        const uint32_t n = static_cast<uint32_t>(j / KS);
        mbar_wait(&kv_empty[st], (n & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * Cfg::kKBytes);
        for (int pn = 0; pn < kPanels; ++pn) {
          tma_load_5d(sm_k + st * Cfg::kKBytes + pn * BKV * 128, &p.map_kv, &kv_full[st], pn * 64,
                      head, p.k_which, j * BKV, b);
          tma_load_5d(sm_v + st * Cfg::kKBytes + pn * BKV * 128, &p.map_kv, &kv_full[st], pn * 64,
                      head, p.v_which, j * BKV, b);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kDK, 0, 1);
      auto issue_s = [&](int j, int t) {  // S_t = Q_t K_j^T
        const int st = j % KS;
        const uint32_t d_tmem = tmem_base + t * BKV;
#pragma unroll
        for (int ks = 0; ks < kDK / 16; ++ks) {
          const int pn = ks >> 2, kk = ks & 3;
          const uint64_t a =
              make_smem_desc_sw128(smem_u32(sm_q + t * Cfg::kQBytes + pn * 128 * 128), 16, 1024) + 2 * kk;
          const uint64_t bd =
              make_smem_desc_sw128(smem_u32(sm_k + st * Cfg::kKBytes + pn * BKV * 128), 16, 1024) + 2 * kk;
          umma_bf16(d_tmem, a, bd, idesc_s, ks > 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int j, int t) {  // O_t += P_t V_j
        const int st = j % KS;
        const uint32_t o_tmem = tmem_base + Cfg::kOCol + t * kDK;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const int pn = ks >> 2, kk = ks & 3;
          const uint64_t a =
              make_smem_desc_sw128(smem_u32(sm_p + t * Cfg::kPBytes + pn * 128 * 128), 16, 1024) + 2 * kk;
          // V tile: rows = kv (128 B each), MN(d)-major; 16 kv rows per k-step = 2048 B
          const uint64_t bd =
              make_smem_desc_sw128(smem_u32(sm_v + st * Cfg::kKBytes + ks * 2048), BKV * 128, 1024);
          umma_bf16(o_tmem, a, bd, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      umma_commit(&s_full[0]);
      issue_s(0, 1);
      umma_commit(&s_full[1]);
      // Event loop: the two query tiles advance independently.  Issuing in a fixed order (wait S-free of tile 0,
      // wait P of tile 0, then tile 1, ...) made each softmax warpgroup wait for the other one's exponentials
      // before its next S product was even issued.
      int js[2] = {1, 1};   // next S product to issue per query tile (S_0 is already in flight)
      int jp[2] = {0, 0};   // next P.V product to issue per query tile
      while (jp[0] < T || jp[1] < T) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          // S_t(j): needs the warpgroup to have copied S_t(j-1) to registers, and K tile j in shared memory
          if (js[t] < T && mbar_test_wait(&s_free[t], static_cast<uint32_t>(js[t] - 1) & 1) &&
              mbar_test_wait(&kv_full[js[t] % KS], static_cast<uint32_t>(js[t] / KS) & 1)) {
            tc_fence_after();
            issue_s(js[t], t);
            umma_commit(&s_full[t]);
            ++js[t];
          }
          // O_t += P_t(j) V_j
          if (jp[t] < T && mbar_test_wait(&p_full[t], static_cast<uint32_t>(jp[t]) & 1)) {
            tc_fence_after();
            const int j = jp[t];
            issue_pv(j, t);
            umma_commit(&pv_done[t]);
            ++jp[t];
            // K/V stage j&1 can be refilled once both P.V products of tile j are queued (the S products that
            // read its K half were issued before the P tiles they lead to)
            if (jp[t ^ 1] > j) umma_commit(&kv_empty[j % KS]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax / correction
    const int t = (warp - 2) >> 2;  // query tile of this warpgroup
    const int q = warp & 3;         // TMEM lane quadrant
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + t * BKV;
    const uint32_t o_addr = tmem_base + lane_off + Cfg::kOCol + t * kDK;
    const uint32_t my_p_u32 = smem_u32(sm_p + t * Cfg::kPBytes);
    const float scale = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[t], static_cast<uint32_t>(j) & 1);
      tc_fence_after();
      // the whole score row goes to registers in one pass; S_t in TMEM is then free for the next product
      uint32_t sr[BKV];
#pragma unroll
      for (int c = 0; c < BKV; c += 32) tmem_ld_32x32(s_addr + c, *reinterpret_cast<uint32_t(*)[32]>(&sr[c]));
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      const int kv_valid = p.ntok_kv - j * BKV;  // columns < kv_valid are real tokens
      if (kv_valid < BKV) {
#pragma unroll
        for (int i = 0; i < BKV; ++i)
          if (i >= kv_valid) sr[i] = 0xff800000u;  // -inf
      }
      // row maximum (3-input max, four independent chains)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < BKV; i += 8) {
        mx0 = max3f(mx0, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
        mx1 = max3f(mx1, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
        mx2 = max3f(mx2, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
        mx3 = max3f(mx3, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale;
      // lazy rescaling: the running maximum only moves when the new one exceeds it by more than 2^8, so P
      // stays <= 256 (exact in bf16's range, fp32 accumulation) and O is corrected a handful of times per row
      float alpha = 1.f;
      if (mx > m_run + kRescaleLog2) {
        alpha = ex2_approx(m_run - mx);
        m_run = mx;
      }
      const float negm = -m_run;
      // p = exp2(s*scale - m): packed FFMA2, one MUFU each, packed FADD2 row sums, bf16 pairs
      float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
      uint32_t pk[BKV / 2];
#pragma unroll
      for (int i = 0; i < BKV; i += 4) {
        float2 a = fma2(make_float2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), scale, negm);
        float2 b = fma2(make_float2(__uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3])), scale, negm);
        a.x = ex2_approx(a.x); a.y = ex2_approx(a.y);
        if (PM == 1 || (PM == 2 && ((i >> 2) & 1))) {
          b = exp2_poly2(b);
        } else {
          b.x = ex2_approx(b.x); b.y = ex2_approx(b.y);
        }
        rs0 = add2(rs0, a);
        rs1 = add2(rs1, b);
        pk[i / 2] = pack_bf16x2(a.x, a.y);
        pk[i / 2 + 1] = pack_bf16x2(b.x, b.y);
      }
      l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
      // the P buffer and O are free once P.V of the previous tile has completed
      if (j > 0) {
        mbar_wait(&pv_done[t], static_cast<uint32_t>(j - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int c8 = 0; c8 < BKV / 8; ++c8) {
        const uint32_t prow = my_p_u32 + (c8 >> 3) * (128 * 128) + row * 128;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (((c8 & 7) ^ (row & 7)) << 4)),
                     "r"(pk[4 * c8]), "r"(pk[4 * c8 + 1]), "r"(pk[4 * c8 + 2]), "r"(pk[4 * c8 + 3])
                     : "memory");
      }
      // correction: O *= alpha (skipped warp-uniformly when no row of this warp moved its max)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
        for (int c = 0; c < kDK; c += 16) {
          uint32_t r[16];
          tmem_ld_32x16(o_addr + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x8(o_addr + c, *reinterpret_cast<uint32_t(*)[8]>(&r[0]));
          tmem_st_32x8(o_addr + c + 8, *reinterpret_cast<uint32_t(*)[8]>(&r[8]));
        }
        tmem_wait_st();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }
    // ---------------------------------------------------------------- final normalisation
    mbar_wait(&pv_done[t], static_cast<uint32_t>(T - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const int tok = q0 + t * 128 + row;
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

// q_src: bf16 [nb*ntok, q_which_n * heads * D] (column block 0 = Q); kv_src: bf16 [nb*ntok_kv, kv_which_n * heads * D]
// with K / V in column blocks k_which / v_which.  Self-attention: q_src = kv_src = the fused QKV, (3, 1, 2).
template <int D, int PM>
static int launch_attn_pm(const void* q_src, int q_which_n, const void* kv_src, int kv_which_n, int k_which,
                          int v_which, int nb, int ntok, int ntok_kv, int heads, void* out, cudaStream_t st) {
  using Cfg = AttnCfg<D>;
  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  const uint64_t C = static_cast<uint64_t>(heads) * D;
  uint64_t dims_q[5] = {static_cast<uint64_t>(D), static_cast<uint64_t>(heads), static_cast<uint64_t>(q_which_n),
                        static_cast<uint64_t>(ntok), static_cast<uint64_t>(nb)};
  uint64_t str_q[4] = {static_cast<uint64_t>(D) * 2, C * 2, q_which_n * C * 2, q_which_n * C * 2 * ntok};
  uint64_t dims_kv[5] = {static_cast<uint64_t>(D), static_cast<uint64_t>(heads), static_cast<uint64_t>(kv_which_n),
                         static_cast<uint64_t>(ntok_kv), static_cast<uint64_t>(nb)};
  uint64_t str_kv[4] = {static_cast<uint64_t>(D) * 2, C * 2, kv_which_n * C * 2, kv_which_n * C * 2 * ntok_kv};
  uint32_t box_q[5] = {64, 1, 1, 128, 1};
  uint32_t box_kv[5] = {64, 1, 1, static_cast<uint32_t>(Cfg::BKV), 1};
  if (int rc = encode_tmap_bf16(&kp.map_q, q_src, 5, dims_q, str_q, box_q)) return rc;
  if (int rc = encode_tmap_bf16(&kp.map_kv, kv_src, 5, dims_kv, str_kv, box_kv)) return rc;
  kp.out = reinterpret_cast<__nv_bfloat16*>(out);
  kp.nb = nb;
  kp.ntok = ntok;
  kp.ntok_kv = ntok_kv;
  kp.k_which = k_which;
  kp.v_which = v_which;
  kp.heads = heads;
  kp.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(D));
  static bool configured = false;
  if (!configured) {
    LDM_CUDA(cudaFuncSetAttribute(attn_kernel<D, PM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::kSmemBytes));
    configured = true;
  }
  const int grid = nb * heads * ((ntok + 255) / 256);
  launch_kernel(attn_kernel<D, PM>, dim3(grid), dim3(kAttnThreads), Cfg::kSmemBytes, st, kp);
  return check_launch("attn_kernel");
}

template <int D>
static int launch_attn(const void* q_src, int q_which_n, const void* kv_src, int kv_which_n, int k_which, int v_which,
                       int nb, int ntok, int ntok_kv, int heads, void* out, cudaStream_t st) {
  static int pm = -1;
  if (pm < 0) {
    const char* e = getenv("LDMSEG_ATTN_POLY");
    pm = e ? atoi(e) : 2;
  }
  switch (pm) {
    case 0: return launch_attn_pm<D, 0>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st);
    case 1: return launch_attn_pm<D, 1>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st);
    default: return launch_attn_pm<D, 2>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st);
  }
}

static int attn_dispatch(const void* q_src, int q_which_n, const void* kv_src, int kv_which_n, int k_which,
                         int v_which, int nb, int ntok, int ntok_kv, int heads, int d, void* out, cudaStream_t st) {
  switch (d) {
    case 40: return launch_attn<40>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st);
    case 80: return launch_attn<80>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st);
    case 160: return launch_attn<160>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st);
    default: set_error("attention: unsupported head dim %d (40, 80, 160)", d); return -2;
  }
}

}  // namespace ldm

using namespace ldm;

extern "C" int ldmseg_attention(const void* qkv, int nb, int ntok, int heads, int d, void* out,
                                void* stream) {
  LDM_REQUIRE(qkv && out && nb > 0 && ntok > 0 && heads > 0, "attention: bad arguments");
  LDM_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0, "attention: qkv not 16-byte aligned");
  return attn_dispatch(qkv, 3, qkv, 3, 1, 2, nb, ntok, ntok, heads, d, out, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ldmseg_cross_attention(const void* q, const void* kv, int nb, int ntok_q, int ntok_kv, int heads,
                                      int d, void* out, void* stream) {
  LDM_REQUIRE(q && kv && out && nb > 0 && ntok_q > 0 && ntok_kv > 0 && heads > 0, "cross_attention: bad arguments");
  LDM_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0,
              "cross_attention: q / kv not 16-byte aligned");
  return attn_dispatch(q, 1, kv, 2, 0, 1, nb, ntok_q, ntok_kv, heads, d, out, reinterpret_cast<cudaStream_t>(stream));
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
