// Self-attention forward (non-causal) on tcgen05: S = Q K^T into TMEM, online softmax in
// registers (exp2, fp32 statistics), P (bf16) back into TMEM as the A operand of O += P V (tcgen05.mma with A
// from tensor memory), O accumulated in TMEM.  (PT = false keeps the older form with P staged in shared memory.)
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
// CTA = 8 HS softmax warps (query tile 0, then 1; per tile the column halves, per half the four TMEM lane quadrants)
//       + MMA issuer of tile 0 (allocates TMEM) | MMA issuer of tile 1 | TMA producer | idle warp      (see AttnRoles)
// TMEM: S0 at [0,BKV), S1 at [BKV,2BKV), O0 at [2BKV, 2BKV+dk), O1 at [2BKV+dk, 2BKV+2dk), then P0, P1 (BKV/2
// columns each: two bf16 per 32-bit cell, K-contiguous -- the layout tcgen05.mma expects of a K-major A operand in
// tensor memory)  (<= 512 columns: 480 / 352 / 512 at d = 40 / 80 / 160).
//
// Why P lives in TMEM: with both operands in shared memory the tensor core fetches A (128 x 16 bf16 = 4 KB) and B for
// every K = 16 instruction at ~74 B/clk, so P.V (N = d <= 64, eight instructions per 128 keys) cost ~75 cycles per
// instruction against a 24-32 cycle math floor, and the P tile crossed shared memory twice (st.shared + operand
// fetch).  The MMA side, not the exponentials, was what the softmax warps waited for (ncu: tensor 24 %, MUFU 50 %).
#include "common.h"
#include <cstdlib>
#include <cstring>
#include "ptx.cuh"
#include "../../include/ldmseg_b200.h"

namespace ldm {

// Warp roles: softmax warpgroups first, then a control warpgroup with ONE MMA-issuing warp PER query tile, the TMA
// producer and an idle warp.  Per tile the order of the tensor-core work is fixed -- S(j+1) becomes issuable (S(j)
// copied to registers) before P(j) is ready -- so each issuer runs a plain blocking sequence of mbarrier waits.  A
// single issuer serving both tiles had to poll four barriers in an event loop; with every stage of the pipeline removed
// but the hand-shakes the kernel still ran at half its full time (tools/attn_timing.py, ablation builds).
//
// HS = column halves per score row.  HS = 1: a thread owns a whole row of its query tile (two softmax warpgroups, 232
// registers each after setmaxnreg).  HS = 2 (the 128-key tiles of d = 40): a row is shared by two threads of two
// different warps, 64 columns each -- four softmax warps per scheduler instead of two.  The softmax is a chain of
// in-order latencies (TMEM load -> row maximum -> scale -> MUFU -> sum -> pack -> TMEM store, ~600 instructions per
// 128-key tile in ~2 800 cycles with two warps per scheduler); what it lacked was independent warps to issue from,
// not pipe throughput.  The two halves agree on the running maximum through shared memory (one named barrier of 64
// threads per iteration); each keeps its own partial row sum.
// NT = query tiles per CTA.  NT = 2 shares every K/V tile between two query tiles and lets the tensor core work for one
// tile while the other exponentiates; NT = 1 halves the work per CTA and doubles the CTA count -- for the launches that
// cannot fill the machine otherwise (batch 1: 1024 tokens = 32 CTAs of two tiles, 256 tokens = 8), where the kernel's
// time is the length of one CTA's serial K/V loop.
template <int HS, int NT = 2>
struct AttnRoles {
  static constexpr int kSoftmaxWarps = 4 * HS * NT;
  static constexpr int kCtrlWarp0 = 4 * HS * NT;   // MMA issuers kCtrlWarp0 (+1 for the second tile; +0 allocates TMEM), TMA producer +2, idle
  static constexpr int kThreads = (4 * HS * NT + 4) * 32;
  static constexpr bool kResplit = NT == 2;        // NT = 1: 256 threads, up to 255 registers each without setmaxnreg
  // setmaxnreg (aligned groups of four warps): the control warpgroup gives its registers to the softmax warps
  // (the pool a warp can grow from holds only what other warps of the CTA released: 384 threads start at 168
  // registers, 128 x (168 - 40) = 256 x (232 - 168); 640 threads start at 96, 128 x (96 - 32) = 512 x (112 - 96))
  static constexpr int kCtrlRegs = HS == 2 ? 32 : 40;
  static constexpr int kSoftmaxRegs = HS == 2 ? 112 : 232;
};

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
  // + barriers, + 6 KB of row-maximum / row-sum exchange between the column halves of a row (HS = 2)
  // (the P staging only when P goes through shared memory)
  static constexpr int smem_bytes(bool p_in_tmem) {
    return 2 * kQBytes + kKVStages * 2 * kKBytes + (p_in_tmem ? 0 : 2 * kPBytes) + 24 * 8 + 6144 + 64 + 1024;
  }
  static constexpr int kTmemCols = 512;
  static constexpr int kOCol = 2 * BKV;   // S0 at [0,BKV), S1 at [BKV,2BKV), then O0, O1 (dk columns each)
  static constexpr int kPCol = kOCol + 2 * kDK;   // P0, P1: BKV/2 columns each (bf16 pairs)
  static_assert(kPCol + BKV <= 512, "TMEM budget");
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
// -DLDMSEG_ATTN_ABLATE=<mask> builds (timing experiments, results are garbage): bit0 no exponentials, bit1 S row read
// from TMEM only once, bit2 no P.V products, bit3 no S products, bit4 no P store, bit5 no row maximum, bit6 no K/V loads after the first ring fill

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

// -DLDMSEG_ATTN_TIMING builds: CTA 0 accumulates clock() deltas per phase (softmax warp 2 -> [0..7], warp 6 ->
// [8..15], MMA thread -> [16..23]); read back with ldmseg_attn_timing_read.
#ifdef LDMSEG_ATTN_TIMING
__device__ unsigned int g_attn_timing[32];
#define ATT_TICK(acc_i)                      \
  do {                                       \
    const unsigned int now_ = clock();       \
    tacc[acc_i] += now_ - tlast;             \
    tlast = now_;                            \
  } while (0)
#else
#define ATT_TICK(acc_i) do {} while (0)
#endif

// PM: share of the exponentials evaluated on the FMA pipe (0 = none, 1 = half, 2 = a quarter)
// PT: P goes to tensor memory (A operand from TMEM) instead of shared memory
template <int D, int PM, bool PT, int HS, int NT>
__global__ void __launch_bounds__(AttnRoles<HS, NT>::kThreads, 1) attn_kernel(const __grid_constant__ AttnKParams p) {
  using Cfg = AttnCfg<D>;
  using Roles = AttnRoles<HS, NT>;
  static_assert(HS == 1 || (PT && Cfg::BKV % 128 == 0), "column halves: P in tensor memory, 128-key tiles");
  static_assert(NT == 2 || (NT == 1 && HS == 1 && PT), "one query tile per CTA: one thread per row, P in tensor memory");
  constexpr int kSoftmaxWarp0 = 0, kCtrlWarp0 = Roles::kCtrlWarp0, kMmaWarp0 = Roles::kCtrlWarp0,
                kTmaWarp = Roles::kCtrlWarp0 + 2;
  constexpr int kCtrlRegs = Roles::kCtrlRegs, kSoftmaxRegs = Roles::kSoftmaxRegs;
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_p + (PT ? 0 : 2 * Cfg::kPBytes));
  uint64_t* q_full = bars + 0;
  uint64_t* s_full = bars + 1;    // [2] per query tile
  uint64_t* p_full = bars + 3;    // [2]
  uint64_t* pv_done = bars + 5;   // [2]
  uint64_t* s_free = bars + 7;    // [2] S_t has been copied to registers
  uint64_t* kv_full = bars + 9;   // [stages]
  uint64_t* kv_empty = bars + 9 + KS;  // [stages]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 9 + 2 * KS);
  // HS = 2: [iteration parity][tile][half][row] row maxima, then [tile][half][row] row sums
  float* xch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem) + 4 + 15) & ~static_cast<uintptr_t>(15));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef LDMSEG_ATTN_ABLATE
  constexpr int dbg = LDMSEG_ATTN_ABLATE;   // compile-time mask: a run-time one costs the d = 40 kernel 320 B of spills
#else
  constexpr int dbg = 0;
#endif

  const int q_blocks = (p.ntok + 128 * NT - 1) / (128 * NT);
  const int qb = blockIdx.x % q_blocks;
  const int head = (blockIdx.x / q_blocks) % p.heads;
  const int b = blockIdx.x / (q_blocks * p.heads);
  const int q0 = qb * 128 * NT;
  const int T = (p.ntok_kv + BKV - 1) / BKV;

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&p.map_q);
    tma_prefetch_desc(&p.map_kv);
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], NT);   // every tile's issuer releases a stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128 * HS);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 128 * HS);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp0) {
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

  if constexpr (Roles::kResplit) {
    if (warp >= kCtrlWarp0 && warp < kCtrlWarp0 + 4) setmaxnreg_dec<kCtrlRegs>();
  }
  if (warp == kTmaWarp) {
    if (elect_one()) {
      mbar_expect_tx(q_full, NT * Cfg::kQBytes);
      for (int t = 0; t < NT; ++t)
        for (int pn = 0; pn < kPanels; ++pn)
          tma_load_5d(sm_q + t * Cfg::kQBytes + pn * 128 * 128, &p.map_q, q_full, pn * 64, head, 0,
                      q0 + t * 128, b);
      for (int j = 0; j < T; ++j) {
        const int st = j % KS;
        const uint32_t n = static_cast<uint32_t>(j / KS);
        mbar_wait(&kv_empty[st], (n & 1) ^ 1);
        if ((dbg & 64) && j >= KS) {   // ablation: stage contents stay stale
          mbar_arrive(&kv_full[st]);
          continue;
        }
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
  } else if (warp >= kMmaWarp0 && warp < kMmaWarp0 + NT) {
    if (elect_one()) {
      const int t = warp - kMmaWarp0;   // the query tile this issuer serves
      constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, kDK, 0, 1);
      auto issue_s = [&](int j) {  // S_t = Q_t K_j^T
        const int st = j % KS;
        const uint32_t d_tmem = tmem_base + t * BKV;
        if ((dbg & 8) && j > 0) return;
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
      auto issue_pv = [&](int j) {  // O_t += P_t V_j
        const int st = j % KS;
        const uint32_t o_tmem = tmem_base + Cfg::kOCol + t * kDK;
        if ((dbg & 4) && j > 0) return;
#pragma unroll
        for (int ks = 0; ks < BKV / 16; ++ks) {
          const int pn = ks >> 2, kk = ks & 3;
          // V tile: rows = kv (128 B each), MN(d)-major; 16 kv rows per k-step = 2048 B
          const uint64_t bd =
              make_smem_desc_sw128(smem_u32(sm_v + st * Cfg::kKBytes + ks * 2048), BKV * 128, 1024);
          if constexpr (PT) {
            // A = P_t from tensor memory: 16 keys = 8 columns of bf16 pairs per k-step
            umma_bf16_ts(o_tmem, tmem_base + Cfg::kPCol + t * (BKV / 2) + ks * 8, bd, idesc_pv,
                         (j > 0 || ks > 0) ? 1u : 0u);
          } else {
            const uint64_t a =
                make_smem_desc_sw128(smem_u32(sm_p + t * Cfg::kPBytes + pn * 128 * 128), 16, 1024) + 2 * kk;
            umma_bf16(o_tmem, a, bd, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
          }
        }
      };
#ifdef LDMSEG_ATTN_TIMING
      unsigned int tacc[4] = {0, 0, 0, 0}, tlast = clock();
#endif
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      umma_commit(&s_full[t]);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) {
          // S_t(j+1): needs the warpgroup to have copied S_t(j) to registers, and K tile j+1 in shared memory
          mbar_wait(&s_free[t], static_cast<uint32_t>(j) & 1);
          mbar_wait(&kv_full[(j + 1) % KS], static_cast<uint32_t>((j + 1) / KS) & 1);
          tc_fence_after();
          ATT_TICK(0);   // waiting
          issue_s(j + 1);
          umma_commit(&s_full[t]);
          ATT_TICK(1);
        }
        // O_t += P_t(j) V_j
        mbar_wait(&p_full[t], static_cast<uint32_t>(j) & 1);
        tc_fence_after();
        ATT_TICK(0);
        issue_pv(j);
        umma_commit(&pv_done[t]);
        // K/V stage j can be refilled once both tiles' products that read it have completed (this issuer's S_t(j)
        // and P_t(j).V_j precede the commit; the barrier counts both issuers)
        umma_commit(&kv_empty[j % KS]);
        ATT_TICK(2);
      }
#ifdef LDMSEG_ATTN_TIMING
      if (blockIdx.x == 0)
        for (int i = 0; i < 3; ++i) g_attn_timing[16 + 4 * t + i] = tacc[i];
#endif
    }
    __syncwarp();
  } else if (warp >= kSoftmaxWarp0 && warp < kSoftmaxWarp0 + Roles::kSoftmaxWarps) {
    // ---------------------------------------------------------------- softmax / correction
    if constexpr (Roles::kResplit) setmaxnreg_inc<kSoftmaxRegs>();
    constexpr int NC = BKV / HS;                      // score columns per thread
    const int t = (warp - kSoftmaxWarp0) / (4 * HS);  // query tile
    const int h = ((warp - kSoftmaxWarp0) >> 2) % HS; // column half of the row
    const int q = warp & 3;                           // TMEM lane quadrant
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + t * BKV + h * NC;
    const uint32_t o_addr = tmem_base + lane_off + Cfg::kOCol + t * kDK;
    // the O columns this thread corrects / writes out: its half of the (padded) head dimension, in groups of 8
    constexpr int kOGroups = kDK / 8;
    const int c_lo = 8 * (h * kOGroups / HS), c_hi = 8 * ((h + 1) * kOGroups / HS);
    const int xbar = 2 + t * 4 + q;                   // named barrier shared by the two threads' warps (HS = 2)
    const uint32_t my_p_u32 = smem_u32(sm_p + t * Cfg::kPBytes);
    const float scale = p.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
#ifdef LDMSEG_ATTN_TIMING
    unsigned int tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock();
#endif
    for (int j = 0; j < T; ++j) {
      mbar_wait(&s_full[t], static_cast<uint32_t>(j) & 1);
      tc_fence_after();
      ATT_TICK(0);   // wait for S
      // the whole score row goes to registers in one pass; S_t in TMEM is then free for the next product
      uint32_t sr[NC];
      if (!(dbg & 2) || j == 0) {
#pragma unroll
        for (int c = 0; c < NC; c += 32) tmem_ld_32x32(s_addr + c, *reinterpret_cast<uint32_t(*)[32]>(&sr[c]));
        tmem_wait_ld();
      }
      tc_fence_before();
      mbar_arrive(&s_free[t]);
      ATT_TICK(1);   // S row -> registers
      const int kv_valid = p.ntok_kv - j * BKV - h * NC;  // my columns < kv_valid are real tokens
      if (kv_valid < NC) {
#pragma unroll
        for (int i = 0; i < NC; ++i)
          if (i >= kv_valid) sr[i] = 0xff800000u;  // -inf
      }
      // row maximum (3-input max, four independent chains)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < NC; i += 8) {
        mx0 = max3f(mx0, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
        mx1 = max3f(mx1, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
        mx2 = max3f(mx2, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
        mx3 = max3f(mx3, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
      }
      float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale;
      if (dbg & 32) mx = __uint_as_float(sr[0]) * scale;
      if constexpr (HS == 2) {
        // both halves of the row must take the same decision about the running maximum.  Slots alternate with the
        // iteration parity: a slot is rewritten two iterations later, after a barrier both threads have passed since
        // the partner read it
        float* mine = xch + (((j & 1) * 2 + t) * 2 + h) * 128 + row;
        *mine = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(xbar) : "memory");
        mx = fmaxf(mx, mine[(h ? -128 : 128)]);
      }
      // lazy rescaling: the running maximum only moves when the new one exceeds it by more than 2^8, so P
      // stays <= 256 (exact in bf16's range, fp32 accumulation) and O is corrected a handful of times per row
      float alpha = 1.f;
      if (mx > m_run + kRescaleLog2) {
        alpha = ex2_approx(m_run - mx);
        m_run = mx;
      }
      const float negm = -m_run;
      ATT_TICK(2);   // row maximum
      // p = exp2(s*scale - m): packed FFMA2, one MUFU each, packed FADD2 row sums, bf16 pairs
      float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
      uint32_t pk[NC / 2];
#pragma unroll
      for (int i = 0; i < NC; i += 4) {
        float2 a = fma2(make_float2(__uint_as_float(sr[i]), __uint_as_float(sr[i + 1])), scale, negm);
        float2 b = fma2(make_float2(__uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3])), scale, negm);
        if (!(dbg & 1)) {
          a.x = ex2_approx(a.x); a.y = ex2_approx(a.y);
          if (PM == 1 || (PM == 2 && ((i >> 2) & 1))) {
            b = exp2_poly2(b);
          } else {
            b.x = ex2_approx(b.x); b.y = ex2_approx(b.y);
          }
        }
        rs0 = add2(rs0, a);
        rs1 = add2(rs1, b);
        pk[i / 2] = pack_bf16x2(a.x, a.y);
        pk[i / 2 + 1] = pack_bf16x2(b.x, b.y);
      }
      l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
      ATT_TICK(4);   // exponentials, row sums, packing
      // the P buffer and O are free once P.V of the previous tile has completed
      if (j > 0) {
        mbar_wait(&pv_done[t], static_cast<uint32_t>(j - 1) & 1);
        tc_fence_after();
      }
      ATT_TICK(5);   // wait for P.V
      if (dbg & 16) {
      } else if constexpr (PT) {
        // thread = query row = TMEM lane: its BKV/2 packed pairs go to consecutive columns of P_t
        const uint32_t p_addr = tmem_base + lane_off + Cfg::kPCol + t * (BKV / 2) + h * (NC / 2);
#pragma unroll
        for (int c = 0; c < NC / 2; c += 32) tmem_st_32x32(p_addr + c, *reinterpret_cast<uint32_t(*)[32]>(&pk[c]));
      } else {
#pragma unroll
        for (int c8 = 0; c8 < BKV / 8; ++c8) {
          const uint32_t prow = my_p_u32 + (c8 >> 3) * (128 * 128) + row * 128;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (((c8 & 7) ^ (row & 7)) << 4)),
                       "r"(pk[4 * c8]), "r"(pk[4 * c8 + 1]), "r"(pk[4 * c8 + 2]), "r"(pk[4 * c8 + 3])
                       : "memory");
        }
      }
      // correction: O *= alpha (skipped warp-uniformly when no row of this warp moved its max)
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 8) {
          uint32_t r[8];
          tmem_ld_32x8(o_addr + c, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_32x8(o_addr + c, r);
        }
      }
      if constexpr (PT) tmem_wait_st();        // P (and the corrected O) are in tensor memory before the MMA reads them
      else {
        tmem_wait_st();
        fence_proxy_async_smem();
      }
      tc_fence_before();
      mbar_arrive(&p_full[t]);
      ATT_TICK(6);   // P store (+ correction)
    }
#ifdef LDMSEG_ATTN_TIMING
    if (blockIdx.x == 0 && lane == 0 && q == 0 && h == 0)
      for (int i = 0; i < 8; ++i) g_attn_timing[t * 8 + i] = tacc[i];
#endif
    // ---------------------------------------------------------------- final normalisation
    mbar_wait(&pv_done[t], static_cast<uint32_t>(T - 1) & 1);
    tc_fence_after();
    if constexpr (HS == 2) {   // the row sum is the two halves' partial sums (same scaling: same running maximum)
      float* mine = xch + 4 * 128 * 2 + (t * 2 + h) * 128 + row;
      *mine = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(xbar) : "memory");
      l_run += mine[(h ? -128 : 128)];
    }
    const float inv_l = 1.f / l_run;
    const int tok = q0 + t * 128 + row;
    __nv_bfloat16* dst = p.out + (static_cast<size_t>(b) * p.ntok + tok) * (p.heads * D) + head * D;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 8) {
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
  if (warp == kMmaWarp0) {
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
template <int D, int PM, bool PT, int HS, int NT>
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
    LDM_CUDA(cudaFuncSetAttribute(attn_kernel<D, PM, PT, HS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg::smem_bytes(PT)));
    configured = true;
  }
  const int grid = nb * heads * ((ntok + 128 * NT - 1) / (128 * NT));
  launch_kernel(attn_kernel<D, PM, PT, HS, NT>, dim3(grid), dim3(AttnRoles<HS, NT>::kThreads), Cfg::smem_bytes(PT), st,
                kp);
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
  static int pt = -1;   // LDMSEG_ATTN_PTMEM=0: P through shared memory (the older form; A/B timing)
  if (pt < 0) {
    const char* e = getenv("LDMSEG_ATTN_PTMEM");
    pt = e ? atoi(e) : 1;
  }
  static int hs = -1;   // LDMSEG_ATTN_HALVES=1: one thread per score row also for the 128-key tiles (A/B timing)
  if (hs < 0) {
    const char* e = getenv("LDMSEG_ATTN_HALVES");
    hs = e ? atoi(e) : 2;
  }
  static int nt1 = -1;  // LDMSEG_ATTN_SINGLE=0: always two query tiles per CTA (A/B timing)
  if (nt1 < 0) {
    const char* e = getenv("LDMSEG_ATTN_SINGLE");
    nt1 = e ? atoi(e) : 1;
  }
#define LDM_ATTN_CASE(PMV, PTV, HSV, NTV) \
  return launch_attn_pm<D, PMV, PTV, HSV, NTV>(q_src, q_which_n, kv_src, kv_which_n, k_which, v_which, nb, ntok, ntok_kv, heads, out, st)
  if constexpr (AttnCfg<D>::BKV == 128) {
    if (pt && hs == 2) {
      switch (pm) {
        case 0: LDM_ATTN_CASE(0, true, 2, 2);
        case 1: LDM_ATTN_CASE(1, true, 2, 2);
        default: LDM_ATTN_CASE(2, true, 2, 2);
      }
    }
  } else {
    // a launch of two-tile CTAs that leaves most SMs idle: one tile per CTA, twice the CTAs, half the serial loop
    if (pt && nt1 && ntok > 128 && static_cast<long long>(nb) * heads * ((ntok + 255) / 256) <= num_sms() / 2) {
      switch (pm) {
        case 0: LDM_ATTN_CASE(0, true, 1, 1);
        case 1: LDM_ATTN_CASE(1, true, 1, 1);
        default: LDM_ATTN_CASE(2, true, 1, 1);
      }
    }
  }
  if (pt) {
    switch (pm) {
      case 0: LDM_ATTN_CASE(0, true, 1, 2);
      case 1: LDM_ATTN_CASE(1, true, 1, 2);
      default: LDM_ATTN_CASE(2, true, 1, 2);
    }
  }
  switch (pm) {
    case 0: LDM_ATTN_CASE(0, false, 1, 2);
    case 1: LDM_ATTN_CASE(1, false, 1, 2);
    default: LDM_ATTN_CASE(2, false, 1, 2);
  }
#undef LDM_ATTN_CASE
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

#ifdef LDMSEG_ATTN_TIMING
extern "C" int ldmseg_attn_timing_read(unsigned int* dst32) {
  return static_cast<int>(cudaMemcpyFromSymbol(dst32, g_attn_timing, sizeof(unsigned int) * 32));
}
#endif

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
