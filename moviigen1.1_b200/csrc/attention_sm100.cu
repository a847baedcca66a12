// Flash-attention forward for the DiT (head_dim 128, bf16, non-causal) on tcgen05 / TMEM / TMA.
// Replaces flash_attn.flash_attn_varlen_func (FA2 mma.sync kernels) called from
// wan/modules/attention.py:113-127 for self-attention (model.py:146-151) and cross-attention (:176).
//
// This file holds two kernels that share one CTA shape (one head x 256 query rows = two 128-row Q tiles, K/V streamed
// by TMA through a shared-memory ring, tcgen05.mma issued by one warp per Q tile, S / P / O in TMEM):
//
//   attention_fwd_k128_kernel    THE DEFAULT.  128-key steps, one 128-column score buffer per tile (S_0 | S_1 | O_0 | O_1
//                                = 512 TMEM columns), M128 N128 MMAs, fixed-reference softmax (STALE = true: no row max
//                                in the steady state, overflow-guarded exact redo).  1.28-1.42 PFLOP/s alone,
//                                1.07-1.19 PFLOP/s inside the power-capped 14B step.  STALE = false: classic online softmax.
//   attention_fwd_kernel         64-key steps with double-buffered scores (MV_ATTN_KSTEP=64): the first design; its N64
//                                score MMAs are shared-memory-bandwidth bound (192 B/clk), 1.19-1.24 PFLOP/s alone.
// Two more were written, validated (all GPU tests green) and measured this round, then removed because they did not
// pay (source: git commit cf8d306; numbers: profiles/README.md items 7-9): two threads per query row with 16 softmax
// warps, and two threads per row with the two warpgroups alternating between the tiles.
//
// What the measurements behind these variants showed (profiles/README.md, "attention" section): per Q tile the chain
// softmax(j) -> P.V(j) -> Q.K(j+1)^T -> softmax(j+1) is serial, the two tiles run in (self-organised, imperfect)
// antiphase, and the per-tile softmax time is the lever: 128 exponentials per thread cost 1233 clk for a lone warp
// (MUFU-bound would be 1024), the row max another ~350 clk — dropping the row max is worth +13 % alone and +6.6 % in
// the step, whereas packing two warps per sub-partition on a tile (the removed row-split kernels) or emulating part of the
// exponentials on the FMA pipe did not pay with 128-key steps.  Inside the 14B step the kernel runs under the 1000 W
// power cap (SM clock 1.5-1.7 GHz), so fewer instructions per score also means a higher clock.
//
// ---- 64-key kernel ----------------------------------------------------------------------------------------------
// One CTA = one head x 256 query rows = two 128-row Q tiles (w = 0, 1).  Keys are consumed in steps of 64.
//   warp 0       TMA producer: Q (once), then K_0, K_1, V_0, K_2, V_1, ... through an 8 x 16 KB ring
//   warps 1, 2   MMA issuers (one per Q tile):  S_w[b] = Q_w K_j^T  (SS, M128 N64 K16 x 8)
//                                               O_w   += P_w[b] V_j (A = P from TMEM, B = V MN-major smem, N128 K16 x 4)
//   warps 4-7    softmax warpgroup for Q tile 0: one thread per query row
//   warps 8-11   same for Q tile 1
// TMEM (512 columns): S_0[0] S_0[1] S_1[0] S_1[1] (4 x 64 fp32 columns) | O_0 | O_1 (2 x 128).  P (bf16, 32 columns)
// overwrites the S buffer it came from.  S is DOUBLE-BUFFERED per Q tile: Q K_{j+2}^T is issued right after
// P V_j, i.e. two steps ahead, so the softmax warpgroups never wait for the tensor pipe and the tensor pipe only
// ever waits for P — the score GEMM is off the softmax -> PV critical path (v1 of this kernel had a single S
// buffer per tile and paid softmax + PV + QK in series).
// Online softmax in fp32 with exp2; the O accumulator is rescaled lazily (only when a row max grew by more than
// 2^8, FA4-style) by the softmax warpgroup itself after waiting for the previous P.V; final 1/l scaling and bf16
// store by the same threads.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int kD = 128;             // head dim
constexpr int kBQ = 128;            // rows per Q tile
constexpr int kBKV = 64;            // keys per step
constexpr int kKVStages = 8;        // ring of 16 KB tiles
constexpr int kAttnThreads = 384;
constexpr int kDefaultEmu64 = 1;    // 64-key kernel: 1/4 of the exponentials on the FMA pipe (+5.7 % measured, 1172 -> 1239 TF/s)
constexpr int kDefaultEmu128 = 0;   // 128-key kernel: MUFU only (the emulation lengthens the per-tile critical path: -18 %)
constexpr int kDefaultStale = 1;     // 128-key kernel: fixed-reference softmax (no per-step row max): 1371 vs 1209 TF/s alone,
                                     // 1137 vs 1067 TF/s inside the power-capped 720P step; MV_ATTN_STALE=0: classic online softmax
constexpr int kDefaultKStep = 128;   // 128-key-step kernel below (in the 14B 720P step: 1019 vs 946 TF/s); MV_ATTN_KSTEP=64: the kernel above
constexpr int kDefaultSkewNs = 0;
constexpr int kDefaultWaitSpin = 0;
constexpr int kDefaultPrmtPack = 0;   // 1: truncating PRMT pack of P (PK) — MV_ATTN_PACK / mv_attention_config
constexpr float kPkScale = 1.0f + 3.0f / 1024.0f;
constexpr float kPkLog2 = 0.0042205915f;   // log2(1 + 3/1024)   // 0: waiting warps are parked in hardware (mbar_wait); 1: they poll
constexpr int kDefaultPingPong = 0;   // skewing the two warpgroups' start had no measurable effect
constexpr uint32_t kQTileBytes = kBQ * kD * 2;       // 32 KB
constexpr uint32_t kQHalfBytes = kQTileBytes / 2;    // [128 x 64] 128B-swizzled sub-tile
constexpr uint32_t kKVTileBytes = kBKV * kD * 2;     // 16 KB
constexpr uint32_t kKVHalfBytes = kKVTileBytes / 2;  // [64 x 64] sub-tile
constexpr uint32_t kAttnSmem = 2 * kQTileBytes + kKVStages * kKVTileBytes + 1024 + 512;

struct AttnParams {
  __nv_bfloat16* o;
  int64_t ldo;
  int Lq, Lk, n_kv;
  float scale_log2;
  // Ulysses return path fused into the epilogue: query row r belongs to rank r / rows_per_rank and is stored
  // straight into that rank's receive buffer o_dst[rank][src_rank][r % rows_per_rank][H*128] (peer pointers over
  // NVLink).  n_dst == 0: plain local output.
  __nv_bfloat16* o_dst[8];
  int n_dst, src_rank, rows_per_rank;
  int pingpong;  // 1: the two softmax warpgroups take turns on the exp phase (named-barrier token), so that one
                 // warpgroup's MUFU-bound exponentials overlap the other's barrier / TMEM / max bookkeeping instead of
                 // both queueing on the 16-lane MUFU pipe at once (FA3/FA4 "ping-pong")
  int skew_ns;  // initial delay of the second softmax warpgroup (MV_ATTN_SKEW): puts the two warpgroups' exp phases in
                // antiphase so that they do not queue on the MUFU pipe at the same time
  unsigned long long* trace;  // diagnostics (mv_attention_fwd_trace): clock64 stamps of CTA (0, 0), see the entry point
  int trace_steps;
  int wait_spin;  // 1: the waits on the per-tile chain (scores ready, P ready, P.V done, K/V landed) poll instead of
                  // parking the warp (MV_ATTN_WAIT_SPIN / mv_attention_config): A/B of the wake-up latency
  int order;  // 0 (default): Q_w K_{j+2}^T is issued after P_w V_j has drained (explicit o_done wait);
              // 1 (MV_ATTN_ORDER=1): issued right behind it, relying on in-order execution of the tensor pipe
};

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// 2^x on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial, max rel. error 1.6e-4 —
// far below the bf16 rounding of P).  Used for a fraction of the exponentials so that the MUFU pipe
// (16 ex2/clk/SM, exactly as many cycles per tile as the tensor pipe needs) is not the co-bottleneck (FA4).
__device__ __forceinline__ float exp2_emu(float x) {
  x = fmaxf(x, -125.f);
  const float xr = x + 12582912.f;            // 1.5 * 2^23: integer part lands in the low mantissa bits
  const float f = x - (xr - 12582912.f);      // fractional part in [-0.5, 0.5]
  float pz = fmaf(0.05360212177038193f, f, 0.24237291514873505f);
  pz = fmaf(pz, f, 0.6935023665428162f);
  pz = fmaf(pz, f, 0.9999481439590454f);
  return __int_as_float(__float_as_int(pz) + (__float_as_int(xr) << 23));
}

template <int EMU>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // 2 tiles
  uint8_t* sKV = smem + 2 * kQTileBytes;     // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQTileBytes + kKVStages * kKVTileBytes);
  uint64_t* q_full = bars;                   // 1
  uint64_t* kv_full = bars + 1;              // kKVStages
  uint64_t* kv_empty = kv_full + kKVStages;  // kKVStages
  uint64_t* s_full = kv_empty + kKVStages;   // [w][b] -> 4
  uint64_t* p_full = s_full + 4;             // [w][b] -> 4
  uint64_t* o_done = p_full + 4;             // [w]    -> 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * kBQ);
  const int n_kv = p.n_kv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);  // released by both tiles' issuing warps
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrive per softmax warp
    }
    mbar_init(&o_done[0], 1);
    mbar_init(&o_done[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    // Both single-thread roles run WARP-UNIFORMLY (all 32 lanes execute the loops and poll the barriers) and only
    // the TMA / tcgen05 instructions themselves are predicated on elect.sync: descriptors and addresses then live
    // in uniform registers and ptxas emits back-to-back UTMALDG / UTCHMMA with no per-instruction
    // "uniformisation" loop (with `if (lane == 0)` around the whole role each MMA cost ~100 issue cycles).
    if (warp == 0) {
      // ------------------------------ TMA producer ------------------------------
      if (elect_one()) {
        mbar_expect_tx(q_full, 2 * kQTileBytes);
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + w * kQTileBytes + h * kQHalfBytes, &tmQ, q_full, h * 64, head, q0 + w * kBQ);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      auto load_tile = [&](const CUtensorMap* tm, int j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[stage], kKVTileBytes);
          tma_load_3d(sKV + stage * kKVTileBytes, tm, &kv_full[stage], 0, head, j * kBKV);
          tma_load_3d(sKV + stage * kKVTileBytes + kKVHalfBytes, tm, &kv_full[stage], 64, head, j * kBKV);
        }
        __syncwarp();
        if (++stage == kKVStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      load_tile(&tmK, 0);
      if (n_kv > 1) load_tile(&tmK, 1);
      for (int j = 0; j < n_kv; ++j) {
        load_tile(&tmV, j);
        if (j + 2 < n_kv) load_tile(&tmK, j + 2);
      }
    } else if (warp == 1 || warp == 2) {
      // ------------------------------ MMA issuers -------------------------------
      constexpr uint32_t idesc_qk = make_idesc_bf16(kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kBQ, kD, 0, 1);
      const uint32_t sQ_addr = smem_u32(sQ);
      const uint32_t sKV_addr = smem_u32(sKV);
      const uint64_t qdesc0 = make_desc_kmajor_sw128(sQ_addr);
      const uint64_t kdesc0 = make_desc_kmajor_sw128(sKV_addr);
      const uint64_t vdesc0 = make_desc_mnmajor_sw128(sKV_addr, kKVHalfBytes);
      // S_w[b] = Q_w K^T : 8 x (M128 N64 K16).  Descriptor start addresses advance in 16-byte units.
      auto issue_qk = [&](int w, int st, int b) {
        const uint32_t d_tmem = tmem_base + (w * 2 + b) * 64;
        const uint64_t qd = qdesc0 + ((w * kQTileBytes) >> 4);
        const uint64_t kd = kdesc0 + ((st * kKVTileBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint32_t qo = ((k >> 2) * kQHalfBytes + (k & 3) * 32) >> 4;
          const uint32_t ko = ((k >> 2) * kKVHalfBytes + (k & 3) * 32) >> 4;
          umma_ss(d_tmem, qd + qo, kd + ko, idesc_qk, k != 0 ? 1u : 0u);
        }
      };
      // O_w (+)= P_w[b] V : 4 x (M128 N128 K16), A = P from TMEM (32 columns of packed bf16 pairs)
      auto issue_pv = [&](int w, int st, int b, uint32_t acc) {
        const uint32_t d_tmem = tmem_base + 256 + w * 128;
        const uint32_t a_tmem = tmem_base + (w * 2 + b) * 64;
        const uint64_t vd = vdesc0 + ((st * kKVTileBytes) >> 4);
#pragma unroll
        for (int k = 0; k < kBKV / 16; ++k) umma_ts(d_tmem, a_tmem + k * 8, vd + ((k * 2048) >> 4), idesc_pv, (acc | k) != 0 ? 1u : 0u);
      };
      // One issuing warp PER Q TILE (warp 1 -> tile 0, warp 2 -> tile 1): each blocks only on its own tile's
      // barriers, so a late softmax warpgroup never delays the other tile's tensor work.  Both walk the K/V ring in
      // the same order; a ring slot is released when both have committed (kv_empty counts 2).
      const int w = warp - 1;
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == kKVStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      mbar_wait(q_full, 0);
      // prologue: scores of steps 0 and 1
      for (int j = 0; j < 2 && j < n_kv; ++j) {
        mbar_wait(&kv_full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          issue_qk(w, stage, j);
          umma_commit(&s_full[w * 2 + j]);
          umma_commit(&kv_empty[stage]);
        }
        __syncwarp();
        advance();
      }
      for (int j = 0; j < n_kv; ++j) {
        const int b = j & 1;
        const uint32_t par = (j >> 1) & 1;
        const int vstage = stage;
        const uint32_t vphase = phase;
        advance();
        const bool more = (j + 2 < n_kv);
        const int kstage = stage;
        const uint32_t kphase = phase;
        if (more) advance();
        mbar_wait(&kv_full[vstage], vphase);
        if (more && p.order == 1) mbar_wait(&kv_full[kstage], kphase);
        mbar_wait(&p_full[w * 2 + b], par);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(w, vstage, b, j > 0 ? 1u : 0u);
          umma_commit(&o_done[w]);
          umma_commit(&kv_empty[vstage]);
          if (more && p.order == 1) {
            // Q_w K_{j+2}^T right behind P_w V_j: the tensor pipe executes this thread's MMAs in issue order, so the
            // score MMA overwrites the S/P buffer only after P_w V_j has read P from it
            issue_qk(w, kstage, b);
            umma_commit(&s_full[w * 2 + b]);
            umma_commit(&kv_empty[kstage]);
          }
        }
        __syncwarp();
        if (more && p.order != 1) {
          // conservative variant (MV_ATTN_ORDER=0): wait until P_w V_j has drained before re-using its buffer
          mbar_wait(&kv_full[kstage], kphase);
          mbar_wait(&o_done[w], j & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_qk(w, kstage, b);
            umma_commit(&s_full[w * 2 + b]);
            umma_commit(&kv_empty[kstage]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax warpgroups ------------------------
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS0 = tmem_base + lane_base + wg * 128;        // S_w[b] = tS0 + b * 64
    const uint32_t tO = tmem_base + lane_base + 256 + wg * 128;
    const float sl2 = p.scale_log2;
    if (wg == 1 && p.skew_ns > 0) __nanosleep(p.skew_ns);
    // exp-phase token: named barrier 1 = "warpgroup 0 may run", 2 = "warpgroup 1 may run" (256 = 128 waiting + 128
    // arriving threads).  Warpgroup 1 primes barrier 1 so that warpgroup 0 goes first.
    if (p.pingpong && wg == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");
    float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
    float l_run = 0.f;
    // o_done[wg] completes one phase per P.V.  A parity wait only means something while waiter and barrier are within
    // one phase of each other: two phases ahead it passes spuriously, two phases behind it blocks for ever.  With
    // the score GEMM running two steps ahead this warpgroup is no longer paced by the P.V's (v1 was), so the phases
    // are consumed strictly in order and P.V(j-1) is always observed BEFORE P(j) is handed over (P.V(j) cannot
    // complete before that, so the barrier can never run ahead) — that P.V had a whole step to finish, the wait is
    // normally free.  A rescale needs the same wait, earlier.
    int o_seen = 0;
    auto wait_pv = [&](int upto) {  // returns when P.V(0..upto-1) of this tile have completed
      while (o_seen < upto) {
        mbar_wait(&o_done[wg], o_seen & 1);
        ++o_seen;
      }
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const uint32_t tS = tS0 + b * 64;
      mbar_wait(&s_full[wg * 2 + b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[2][32];
      tmem_ld_x32(tS, s[0]);
      tmem_ld_x32(tS + 32, s[1]);
      tc_wait_ld();
      const int valid = p.Lk - j * kBKV;
      if (valid < kBKV) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmax3(mx0, __uint_as_float(s[c][i + 0]), __uint_as_float(s[c][i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]));
        }
      const float m_new = fmax3(m_run, mx0, mx1);
      if (j == 0) {
        m_run = m_new;
      } else {
        const bool need = (m_new - m_run) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rescale O_w: the previous P.V of this tile must have landed first
          wait_pv(j);
          tc_fence_after();
          const float alpha = fast_exp2((m_run - m_new) * sl2);
          l_run *= alpha;
          m_run = m_new;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_x32(tO + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO + c * 32, o);
          }
        }
      }
      if (p.pingpong) {
        if (wg == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      const float neg_m = -m_run * sl2;
      const float2 sc2 = make_float2(sl2, sl2);
      const float2 nm2 = make_float2(neg_m, neg_m);
      float2 sum2 = make_float2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
          const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])), sc2, nm2);
          float2 e01, e23;
          e01.x = fast_exp2(x01.x);
          e01.y = (EMU >= 2) ? exp2_emu(x01.y) : fast_exp2(x01.y);
          e23.x = fast_exp2(x23.x);
          e23.y = (EMU >= 1) ? exp2_emu(x23.y) : fast_exp2(x23.y);
          sum2 = __fadd2_rn(sum2, __fadd2_rn(e01, e23));
          pk[c * 16 + (i >> 1)] = pack_bf16(e01.x, e01.y);
          pk[c * 16 + (i >> 1) + 1] = pack_bf16(e23.x, e23.y);
        }
      l_run += sum2.x + sum2.y;
      if (p.pingpong) {   // hand the token over (warpgroup 1 keeps it after its last step: arrivals == waits)
        if (wg == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
        else if (j + 1 < n_kv) asm volatile("bar.arrive 1, 256;" ::: "memory");
      }
      wait_pv(j);
      tmem_st_x32(tS, pk);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[wg * 2 + b]);
    }

    // ------------------------------ final epilogue ----------------------------
    wait_pv(n_kv);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int row = q0 + wg * kBQ + quad * 32 + lane;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row) * p.ldo + head * kD;
    if (p.n_dst > 0 && row < p.Lq) {
      const int dst = row / p.rows_per_rank;
      const int rl = row - dst * p.rows_per_rank;
      orow = p.o_dst[dst] + (static_cast<int64_t>(p.src_rank) * p.rows_per_rank + rl) * p.ldo + head * kD;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld_x32(tO + c * 32, o);
      tc_wait_ld();
      if (row < p.Lq) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          dst[i] = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ================================================================================================================
// 128-key-step variant (MV_ATTN_KSTEP=128): TMEM = S_0 | S_1 | O_0 | O_1 (4 x 128 fp32 columns), ONE score buffer per
// Q tile, K/V tiles of 128 keys (4 x 32 KB ring).  The score MMA is M128 N128 K16: per instruction the A slice
// (4 KB) + B slice (4 KB) are read from shared memory in the 64 cycles the MMA takes — at the 128 B/clk shared-memory
// limit, whereas the N64 score MMA of the 64-key kernel needs 6 KB per 32 cycles (192 B/clk: smem-bound at 2/3 of the
// tensor rate).  Per tile the chain softmax(j) -> P.V(j) -> Q.K(j+1)^T -> softmax(j+1) is serial (the score MMA is
// issued right behind the P.V and overwrites the P/S buffer in tensor-pipe issue order); the two Q tiles fill each
// other's gaps.  Half as many barrier hand-overs per key as the 64-key kernel.
// ================================================================================================================
constexpr int kBKV2 = 128;
constexpr int kKVStages2 = 4;
constexpr uint32_t kKVTileBytes2 = kBKV2 * kD * 2;     // 32 KB
constexpr uint32_t kKVHalfBytes2 = kKVTileBytes2 / 2;  // [128 x 64] sub-tile
constexpr uint32_t kAttnSmem2 = 2 * kQTileBytes + kKVStages2 * kKVTileBytes2 + 1024 + 512;

template <int EMU, bool PP, bool TRACE, bool STALE, bool HI = false, bool PK = false>
__global__ void __launch_bounds__(kAttnThreads, 1)
attention_fwd_k128_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // 2 tiles
  uint8_t* sKV = smem + 2 * kQTileBytes;     // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQTileBytes + kKVStages2 * kKVTileBytes2);
  uint64_t* q_full = bars;                    // 1
  uint64_t* kv_full = bars + 1;               // kKVStages2
  uint64_t* kv_empty = kv_full + kKVStages2;  // kKVStages2
  uint64_t* s_full = kv_empty + kKVStages2;   // [w] -> 2
  uint64_t* p_full = s_full + 2;              // [w] -> 2
  uint64_t* o_done = p_full + 2;              // [w] -> 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  // HI: the role warps (TMA producer, the two MMA issuers) are physical warps 8-11, the softmax warpgroups 0-7: the
  // sub-partition arbiter favours the highest warp id, so a pending tcgen05.mma / TMA issue never queues behind the
  // softmax instruction stream sharing its sub-partition (the shift is a multiple of 4: TMEM quadrants are unchanged)
  const int warp = HI ? ((threadIdx.x >> 5) + 4) % 12 : (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * kBQ);
  const int n_kv = (p.Lk + kBKV2 - 1) / kBKV2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kKVStages2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 2);  // released by both tiles' issuing warps
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one arrive per softmax warp
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool spin = p.wait_spin != 0;
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    if (spin) mbar_wait_spin(bar, parity);
    else mbar_wait(bar, parity);
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ------------------------------ TMA producer: Q, K_0, then V_j, K_{j+1} ------------------------------
      if (elect_one()) {
        mbar_expect_tx(q_full, 2 * kQTileBytes);
#pragma unroll
        for (int w = 0; w < 2; ++w)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_3d(sQ + w * kQTileBytes + h * kQHalfBytes, &tmQ, q_full, h * 64, head, q0 + w * kBQ);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      auto load_tile = [&](const CUtensorMap* tm, int j) {
        wait(&kv_empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[stage], kKVTileBytes2);
          tma_load_3d(sKV + stage * kKVTileBytes2, tm, &kv_full[stage], 0, head, j * kBKV2);
          tma_load_3d(sKV + stage * kKVTileBytes2 + kKVHalfBytes2, tm, &kv_full[stage], 64, head, j * kBKV2);
        }
        __syncwarp();
        if (++stage == kKVStages2) {
          stage = 0;
          phase ^= 1;
        }
      };
      load_tile(&tmK, 0);
      for (int j = 0; j < n_kv; ++j) {
        load_tile(&tmV, j);
        if (j + 1 < n_kv) load_tile(&tmK, j + 1);
      }
    } else if (warp == 1 || warp == 2) {
      // ------------------------------ MMA issuers: one warp per Q tile -------------------------------
      constexpr uint32_t idesc_qk = make_idesc_bf16(kBQ, kBKV2, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kBQ, kD, 0, 1);
      const int w = warp - 1;
      const uint64_t qdesc0 = make_desc_kmajor_sw128(smem_u32(sQ)) + ((w * kQTileBytes) >> 4);
      const uint64_t kdesc0 = make_desc_kmajor_sw128(smem_u32(sKV));
      const uint64_t vdesc0 = make_desc_mnmajor_sw128(smem_u32(sKV), kKVHalfBytes2);
      const uint32_t tS = tmem_base + w * 128;
      const uint32_t tO = tmem_base + 256 + w * 128;
      // S_w = Q_w K^T : 8 x (M128 N128 K16)
      auto issue_qk = [&](int st) {
        const uint64_t kd = kdesc0 + ((st * kKVTileBytes2) >> 4);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint32_t qo = ((k >> 2) * kQHalfBytes + (k & 3) * 32) >> 4;
          const uint32_t ko = ((k >> 2) * kKVHalfBytes2 + (k & 3) * 32) >> 4;
          umma_ss(tS, qdesc0 + qo, kd + ko, idesc_qk, k != 0 ? 1u : 0u);
        }
      };
      // O_w (+)= P_w V : 8 x (M128 N128 K16), A = P from TMEM (64 columns of packed bf16 pairs).
      // (Handing P over in two 64-key halves so that the first half of P.V overlaps the second half's exponentials
      // was measured: the mid-step tcgen05.wait::st + barrier arrive cost 14 % (1358 -> 1172 TF/s) — not done.)
      auto issue_pv = [&](int st, uint32_t acc) {
        const uint64_t vd = vdesc0 + ((st * kKVTileBytes2) >> 4);
#pragma unroll
        for (int k = 0; k < kBKV2 / 16; ++k) umma_ts(tO, tS + k * 8, vd + ((k * 2048) >> 4), idesc_pv, (acc | k) != 0 ? 1u : 0u);
      };
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == kKVStages2) {
          stage = 0;
          phase ^= 1;
        }
      };
      wait(q_full, 0);
      wait(&kv_full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(stage);
        umma_commit(&s_full[w]);
        umma_commit(&kv_empty[stage]);
      }
      __syncwarp();
      advance();
      for (int j = 0; j < n_kv; ++j) {
        const int vstage = stage;
        const uint32_t vphase = phase;
        advance();
        const bool more = (j + 1 < n_kv);
        const int kstage = stage;
        const uint32_t kphase = phase;
        if (more) advance();
        wait(&kv_full[vstage], vphase);
        if (more) wait(&kv_full[kstage], kphase);
        wait(&p_full[w], j & 1);
        tc_fence_after();
        if constexpr (TRACE) {
          if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && j < p.trace_steps)
            p.trace[(w * p.trace_steps + j) * 8 + 5] = clock64();
        }
        if (elect_one()) {
          issue_pv(vstage, j > 0 ? 1u : 0u);
          umma_commit(&o_done[w]);
          umma_commit(&kv_empty[vstage]);
          if (more) {
            // right behind P.V(j): this thread's MMAs execute in issue order, so the scores of step j+1 overwrite
            // the S/P buffer only after P.V(j) has read P from it
            issue_qk(kstage);
            umma_commit(&s_full[w]);
            umma_commit(&kv_empty[kstage]);
          }
        }
        __syncwarp();
        if constexpr (TRACE) {
          if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && j < p.trace_steps)
            p.trace[(w * p.trace_steps + j) * 8 + 6] = clock64();
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------ softmax warpgroups ------------------------
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + wg * 128;
    const uint32_t tO = tmem_base + lane_base + 256 + wg * 128;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
    float l_run = 0.f;
    // optional exp-phase token (MV_ATTN_PINGPONG=1): named barrier 1 = "warpgroup 0 may exponentiate", 2 = "warpgroup 1
    // may"; warpgroup 1 primes barrier 1 so that warpgroup 0 goes first
    if (PP && wg == 1) asm volatile("bar.arrive 1, 256;" ::: "memory");

    if constexpr (STALE) {
      // ---- fixed-reference variant (MV_ATTN_STALE=1): NO per-step row max.  The exponentials use a reference that is
      // only moved when it has to be: P = 2^(s * scale_log2 - ref2) is exact for any ref2 as long as nothing
      // overflows (bf16 P and the fp32 sums keep their relative precision at every magnitude; terms that underflow
      // are below 2^-126 of the row's largest weight), so the lazy-rescale threshold of the classic loop (2^8) can
      // be pushed to "when the step's exponentials sum to more than 2^64".  That test needs only the row sum, which
      // exists anyway: the row-max instructions (1/6 of the softmax instruction count), the rescale vote and the
      // max -> exp dependency are gone from the steady state.  A step that trips the test (a score ~2^57 above the
      // reference; step 0 sets the reference to its exact row max) is redone from the scores still in TMEM with
      // its true max, O and l rescaled, before anything is stored.  References are kept in scaled units.
      float ref2 = 0.f;      // P = 2^(s * scale_log2 - ref2)
      auto rescale_by = [&](float up) {   // ref2 += up (up >= 0); O and l follow
        const float alpha = fast_exp2(-up);
        l_run *= alpha;
        ref2 += up;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          tmem_ld_x32(tO + c * 32, o);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x32(tO + c * 32, o);
        }
      };
      for (int j = 0; j < n_kv; ++j) {
        const bool tr = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && quad == 0 && lane == 0 && j < p.trace_steps;
        unsigned long long* trow = TRACE ? p.trace + (wg * p.trace_steps + (tr ? j : 0)) * 8 : nullptr;
        // P.V(j-1) completes before the scores of step j (same issuing thread, committed first): taking its phase
        // FIRST keeps the (free) probe off the path between "scores ready" and the first tcgen05.ld
        if (j > 0) wait(&o_done[wg], (j - 1) & 1);
        wait(&s_full[wg], j & 1);
        if (j == 0 && wg == 1 && p.skew_ns > 0) {   // MV_ATTN_SKEW (clocks): one-time phase offset between the two tiles
          const long long t0 = clock64();
          while (clock64() - t0 < p.skew_ns) {}
        }
        tc_fence_after();
        if constexpr (TRACE) { if (tr) trow[0] = clock64(); }
        uint32_t s[4][32];
        const int valid = p.Lk - j * kBKV2;
        auto load_scores = [&]() {
#pragma unroll
          for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, s[c]);
          tc_wait_ld();
          if (valid < kBKV2) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
          }
        };
        auto row_max = [&]() {
          float mx[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            mx[c] = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx[c] = fmax3(mx[c], __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
          }
          return fmax3(fmax3(mx[0], mx[1], mx[2]), mx[3], -INFINITY);
        };
        load_scores();
        if constexpr (TRACE) { if (tr) trow[1] = clock64(); }
        if (j == 0) ref2 = row_max() * sl2;   // the first step fixes the reference at its exact row max
        if constexpr (TRACE) { if (tr) trow[2] = clock64(); }
        uint32_t pk[2][32];
        float2 sum2;
        bool redo;
        int redone = 0;   // warp-uniform
#pragma unroll 1
        do {
          const float2 sc2 = make_float2(sl2, sl2);
          // PK: P is rounded to bf16 by TRUNCATING 2^(x + log2(1 + 3/1024)) — the pre-scale adds 0.375..0.75 bf16 ulp, i.e.
          // round-to-nearest up to a sub-ulp threshold shift with ~zero mean bias — so the pack is one PRMT on the ALU pipe
          // instead of one F2FP conversion per pair on the pipe the exponentials run on; the constant is folded into the
          // FFMA that exists anyway and divided out of the row sum at the end (kPkScale).
          const float2 nm2 = make_float2(-ref2 + (PK ? kPkLog2 : 0.f), -ref2 + (PK ? kPkLog2 : 0.f));
          sum2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              const int c = h * 2 + cc;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
                const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])), sc2, nm2);
                float2 e01, e23;
                e01.x = fast_exp2(x01.x);
                e01.y = (EMU >= 2) ? exp2_emu(fminf(x01.y, 126.f)) : fast_exp2(x01.y);
                e23.x = fast_exp2(x23.x);
                e23.y = (EMU >= 1) ? exp2_emu(fminf(x23.y, 126.f)) : fast_exp2(x23.y);
                sum2 = __fadd2_rn(sum2, __fadd2_rn(e01, e23));
                if constexpr (PK) {
                  pk[h][cc * 16 + (i >> 1)] = __byte_perm(__float_as_uint(e01.x), __float_as_uint(e01.y), 0x7632);
                  pk[h][cc * 16 + (i >> 1) + 1] = __byte_perm(__float_as_uint(e23.x), __float_as_uint(e23.y), 0x7632);
                } else {
                  pk[h][cc * 16 + (i >> 1)] = pack_bf16(e01.x, e01.y);
                  pk[h][cc * 16 + (i >> 1) + 1] = pack_bf16(e23.x, e23.y);
                }
              }
            }
          }
          // overflow guard on the sum (inf included; !(x <= t) also catches NaN)
          redo = __any_sync(0xffffffffu, !(sum2.x + sum2.y <= 1.8446744073709552e19f));   // 2^64
          if (redo) {
            if (redone++ > 0) break;   // still not finite with the exact row max as reference: the INPUT holds inf / NaN.
                                       // Let it propagate into this row's output (as the classic loop does); never spin.
            load_scores();                                     // the scores are still in TMEM: P has not been stored yet
            rescale_by(fmaxf(row_max() * sl2 - ref2, 0.f));    // exact reference; the second pass cannot overflow
          }
        } while (redo);
        l_run += sum2.x + sum2.y;
        if constexpr (TRACE) { if (tr) trow[3] = clock64(); }
        tmem_st_x32(tS, pk[0]);
        tmem_st_x32(tS + 32, pk[1]);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[wg]);
        if constexpr (TRACE) { if (tr) trow[4] = clock64(); }
      }
    } else {
    for (int j = 0; j < n_kv; ++j) {
      // S(j) complete implies P.V(j-1) complete (same issuing thread, in order, and its commit came first); the o_done
      // phase is taken every step so that every phase of the barrier is observed in order.
      const bool tr = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && quad == 0 && lane == 0 && j < p.trace_steps;
      unsigned long long* trow = TRACE ? p.trace + (wg * p.trace_steps + (tr ? j : 0)) * 8 : nullptr;
      if (j > 0) wait(&o_done[wg], (j - 1) & 1);   // completes before s_full(j): probed first, off the S -> ld path
      wait(&s_full[wg], j & 1);
      tc_fence_after();
      if constexpr (TRACE) { if (tr) trow[0] = clock64(); }
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, s[c]);
      tc_wait_ld();
      if constexpr (TRACE) { if (tr) trow[1] = clock64(); }
      const int valid = p.Lk - j * kBKV2;
      if (valid < kBKV2) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
      }
      float mx[4];   // one dependent chain per 32-column chunk (4 x 16 max3 deep instead of 2 x 32)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mx[c] = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx[c] = fmax3(mx[c], __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
      }
      const float m_new = fmax3(m_run, fmax3(mx[0], mx[1], mx[2]), mx[3]);
      if (j == 0) {
        m_run = m_new;
      } else {
        const bool need = (m_new - m_run) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = fast_exp2((m_run - m_new) * sl2);
          l_run *= alpha;
          m_run = m_new;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_x32(tO + c * 32, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x32(tO + c * 32, o);
          }
        }
      }
      if constexpr (PP) {
        if (wg == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 2, 256;" ::: "memory");
      }
      if constexpr (TRACE) { if (tr) trow[2] = clock64(); }
      const float neg_m = -m_run * sl2;
      const float2 sc2 = make_float2(sl2, sl2);
      const float2 nm2 = make_float2(neg_m, neg_m);
      float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t pk[32];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = h * 2 + cc;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float2 x01 = __ffma2_rn(make_float2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
            const float2 x23 = __ffma2_rn(make_float2(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])), sc2, nm2);
            float2 e01, e23;
            e01.x = fast_exp2(x01.x);
            e01.y = (EMU >= 2) ? exp2_emu(x01.y) : fast_exp2(x01.y);
            e23.x = fast_exp2(x23.x);
            e23.y = (EMU >= 1) ? exp2_emu(x23.y) : fast_exp2(x23.y);
            sum2 = __fadd2_rn(sum2, __fadd2_rn(e01, e23));
            pk[cc * 16 + (i >> 1)] = pack_bf16(e01.x, e01.y);
            pk[cc * 16 + (i >> 1) + 1] = pack_bf16(e23.x, e23.y);
          }
        }
        tmem_st_x32(tS + h * 32, pk);
      }
      l_run += sum2.x + sum2.y;
      if constexpr (TRACE) { if (tr) trow[3] = clock64(); }
      if constexpr (PP) {   // hand the token over (warpgroup 1 keeps it after its last step: arrivals == waits)
        if (wg == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
        else if (j + 1 < n_kv) asm volatile("bar.arrive 1, 256;" ::: "memory");
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[wg]);
      if constexpr (TRACE) { if (tr) trow[4] = clock64(); }
    }

    }

    // ------------------------------ final epilogue ----------------------------
    wait(&o_done[wg], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = (PK && STALE ? kPkScale : 1.0f) / l_run;   // the row sum carries the pre-scale of the truncating pack
    const int row = q0 + wg * kBQ + quad * 32 + lane;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row) * p.ldo + head * kD;
    if (p.n_dst > 0 && row < p.Lq) {
      const int dst = row / p.rows_per_rank;
      const int rl = row - dst * p.rows_per_rank;
      orow = p.o_dst[dst] + (static_cast<int64_t>(p.src_rank) * p.rows_per_rank + rl) * p.ldo + head * kD;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld_x32(tO + c * 32, o);
      tc_wait_ld();
      if (row < p.Lq) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
          dst[i] = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


}  // namespace mv

// Kernel-variant knobs: environment defaults (read once), overridable at run time through mv_attention_config
// (A/B measurements inside one process; the product path never calls it).
namespace {
struct AttnKnobs {
  int kstep, emu, stale, pingpong, order, skew, wait_spin, pack;
  bool init;
};
AttnKnobs g_knobs = {0, 0, 0, 0, 0, 0, 0, 0, false};
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}
AttnKnobs& attn_knobs() {
  if (!g_knobs.init) {
    const int ks = env_int("MV_ATTN_KSTEP", mv::kDefaultKStep);
    g_knobs.kstep = (ks == 128 || ks == 64) ? ks : mv::kDefaultKStep;
    const int emu_dflt = g_knobs.kstep == 128 ? mv::kDefaultEmu128 : mv::kDefaultEmu64;
    const int em = env_int("MV_ATTN_EMU", emu_dflt);
    g_knobs.emu = (em >= 0 && em <= 2) ? em : emu_dflt;
    g_knobs.stale = env_int("MV_ATTN_STALE", mv::kDefaultStale);
    g_knobs.pingpong = env_int("MV_ATTN_PINGPONG", mv::kDefaultPingPong);
    g_knobs.order = env_int("MV_ATTN_ORDER", 0) == 1 ? 1 : 0;   // default 0: same speed since each tile has its own issuer
    g_knobs.skew = env_int("MV_ATTN_SKEW", mv::kDefaultSkewNs);
    g_knobs.wait_spin = env_int("MV_ATTN_WAIT_SPIN", mv::kDefaultWaitSpin) != 0 ? 1 : 0;
    g_knobs.pack = env_int("MV_ATTN_PACK", mv::kDefaultPrmtPack) != 0 ? 1 : 0;
    g_knobs.init = true;
  }
  return g_knobs;
}
}  // namespace

extern "C" int mv_attention_config(int kstep, int emu, int stale, int pingpong, int skew, int wait_spin, int pack) {
  AttnKnobs& kn = attn_knobs();
  if ((kstep >= 0 && kstep != 64 && kstep != 128) || emu > 2) {
    mv::set_error("mv_attention_config: kstep must be 64 or 128, emu 0..2 (negative = keep)");
    return MV_E_SHAPE;
  }
  if (kstep >= 0) kn.kstep = kstep;
  if (emu >= 0) kn.emu = emu;
  if (stale >= 0) kn.stale = stale;
  if (pingpong >= 0) kn.pingpong = pingpong;
  if (skew >= 0) kn.skew = skew;
  if (wait_spin >= 0) kn.wait_spin = wait_spin != 0 ? 1 : 0;
  if (pack >= 0) kn.pack = pack != 0 ? 1 : 0;
  return MV_OK;
}

static int attention_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                          int64_t ldo, int Lq, int Lk, int H, float softmax_scale, void* const* o_dst, int n_dst,
                          int src_rank, int rows_per_rank, mv_stream_t stream,
                          unsigned long long* trace = nullptr, int trace_steps = 0) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(Lq > 0 && Lk > 0 && H > 0, "mv_attention_fwd: empty problem Lq=%d Lk=%d H=%d", Lq, Lk, H);
  MV_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
             "mv_attention_fwd: row strides must be multiples of 8 elements");
  MV_REQUIRE(ldq >= (int64_t)H * kD && ldk >= (int64_t)H * kD && ldv >= (int64_t)H * kD && ldo >= (int64_t)H * kD,
             "mv_attention_fwd: row stride smaller than H*128");
  MV_REQUIRE(n_dst > 0 || (o != nullptr && (reinterpret_cast<uintptr_t>(o) & 15) == 0),
             "mv_attention_fwd: o must be 16-byte aligned");
  MV_REQUIRE(n_dst >= 0 && n_dst <= 8, "mv_attention_fwd_scatter: at most 8 destinations");
  MV_REQUIRE(H <= 65535, "mv_attention_fwd: too many heads");

  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int L, int box_rows) {
    uint64_t dims[3] = {static_cast<uint64_t>(kD), static_cast<uint64_t>(H), static_cast<uint64_t>(L)};
    uint64_t str[3] = {2, static_cast<uint64_t>(kD) * 2, static_cast<uint64_t>(ld) * 2};
    uint32_t box[3] = {64, 1, static_cast<uint32_t>(box_rows)};
    return make_tmap_bf16(tm, base, 3, dims, str, box, true);
  };
  const AttnKnobs& kn = attn_knobs();
  const int kstep = trace != nullptr ? 128 : kn.kstep;   // keys per softmax step: 64 (double-buffered scores) or 128
                                                         // (the trace entry point exists for the 128-key kernel only)
  if ((rc = mk(&tmQ, q, ldq, Lq, kBQ)) != MV_OK) return rc;
  if ((rc = mk(&tmK, k, ldk, Lk, kstep)) != MV_OK) return rc;
  if ((rc = mk(&tmV, v, ldv, Lk, kstep)) != MV_OK) return rc;

  AttnParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.Lq = Lq;
  p.Lk = Lk;
  p.n_kv = (Lk + kBKV - 1) / kBKV;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.order = kn.order;
  p.skew_ns = kn.skew;
  p.wait_spin = kn.wait_spin;
  p.pingpong = kn.pingpong;
  p.trace = trace;
  p.trace_steps = trace_steps;
  p.n_dst = n_dst;
  p.src_rank = src_rank;
  p.rows_per_rank = rows_per_rank > 0 ? rows_per_rank : 1;
  for (int i = 0; i < 8; ++i) p.o_dst[i] = nullptr;
  if (n_dst > 0) {
    MV_REQUIRE(rows_per_rank > 0 && static_cast<int64_t>(rows_per_rank) * n_dst >= Lq && src_rank >= 0 && src_rank < n_dst,
               "mv_attention_fwd_scatter: rows_per_rank*n_dst must cover Lq");
    for (int i = 0; i < n_dst; ++i) {
      MV_REQUIRE(o_dst[i] != nullptr && (reinterpret_cast<uintptr_t>(o_dst[i]) & 15) == 0,
                 "mv_attention_fwd_scatter: destination %d null or misaligned", i);
      p.o_dst[i] = reinterpret_cast<__nv_bfloat16*>(o_dst[i]);
    }
  }

  // fraction of exponentials evaluated on the FMA pipe: 0 = none, 1 = 1/4, 2 = 1/2 (MV_ATTN_EMU overrides)
  const int emu = kn.emu;
  static bool attr_done = false;
  if (!attr_done) {
    attr_done = true;
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<1, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<2, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, false, false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_k128_kernel<0, false, false, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem2)));
  }
  dim3 grid((Lq + 2 * kBQ - 1) / (2 * kBQ), H);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (kstep == 128) {
    const int stale = kn.stale;   // 1 (default): fixed-reference softmax, no per-step row max
    if (stale && !p.pingpong && emu <= 1 && softmax_scale > 0.f) {
      if (p.trace != nullptr) attention_fwd_k128_kernel<0, false, true, true><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
      else if (emu == 1) attention_fwd_k128_kernel<1, false, false, true><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
      else if (kn.pack) attention_fwd_k128_kernel<0, false, false, true, false, true><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
      else if (roles_hi()) attention_fwd_k128_kernel<0, false, false, true, true><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
      else attention_fwd_k128_kernel<0, false, false, true><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    } else
    if (p.trace != nullptr && p.pingpong) attention_fwd_k128_kernel<0, true, true, false><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    else if (p.trace != nullptr) attention_fwd_k128_kernel<0, false, true, false><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    else if (p.pingpong) attention_fwd_k128_kernel<0, true, false, false><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    else if (emu == 2) attention_fwd_k128_kernel<2, false, false, false><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    else if (emu == 1) attention_fwd_k128_kernel<1, false, false, false><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    else attention_fwd_k128_kernel<0, false, false, false><<<grid, kAttnThreads, kAttnSmem2, st>>>(tmQ, tmK, tmV, p);
    MV_CHECK_LAUNCH("attention_fwd_k128_kernel");
    return MV_OK;
  }
  if (emu == 2) attention_fwd_kernel<2><<<grid, kAttnThreads, kAttnSmem, st>>>(tmQ, tmK, tmV, p);
  else if (emu == 1) attention_fwd_kernel<1><<<grid, kAttnThreads, kAttnSmem, st>>>(tmQ, tmK, tmV, p);
  else attention_fwd_kernel<0><<<grid, kAttnThreads, kAttnSmem, st>>>(tmQ, tmK, tmV, p);
  MV_CHECK_LAUNCH("attention_fwd_kernel");
  return MV_OK;
}

extern "C" int mv_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                void* o, int64_t ldo, int Lq, int Lk, int H, float softmax_scale,
                                mv_stream_t stream) {
  return attention_impl(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, H, softmax_scale, nullptr, 0, 0, 0, stream);
}

extern "C" int mv_attention_fwd_scatter(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                        int64_t ldv, void* const* o_dst, int n_dst, int src_rank, int rows_per_rank,
                                        int64_t ldo, int Lq, int Lk, int H, float softmax_scale, mv_stream_t stream) {
  return attention_impl(q, ldq, k, ldk, v, ldv, nullptr, ldo, Lq, Lk, H, softmax_scale, o_dst, n_dst, src_rank,
                        rows_per_rank, stream);
}

// Diagnostics: mv_attention_fwd on the 128-key-step kernel with clock64 stamps of CTA (0, head 0) written to
// trace[tile w (2)][step (trace_steps)][8]: 0 = scores visible to the softmax warpgroup, 1 = scores in registers,
// 2 = row max done, 3 = exponentials done, 4 = P handed over, 5 = P seen by the MMA warp, 6 = P.V + next Q.K^T issued.
extern "C" int mv_attention_fwd_trace(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                      void* o, int64_t ldo, int Lq, int Lk, int H, float softmax_scale,
                                      unsigned long long* trace, int trace_steps, mv_stream_t stream) {
  if (trace == nullptr || trace_steps <= 0) {
    mv::set_error("mv_attention_fwd_trace: trace buffer required");
    return MV_E_SHAPE;
  }
  return attention_impl(q, ldq, k, ldk, v, ldv, o, ldo, Lq, Lk, H, softmax_scale, nullptr, 0, 0, 0, stream, trace,
                        trace_steps);
}
