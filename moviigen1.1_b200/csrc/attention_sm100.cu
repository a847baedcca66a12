// Flash-attention forward for the DiT (head_dim 128, bf16, non-causal) on tcgen05 / TMEM / TMA.
//
// One CTA = one head x 256 query rows (two 128-row Q tiles that ping-pong on the tensor pipe):
//   warp 0       TMA producer: Q (once), then K_0, V_0, K_1, V_1, ... through a 4 x 32 KB ring
//   warp 1       MMA issuer (one thread): S_w = Q_w K_j^T (SS), O_w += P_w V_j (A = P from TMEM, B = V
//                MN-major from smem); TMEM: S0|S1|O0|O1 = 4 x 128 fp32 columns; P_w (bf16) aliases S_w
//   warps 4-7    softmax warpgroup for Q tile 0: one thread per query row, S from TMEM, online softmax
//   warps 8-11   same for Q tile 1
// The O accumulator is rescaled lazily (only when a row max grew by more than 2^8, FA4-style), by the
// softmax warpgroup itself, after waiting for the previous P.V of its tile; the final 1/l scaling and the
// bf16 store are done by the same threads.
//
// Replaces flash_attn.flash_attn_varlen_func (FA2 mma.sync kernels) called from
// wan/modules/attention.py:113-127 for self-attention (model.py:146-151) and cross-attention (:176).
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int kD = 128;             // head dim
constexpr int kBQ = 128;            // rows per Q tile
constexpr int kBKV = 128;           // keys per KV tile
constexpr int kKVStages = 4;        // ring of 32 KB tiles (K and V alternate)
constexpr int kAttnThreads = 384;
constexpr uint32_t kTileBytes = kBQ * kD * 2;       // 32 KB
constexpr uint32_t kHalfBytes = kTileBytes / 2;     // one [128 x 64] 128B-swizzled sub-tile
constexpr uint32_t kAttnSmem = 2 * kTileBytes + kKVStages * kTileBytes + 1024 + 256;

struct AttnParams {
  __nv_bfloat16* o;
  int64_t ldo;
  int Lq, Lk, n_kv;
  float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // 2 tiles
  uint8_t* sKV = smem + 2 * kTileBytes;     // ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + kKVStages) * kTileBytes);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // kKVStages
  uint64_t* kv_empty = kv_full + kKVStages; // kKVStages
  uint64_t* s_full = kv_empty + kKVStages;  // 2
  uint64_t* p_full = s_full + 2;            // 2
  uint64_t* o_done = p_full + 2;            // 2
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
      mbar_init(&kv_empty[i], 1);
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

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0 && lane == 0) {
      // ------------------------------ TMA producer ------------------------------
      mbar_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_3d(sQ + w * kTileBytes + h * kHalfBytes, &tmQ, q_full, h * 64, head, q0 + w * kBQ);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < 2 * n_kv; ++i) {
        const int j = i >> 1;
        const CUtensorMap* tm = (i & 1) ? &tmV : &tmK;
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&kv_full[stage], kTileBytes);
        tma_load_3d(sKV + stage * kTileBytes, tm, &kv_full[stage], 0, head, j * kBKV);
        tma_load_3d(sKV + stage * kTileBytes + kHalfBytes, tm, &kv_full[stage], 64, head, j * kBKV);
        if (++stage == kKVStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ------------------------------ MMA issuer --------------------------------
      constexpr uint32_t idesc_qk = make_idesc_bf16(kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(kBQ, kD, 0, 1);
      const uint32_t sQ_addr = smem_u32(sQ);
      const uint32_t sKV_addr = smem_u32(sKV);
      auto issue_qk = [&](int w, int st) {
        const uint32_t d_tmem = tmem_base + w * 128;
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint32_t off = (k >> 2) * kHalfBytes + (k & 3) * 32;
          const uint64_t adesc = make_desc_kmajor_sw128(sQ_addr + w * kTileBytes + off);
          const uint64_t bdesc = make_desc_kmajor_sw128(sKV_addr + st * kTileBytes + off);
          umma_ss(d_tmem, adesc, bdesc, idesc_qk, k != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int w, int st, uint32_t acc) {
        const uint32_t d_tmem = tmem_base + 256 + w * 128;
        const uint32_t a_tmem = tmem_base + w * 128;  // P_w: 64 columns of packed bf16 pairs
#pragma unroll
        for (int k = 0; k < kBKV / 16; ++k) {
          // 16 keys = 2 eight-row groups of 1024 B; the two 64-wide d halves are kHalfBytes apart
          const uint64_t bdesc = make_desc_mnmajor_sw128(sKV_addr + st * kTileBytes + k * 2048, kHalfBytes);
          umma_ts(d_tmem, a_tmem + k * 8, bdesc, idesc_pv, (acc | k) != 0 ? 1u : 0u);
        }
      };
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == kKVStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[stage], phase);
      tc_fence_after();
      issue_qk(0, stage);
      umma_commit(&s_full[0]);
      issue_qk(1, stage);
      umma_commit(&s_full[1]);
      umma_commit(&kv_empty[stage]);
      advance();
      for (int j = 0; j < n_kv; ++j) {
        const int vstage = stage;
        const uint32_t vphase = phase;
        advance();
        const bool more = (j + 1 < n_kv);
        const int kstage = stage;
        const uint32_t kphase = phase;
        if (more) advance();
        const uint32_t par = j & 1;
        mbar_wait(&kv_full[vstage], vphase);
        mbar_wait(&p_full[0], par);
        tc_fence_after();
        issue_pv(0, vstage, j > 0 ? 1u : 0u);
        umma_commit(&o_done[0]);
        if (more) {
          mbar_wait(&kv_full[kstage], kphase);
          tc_fence_after();
          issue_qk(0, kstage);
          umma_commit(&s_full[0]);
        }
        mbar_wait(&p_full[1], par);
        tc_fence_after();
        issue_pv(1, vstage, j > 0 ? 1u : 0u);
        umma_commit(&o_done[1]);
        umma_commit(&kv_empty[vstage]);
        if (more) {
          issue_qk(1, kstage);
          umma_commit(&s_full[1]);
          umma_commit(&kv_empty[kstage]);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ------------------------------ softmax warpgroups ------------------------
    const int wg = (warp - 4) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_base + wg * 128;
    const uint32_t tO = tmem_base + lane_base + 256 + wg * 128;
    const float sl2 = p.scale_log2;
    float m_run = -INFINITY;  // running (possibly stale) row max of raw scores
    float l_run = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[wg], j & 1);
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_x32(tS + c * 32, s[c]);
      tc_wait_ld();
      const int valid = p.Lk - j * kBKV;
      if (valid < kBKV) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[c][i + 0]));
          mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
        }
      const float m_new = fmaxf(fmaxf(m_run, fmaxf(mx0, mx1)), fmaxf(mx2, mx3));
      if (j == 0) {
        m_run = m_new;
      } else {
        const bool need = (m_new - m_run) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rescale O_w: the previous P.V of this tile must have landed first
          mbar_wait(&o_done[wg], (j - 1) & 1);
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
      const float neg_m = -m_run * sl2;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pk[2][32];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[c][i]), sl2, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[c][i + 1]), sl2, neg_m));
          sum0 += p0;
          sum1 += p1;
          pk[c >> 1][(c & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
        }
      l_run += sum0 + sum1;
      tmem_st_x32(tS, pk[0]);
      tmem_st_x32(tS + 32, pk[1]);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[wg]);
    }

    // ------------------------------ final epilogue ----------------------------
    mbar_wait(&o_done[wg], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_run;
    const int row = q0 + wg * kBQ + quad * 32 + lane;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row) * p.ldo + head * kD;
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

extern "C" int mv_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                void* o, int64_t ldo, int Lq, int Lk, int H, float softmax_scale,
                                mv_stream_t stream) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(Lq > 0 && Lk > 0 && H > 0, "mv_attention_fwd: empty problem Lq=%d Lk=%d H=%d", Lq, Lk, H);
  MV_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
             "mv_attention_fwd: row strides must be multiples of 8 elements");
  MV_REQUIRE(ldq >= (int64_t)H * kD && ldk >= (int64_t)H * kD && ldv >= (int64_t)H * kD && ldo >= (int64_t)H * kD,
             "mv_attention_fwd: row stride smaller than H*128");
  MV_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0, "mv_attention_fwd: o must be 16-byte aligned");
  MV_REQUIRE(H <= 65535, "mv_attention_fwd: too many heads");

  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int L) {
    uint64_t dims[3] = {static_cast<uint64_t>(kD), static_cast<uint64_t>(H), static_cast<uint64_t>(L)};
    uint64_t str[3] = {2, static_cast<uint64_t>(kD) * 2, static_cast<uint64_t>(ld) * 2};
    uint32_t box[3] = {64, 1, static_cast<uint32_t>(kBQ)};
    return make_tmap_bf16(tm, base, 3, dims, str, box, true);
  };
  if ((rc = mk(&tmQ, q, ldq, Lq)) != MV_OK) return rc;
  if ((rc = mk(&tmK, k, ldk, Lk)) != MV_OK) return rc;
  if ((rc = mk(&tmV, v, ldv, Lk)) != MV_OK) return rc;

  AttnParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.Lq = Lq;
  p.Lk = Lk;
  p.n_kv = (Lk + kBKV - 1) / kBKV;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;

  static bool attr_set = false;
  if (!attr_set) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kAttnSmem)));
    attr_set = true;
  }
  dim3 grid((Lq + 2 * kBQ - 1) / (2 * kBQ), H);
  attention_fwd_kernel<<<grid, kAttnThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  MV_CHECK_LAUNCH("attention_fwd_kernel");
  return MV_OK;
}
