// umT5 text-encoder kernels (reference wan/modules/t5.py).
//   t5_attention_kernel : softmax(q k^T + rel_pos_bias + key mask) v for one (head, 128-query tile) per CTA,
//                         head_dim 64, at most 512 keys — the whole score row block (128 x 512 fp32) lives in
//                         TMEM, so the softmax is exact two-pass (no online rescaling); P overwrites S in TMEM as
//                         packed bf16 and feeds the P.V MMA from TMEM (tcgen05.mma A-from-TMEM).
//   t5_rmsnorm_kernel   : T5LayerNorm (RMS, no mean subtraction, no bias)           t5.py:53-66
//   embed_gather_kernel : token embedding lookup                                    t5.py:304
//   mul_bf16_kernel     : fc1(x) * gelu(gate(x)) product of the gated FFN            t5.py:136-137
// The dense projections run on mv_gemm_bf16.
#include <algorithm>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int kT5D = 64;         // head dim (umT5-XXL: 4096 / 64 heads)
constexpr int kT5BQ = 128;       // queries per CTA = TMEM lanes
constexpr int kT5MaxKeys = 512;  // 512 fp32 TMEM columns
constexpr int kT5Threads = 128;
constexpr uint32_t kT5QBytes = kT5BQ * kT5D * 2;            // 16 KB
constexpr uint32_t kT5KVBytes = kT5MaxKeys * kT5D * 2;      // 64 KB
constexpr uint32_t kT5BoxRows = 256;                        // TMA box limit
constexpr uint32_t kT5BoxBytes = kT5BoxRows * kT5D * 2;     // 32 KB
constexpr uint32_t kT5LutFloats = 2 * kT5MaxKeys;           // relative positions -(Lq-1) .. Lk-1
constexpr size_t kT5Smem = 1024 + kT5QBytes + 2 * kT5KVBytes + kT5LutFloats * 4 + 64;

struct T5AttnParams {
  __nv_bfloat16* o;
  int64_t ldo;
  const float* bias;  // [H, bias_ld]: bias[h, (j - i) + bias_center]
  int64_t bias_ld;
  int bias_center;
  int Lq, Lk, kv_len;
};

__global__ void __launch_bounds__(kT5Threads, 1)
t5_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const T5AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kT5QBytes;
  uint8_t* sV = sK + kT5KVBytes;
  float* lut = reinterpret_cast<float*>(sV + kT5KVBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(lut + kT5LutFloats);
  uint64_t* tma_full = bars;
  uint64_t* mma_done = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * kT5BQ;
  // keys that take part: [0, kv_len); processed in 32-key chunks
  const int kv_pad = min((p.kv_len + 31) & ~31, kT5MaxKeys);
  const int n_box = (min(p.Lk, kv_pad) + kT5BoxRows - 1) / kT5BoxRows;   // 256-key TMA boxes / score MMAs

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(tma_full, 1);
    mbar_init(mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // relative-position bias of this head for every (key - query) distance this tile can see
  {
    const int lo = -(q0 + kT5BQ - 1);  // smallest j - i
    for (int t = threadIdx.x; t < kT5MaxKeys + kT5BQ; t += kT5Threads) {
      int idx = lo + t + p.bias_center;
      idx = max(0, min(idx, static_cast<int>(p.bias_ld) - 1));
      lut[t] = p.bias != nullptr ? p.bias[static_cast<int64_t>(head) * p.bias_ld + idx] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(tma_full, kT5QBytes + 2 * n_box * kT5BoxBytes);
      tma_load_3d(sQ, &tmQ, tma_full, 0, head, q0);
      for (int b = 0; b < n_box; ++b) {
        tma_load_3d(sK + b * kT5BoxBytes, &tmK, tma_full, 0, head, b * kT5BoxRows);
        tma_load_3d(sV + b * kT5BoxBytes, &tmV, tma_full, 0, head, b * kT5BoxRows);
      }
    }
    __syncwarp();
    mbar_wait(tma_full, 0);
    tc_fence_after();
    // S[128, 256 b .. 256 b + 255] = Q K_b^T : 4 x (M128 N256 K16) per box, fp32 in TMEM columns [256 b, 256 b + 256)
    constexpr uint32_t idesc_qk = make_idesc_bf16(kT5BQ, 256, 0, 0);
    const uint64_t qdesc = make_desc_kmajor_sw128(smem_u32(sQ));
    const uint64_t kdesc = make_desc_kmajor_sw128(smem_u32(sK));
    for (int b = 0; b < n_box; ++b) {
#pragma unroll
      for (int k = 0; k < kT5D / 16; ++k) {
        if (elect_one())
          umma_ss(tmem_base + b * 256, qdesc + ((k * 32) >> 4), kdesc + ((b * kT5BoxBytes + k * 32) >> 4), idesc_qk,
                  k != 0 ? 1u : 0u);
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(mma_done);
    __syncwarp();
  }
  mbar_wait(mma_done, 0);
  tc_fence_after();

  // ---- exact two-pass softmax, one thread per query row (TMEM lane) ----
  const uint32_t tS = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  const int row = q0 + warp * 32 + lane;
  // lut index of (key j, this row):  j - row - lo  with lo = -(q0 + 127)
  const float* my_lut = lut + (q0 + kT5BQ - 1 - row);
  const int n_chunk = kv_pad >> 5;
  constexpr float kLog2e = 1.4426950408889634f;
  float m = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < n_chunk; ++c) {
    uint32_t s[32];
    tmem_ld_x32(tS + c * 32, s);
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int j = c * 32 + i;
      const float v = __uint_as_float(s[i]) + my_lut[j];
      m = (j < p.kv_len) ? fmaxf(m, v) : m;
    }
  }
  float l = 0.f;
  const float neg_m = -m * kLog2e;
#pragma unroll 1
  for (int c = 0; c < n_chunk; ++c) {
    uint32_t s[32];
    tmem_ld_x32(tS + c * 32, s);
    tc_wait_ld();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const int j = c * 32 + i;
      float e0 = fast_exp2(fmaf(__uint_as_float(s[i]) + my_lut[j], kLog2e, neg_m));
      float e1 = fast_exp2(fmaf(__uint_as_float(s[i + 1]) + my_lut[j + 1], kLog2e, neg_m));
      e0 = (j < p.kv_len) ? e0 : 0.f;
      e1 = (j + 1 < p.kv_len) ? e1 : 0.f;
      l += e0 + e1;
      pk[i >> 1] = pack_bf16(e0, e1);
    }
    // P (bf16 pairs) overwrites S in place: columns [16c, 16c+16) were read in chunk c/2 <= c of this pass
    tmem_st_x16(tS + c * 16, pk);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();

  // ---- O = P V : (kv_pad / 16) x (M128 N64 K16), A = P from TMEM, accumulator at columns [256, 320) ----
  // (columns 256.. held scores of keys 256..511; every thread has consumed them before the barrier above)
  if (warp == 0) {
    tc_fence_after();
    constexpr uint32_t idesc_pv = make_idesc_bf16(kT5BQ, kT5D, 0, 1);
    const uint64_t vdesc = make_desc_mnmajor_sw128(smem_u32(sV), kT5KVBytes);
    const int n_k = kv_pad >> 4;
#pragma unroll 1
    for (int k = 0; k < n_k; ++k) {
      if (elect_one())
        umma_ts(tmem_base + 256, tmem_base + k * 8, vdesc + ((k * 2048) >> 4), idesc_pv, k != 0 ? 1u : 0u);
      __syncwarp();
    }
    if (elect_one()) umma_commit(mma_done);
    __syncwarp();
  }
  mbar_wait(mma_done, 1);
  tc_fence_after();

  {
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row) * p.ldo + head * kT5D;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_x32(tS + 256 + c * 32, o);
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
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// --------------------------------------------------------------------------------------------
// T5LayerNorm: y = bf16( bf16(x * rsqrt(mean(x^2) + eps)) * w )                   t5.py:61-66
// One CTA per row, row in registers (C <= 256 * 4 * kT5Vec).
// --------------------------------------------------------------------------------------------
constexpr int kT5RowThreads = 256;
constexpr int kT5Vec = 8;

__global__ void __launch_bounds__(kT5RowThreads)
t5_rmsnorm_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                  __nv_bfloat16* __restrict__ out, int64_t ldo, int C, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * ldx);
  const int nvec = C >> 2;
  float4 v[kT5Vec];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kT5Vec; ++i) {
    const int idx = threadIdx.x + i * kT5RowThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  float t = (threadIdx.x & 31) < (kT5RowThreads >> 5) ? red[threadIdx.x & 31] : 0.f;
  t = warp_sum(t);
  const float r = rsqrtf(t / static_cast<float>(C) + eps);
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<int64_t>(row) * ldo);
  const float4* wr = reinterpret_cast<const float4*>(w);
#pragma unroll
  for (int i = 0; i < kT5Vec; ++i) {
    const int idx = threadIdx.x + i * kT5RowThreads;
    if (idx < nvec) {
      const float4 g = wr[idx];
      uint2 o;
      o.x = pack_bf16(bf16_round(v[i].x * r) * g.x, bf16_round(v[i].y * r) * g.y);
      o.y = pack_bf16(bf16_round(v[i].z * r) * g.z, bf16_round(v[i].w * r) * g.w);
      orow[idx] = o;
    }
  }
}

// out_f32[n, :] = float(table_bf16[ids[n], :])                                     t5.py:304
__global__ void __launch_bounds__(256)
embed_gather_kernel(const __nv_bfloat16* __restrict__ table, int64_t ldt, const int64_t* __restrict__ ids,
                    int64_t vocab, float* __restrict__ out, int64_t ldo, int C) {
  const int n = blockIdx.x;
  int64_t id = ids[n];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const uint2* src = reinterpret_cast<const uint2*>(table + id * ldt);
  float4* dst = reinterpret_cast<float4*>(out + static_cast<int64_t>(n) * ldo);
  for (int i = threadIdx.x; i < (C >> 2); i += blockDim.x) {
    const uint2 u = src[i];
    dst[i] = make_float4(bf16_lo(u.x), bf16_hi(u.x), bf16_lo(u.y), bf16_hi(u.y));
  }
}

// out = bf16(a * b) over a [rows, C] bf16 matrix pair                               t5.py:137
__global__ void __launch_bounds__(256)
mul_bf16_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda, const __nv_bfloat16* __restrict__ b, int64_t ldb,
                __nv_bfloat16* __restrict__ out, int64_t ldo, int rows, int C) {
  const int cvec = C >> 3;
  const int64_t total = static_cast<int64_t>(rows) * cvec;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(t / cvec), c = static_cast<int>(t % cvec);
    const uint4 x = reinterpret_cast<const uint4*>(a + static_cast<int64_t>(r) * lda)[c];
    const uint4 y = reinterpret_cast<const uint4*>(b + static_cast<int64_t>(r) * ldb)[c];
    uint4 z;
    z.x = pack_bf16(bf16_lo(x.x) * bf16_lo(y.x), bf16_hi(x.x) * bf16_hi(y.x));
    z.y = pack_bf16(bf16_lo(x.y) * bf16_lo(y.y), bf16_hi(x.y) * bf16_hi(y.y));
    z.z = pack_bf16(bf16_lo(x.z) * bf16_lo(y.z), bf16_hi(x.z) * bf16_hi(y.z));
    z.w = pack_bf16(bf16_lo(x.w) * bf16_lo(y.w), bf16_hi(x.w) * bf16_hi(y.w));
    reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * ldo)[c] = z;
  }
}

}  // namespace mv

extern "C" int mv_t5_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                               void* o, int64_t ldo, const float* bias, int64_t bias_ld, int bias_center, int Lq,
                               int Lk, int kv_len, int H, mv_stream_t stream) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(Lq > 0 && Lk > 0 && H > 0 && H <= 65535, "mv_t5_attention: empty problem Lq=%d Lk=%d H=%d", Lq, Lk, H);
  MV_REQUIRE(Lk <= kT5MaxKeys, "mv_t5_attention: at most %d keys (got %d)", kT5MaxKeys, Lk);
  MV_REQUIRE(kv_len >= 1 && kv_len <= Lk, "mv_t5_attention: kv_len=%d outside [1, Lk=%d]", kv_len, Lk);
  MV_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
             "mv_t5_attention: row strides must be multiples of 8 elements");
  MV_REQUIRE(ldq >= (int64_t)H * kT5D && ldk >= (int64_t)H * kT5D && ldv >= (int64_t)H * kT5D &&
                 ldo >= (int64_t)H * kT5D,
             "mv_t5_attention: row stride smaller than H*64");
  MV_REQUIRE(o != nullptr && (reinterpret_cast<uintptr_t>(o) & 15) == 0, "mv_t5_attention: o must be 16-byte aligned");
  MV_REQUIRE(bias == nullptr || (bias_ld > 0 && bias_center >= 0 && bias_center < bias_ld),
             "mv_t5_attention: bias_center=%d outside the table (ld=%lld)", bias_center, (long long)bias_ld);

  CUtensorMap tmQ, tmK, tmV;
  auto mk = [&](CUtensorMap* tm, const void* base, int64_t ld, int L, int box_rows) {
    uint64_t dims[3] = {static_cast<uint64_t>(kT5D), static_cast<uint64_t>(H), static_cast<uint64_t>(L)};
    uint64_t str[3] = {2, static_cast<uint64_t>(kT5D) * 2, static_cast<uint64_t>(ld) * 2};
    uint32_t box[3] = {64, 1, static_cast<uint32_t>(box_rows)};
    return make_tmap_bf16(tm, base, 3, dims, str, box, true);
  };
  if ((rc = mk(&tmQ, q, ldq, Lq, kT5BQ)) != MV_OK) return rc;
  if ((rc = mk(&tmK, k, ldk, Lk, kT5BoxRows)) != MV_OK) return rc;
  if ((rc = mk(&tmV, v, ldv, Lk, kT5BoxRows)) != MV_OK) return rc;

  T5AttnParams p;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.bias = bias;
  p.bias_ld = bias_ld;
  p.bias_center = bias_center;
  p.Lq = Lq;
  p.Lk = Lk;
  p.kv_len = kv_len;
  static bool attr_set = false;
  if (!attr_set) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kT5Smem)));
    attr_set = true;
  }
  dim3 grid((Lq + kT5BQ - 1) / kT5BQ, H);
  t5_attention_kernel<<<grid, kT5Threads, kT5Smem, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  MV_CHECK_LAUNCH("t5_attention_kernel");
  return MV_OK;
}

extern "C" int mv_t5_rmsnorm(const float* x, int64_t ldx, const float* weight, void* out, int64_t ldo, int rows,
                             int C, float eps, mv_stream_t stream) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  if (rows == 0) return MV_OK;
  MV_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= kT5RowThreads * 4 * kT5Vec,
             "mv_t5_rmsnorm: C=%d must be a multiple of 4 and <= %d", C, kT5RowThreads * 4 * kT5Vec);
  MV_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && ldx >= C && ldo >= C, "mv_t5_rmsnorm: bad row strides");
  t5_rmsnorm_kernel<<<rows, kT5RowThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, weight, reinterpret_cast<__nv_bfloat16*>(out), ldo, C, eps);
  MV_CHECK_LAUNCH("t5_rmsnorm_kernel");
  return MV_OK;
}

extern "C" int mv_embed_gather(const void* table, int64_t ldt, int64_t vocab, const int64_t* ids, float* out,
                               int64_t ldo, int n, int C, mv_stream_t stream) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  if (n == 0) return MV_OK;
  MV_REQUIRE(n > 0 && C > 0 && C % 4 == 0 && vocab > 0, "mv_embed_gather: bad sizes n=%d C=%d", n, C);
  MV_REQUIRE(ldt % 4 == 0 && ldo % 4 == 0 && ldt >= C && ldo >= C, "mv_embed_gather: bad row strides");
  embed_gather_kernel<<<n, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(table), ldt, ids, vocab, out, ldo, C);
  MV_CHECK_LAUNCH("embed_gather_kernel");
  return MV_OK;
}

extern "C" int mv_mul_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int rows,
                           int C, mv_stream_t stream) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  if (rows == 0) return MV_OK;
  MV_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, "mv_mul_bf16: C=%d must be a multiple of 8", C);
  MV_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 && lda >= C && ldb >= C && ldo >= C,
             "mv_mul_bf16: bad row strides");
  const int64_t total = static_cast<int64_t>(rows) * (C >> 3);
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count()) * 8));
  mul_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), lda, reinterpret_cast<const __nv_bfloat16*>(b), ldb,
      reinterpret_cast<__nv_bfloat16*>(out), ldo, rows, C);
  MV_CHECK_LAUNCH("mul_bf16_kernel");
  return MV_OK;
}
