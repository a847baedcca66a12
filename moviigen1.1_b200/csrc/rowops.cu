// HBM-bound fused row kernels of the DiT block: LayerNorm + adaLN modulation, full-row RMSNorm + RoPE,
// patchify gather, fp32 head + unpatchify, fp32 time-embedding GEMV, sinusoidal embedding.
// Each replaces a chain of ATen elementwise launches in wan/modules/model.py (cited per kernel).
#include <math.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int kRowThreads = 256;

// Sum over the whole block; every thread gets the result.  `red` holds >= 32 floats.  Safe to call
// repeatedly with the same scratch.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

// --------------------------------------------------------------------------------------------
// LayerNorm (+affine) (+bf16 rounding) + (1+scale)*y+shift -> bf16      model.py:89-99,299,306,307
// One CTA per row; the row lives in registers (C <= 256*4*kLnVec).
// --------------------------------------------------------------------------------------------
constexpr int kLnVec = 8;  // float4 per thread -> C <= 8192

__global__ void __launch_bounds__(kRowThreads)
ln_modulate_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ shift,
                   const float* __restrict__ scale, const float* __restrict__ w, const float* __restrict__ b,
                   __nv_bfloat16* __restrict__ out, int64_t ldo, int C, float eps, int round_ln) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * ldx);
  const int nvec = C >> 2;
  float4 v[kLnVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kLnVec; ++i) {
    const int idx = threadIdx.x + i * kRowThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = block_sum(s, red) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kLnVec; ++i) {
    const int idx = threadIdx.x + i * kRowThreads;
    if (idx < nvec) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + bb * bb) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(block_sum(q, red) / static_cast<float>(C) + eps);
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<int64_t>(row) * ldo);
#pragma unroll
  for (int i = 0; i < kLnVec; ++i) {
    const int idx = threadIdx.x + i * kRowThreads;
    if (idx < nvec) {
      float y[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                    (v[i].w - mean) * rstd};
      if (w != nullptr) {
        const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + idx);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + idx);
        y[0] = y[0] * ww.x + bb.x; y[1] = y[1] * ww.y + bb.y; y[2] = y[2] * ww.z + bb.z; y[3] = y[3] * ww.w + bb.w;
      }
      if (round_ln) {
#pragma unroll
        for (int k = 0; k < 4; ++k) y[k] = bf16_round(y[k]);
      }
      if (scale != nullptr) {
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + idx);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + idx);
        y[0] = y[0] * (1.f + sc.x) + sh.x; y[1] = y[1] * (1.f + sc.y) + sh.y;
        y[2] = y[2] * (1.f + sc.z) + sh.z; y[3] = y[3] * (1.f + sc.w) + sh.w;
      }
      uint2 o;
      o.x = pack_bf16(y[0], y[1]);
      o.y = pack_bf16(y[2], y[3]);
      orow[idx] = o;
    }
  }
}

// --------------------------------------------------------------------------------------------
// WanRMSNorm over the full row + RoPE, in place on bf16.             model.py:70-86, 39-67
// One CTA per row; 8 bf16 (one uint4) per thread per step.
// --------------------------------------------------------------------------------------------
constexpr int kRmsVec = 4;  // uint4 per thread -> C <= 256*8*4 = 8192

// When `out` is given the result is written out of place in the Ulysses send layout
// out[dst][row][c'] with dst = col / (C / sp_world), c' = col % (C / sp_world) (rows = gridDim.x), i.e. the
// head-scatter of xdit_context_parallel.py:185-190 is fused into this pass.  weight == nullptr skips the
// norm (plain scatter copy, used for V).
struct ScatterTable {
  __nv_bfloat16* dst[8];  // per destination rank: base of its receive buffer [src rank][rows][C / sp_world]
  int n;                  // 0: no table (use `out`), else sp_world
  int src_rank;
};

// With a ScatterTable the head group of destination rank d is stored straight into rank d's HBM (a peer pointer
// mapped over NVLink): the all-to-all of xdit_context_parallel.py:185-190 happens inside this pass.
__global__ void __launch_bounds__(kRowThreads)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ weight,
                    const float* __restrict__ cs, int C, int head_dim, float eps,
                    __nv_bfloat16* __restrict__ out, int sp_world, const ScatterTable tab) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld);
  const int nvec = C >> 3;
  uint4 v[kRmsVec];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kRmsVec; ++i) {
    const int idx = threadIdx.x + i * kRowThreads;
    if (idx < nvec) {
      v[i] = xr[idx];
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = bf16_lo(u[k]), bb = bf16_hi(u[k]);
        ss += a * a + bb * bb;
      }
    }
  }
  const float rinv = weight != nullptr ? rsqrtf(block_sum(ss, red) / static_cast<float>(C) + eps) : 1.f;
  const int half = head_dim >> 1;
  const int cols_per_rank = C / sp_world;
  const int rows_total = gridDim.x;
  const float* csr = cs != nullptr ? cs + static_cast<int64_t>(row) * head_dim : nullptr;  // [half][2]
#pragma unroll
  for (int i = 0; i < kRmsVec; ++i) {
    const int idx = threadIdx.x + i * kRowThreads;
    if (idx < nvec) {
      const int col = idx << 3;
      uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      if (weight != nullptr) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(weight + col));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(weight + col) + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // (x.float() * rsqrt(..)).type_as(x) * weight     -> bf16 rounding before the fp32 weight
        float a = bf16_round(bf16_lo(u[k]) * rinv) * wv[2 * k];
        float bb = bf16_round(bf16_hi(u[k]) * rinv) * wv[2 * k + 1];
        if (csr != nullptr) {
          const int pair = ((col + 2 * k) % head_dim) >> 1;
          const float2 c2 = __ldg(reinterpret_cast<const float2*>(csr) + pair);
          const float ra = a * c2.x - bb * c2.y;
          const float rb = a * c2.y + bb * c2.x;
          a = ra;
          bb = rb;
        }
        u[k] = pack_bf16(a, bb);
      }
      }
      (void)half;
      if (out == nullptr && tab.n == 0) {
        xr[idx] = make_uint4(u[0], u[1], u[2], u[3]);
      } else {
        const int dst = col / cols_per_rank, cc = col - dst * cols_per_rank;
        __nv_bfloat16* base = tab.n > 0 ? tab.dst[dst] : out;
        const int slot = tab.n > 0 ? tab.src_rank : dst;
        uint4* o = reinterpret_cast<uint4*>(base + (static_cast<int64_t>(slot) * rows_total + row) * cols_per_rank + cc);
        *o = make_uint4(u[0], u[1], u[2], u[3]);
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// patchify gather: latent fp32 [C,F,H,W] -> A bf16 [L, C*ph*pw]              model.py:529-533
// --------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const float* __restrict__ lat, __nv_bfloat16* __restrict__ a, int C, int F, int H,
                                int W, int ph, int pw) {
  const int Hp = H / ph, Wp = W / pw;
  const int Kc = C * ph * pw;
  const int64_t total = static_cast<int64_t>(F) * Hp * Wp * Kc;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(t % Kc);
    const int64_t n = t / Kc;
    const int j = col % pw;
    const int i = (col / pw) % ph;
    const int c = col / (pw * ph);
    const int wq = static_cast<int>(n % Wp);
    const int hq = static_cast<int>((n / Wp) % Hp);
    const int f = static_cast<int>(n / (static_cast<int64_t>(Wp) * Hp));
    const float v = lat[((static_cast<int64_t>(c) * F + f) * H + (hq * ph + i)) * W + (wq * pw + j)];
    a[t] = __float2bfloat16_rn(v);
  }
}

// --------------------------------------------------------------------------------------------
// Head (fp32 LN + modulation + Linear(C -> ph*pw*Cout)) fused with unpatchify. model.py:333-343,581-609
// CTA = 256 threads, kHeadRows tokens.  Stage 1: per-row LN statistics (one warp per row at a time).
// Stage 2: K-chunked fp32 GEMM [rows x 64] through shared memory.
// --------------------------------------------------------------------------------------------
constexpr int kHeadRows = 32;
constexpr int kHeadKC = 64;
constexpr int kHeadNOut = 64;  // ph*pw*Cout must be <= 64

__global__ void __launch_bounds__(256)
head_unpatchify_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ shift,
                       const float* __restrict__ scale, const float* __restrict__ Wh, const float* __restrict__ bh,
                       float* __restrict__ out, int L, int F, int Hp, int Wp, int ph, int pw, int Cout, int C,
                       float eps) {
  __shared__ float s_mean[kHeadRows], s_rstd[kHeadRows];
  __shared__ float s_x[kHeadRows][kHeadKC + 1];
  __shared__ float s_w[kHeadNOut][kHeadKC + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kHeadRows;
  const int nout = ph * pw * Cout;

  for (int r = warp; r < kHeadRows; r += 8) {
    const int row = row0 + r;
    float mean = 0.f, rstd = 0.f;
    if (row < L) {
      const float* xr = x + static_cast<int64_t>(row) * ldx;
      float s = 0.f;
      for (int k = lane; k < C; k += 32) s += xr[k];
      mean = warp_sum(s) / static_cast<float>(C);
      float q = 0.f;
      for (int k = lane; k < C; k += 32) {
        const float d = xr[k] - mean;
        q += d * d;
      }
      rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + eps);
    }
    if (lane == 0) {
      s_mean[r] = mean;
      s_rstd[r] = rstd;
    }
  }
  __syncthreads();

  // thread -> (row pair, 4 outputs): 256 threads = 16 row-pairs x 16 output-quads
  const int tr = (threadIdx.x >> 4) * 2;   // rows tr, tr+1
  const int tn = (threadIdx.x & 15) * 4;   // outputs tn..tn+3
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  for (int k0 = 0; k0 < C; k0 += kHeadKC) {
    for (int t = threadIdx.x; t < kHeadRows * kHeadKC; t += 256) {
      const int r = t / kHeadKC, kk = t % kHeadKC;
      const int row = row0 + r, k = k0 + kk;
      float v = 0.f;
      if (row < L && k < C) {
        v = (x[static_cast<int64_t>(row) * ldx + k] - s_mean[r]) * s_rstd[r];
        v = v * (1.f + scale[k]) + shift[k];
      }
      s_x[r][kk] = v;
    }
    for (int t = threadIdx.x; t < kHeadNOut * kHeadKC; t += 256) {
      const int n = t / kHeadKC, kk = t % kHeadKC;
      const int k = k0 + kk;
      s_w[n][kk] = (n < nout && k < C) ? Wh[static_cast<int64_t>(n) * C + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kHeadKC; ++kk) {
      const float a0 = s_x[tr][kk], a1 = s_x[tr + 1][kk];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float wv = s_w[tn + n][kk];
        acc[0][n] = fmaf(a0, wv, acc[0][n]);
        acc[1][n] = fmaf(a1, wv, acc[1][n]);
      }
    }
    __syncthreads();
  }
  const int Ho = Hp * ph, Wo = Wp * pw;
  if (F == 0) {  // token-major output [L, nout] (sequence-parallel path: all_gather then mv_unpatchify)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = row0 + tr + rr;
      if (row >= L) continue;
#pragma unroll
      for (int n = 0; n < 4; ++n)
        if (tn + n < nout) out[static_cast<int64_t>(row) * nout + tn + n] = acc[rr][n] + bh[tn + n];
    }
    return;
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int row = row0 + tr + rr;
    if (row >= L || row >= F * Hp * Wp) continue;
    const int wq = row % Wp, hq = (row / Wp) % Hp, f = row / (Wp * Hp);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int o = tn + n;
      if (o >= nout) continue;
      const int c = o % Cout;
      const int j = (o / Cout) % pw;
      const int i = o / (Cout * pw);
      out[((static_cast<int64_t>(c) * F + f) * Ho + (hq * ph + i)) * Wo + (wq * pw + j)] = acc[rr][n] + bh[o];
    }
  }
}

// tokens [L, ph*pw*Cout] -> video [Cout, F, Hp*ph, Wp*pw]                       model.py:581-609
__global__ void unpatchify_kernel(const float* __restrict__ tok, float* __restrict__ out, int F, int Hp, int Wp,
                                  int ph, int pw, int Cout) {
  const int nout = ph * pw * Cout;
  const int64_t total = static_cast<int64_t>(F) * Hp * Wp * nout;
  const int Ho = Hp * ph, Wo = Wp * pw;
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // iterate in OUTPUT order so that stores coalesce: t = ((c*F + f)*Ho + y)*Wo + x
    const int x = static_cast<int>(t % Wo);
    const int y = static_cast<int>((t / Wo) % Ho);
    const int f = static_cast<int>((t / (static_cast<int64_t>(Wo) * Ho)) % F);
    const int c = static_cast<int>(t / (static_cast<int64_t>(Wo) * Ho * F));
    const int wq = x / pw, j = x % pw, hq = y / ph, i = y % ph;
    const int64_t row = (static_cast<int64_t>(f) * Hp + hq) * Wp + wq;
    out[t] = tok[row * nout + (i * pw + j) * Cout + c];
  }
}

// --------------------------------------------------------------------------------------------
// fp32 GEMV for the time embedding MLP (M = 1)                            model.py:455-457,541-545
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_f32_vec_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                      float* __restrict__ out, int N, int K, int act_in) {
  extern __shared__ float sx[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v = x[k];
    if (act_in == 1) v = v / (1.f + expf(-v));  // SiLU
    sx[k] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const float* wr = W + static_cast<int64_t>(n) * K;
  float acc = 0.f;
  if ((K & 3) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(wr);
    const float4* x4 = reinterpret_cast<const float4*>(sx);
    for (int k = lane; k < (K >> 2); k += 32) {
      const float4 a = __ldg(w4 + k);
      const float4 c = x4[k];
      acc += (a.x * c.x + a.y * c.y) + (a.z * c.z + a.w * c.w);
    }
  } else {
    for (int k = lane; k < K; k += 32) acc += wr[k] * sx[k];
  }
  acc = warp_sum(acc);
  if (lane == 0) out[n] = acc + (b != nullptr ? b[n] : 0.f);
}

// sinusoidal_embedding_1d in fp64                                                  model.py:15-25
__global__ void sinusoid_kernel(const void* t, int t_is_int64, float* out, int dim) {
  const int half = dim >> 1;
  const double pos = t_is_int64 ? static_cast<double>(*reinterpret_cast<const long long*>(t))
                                : static_cast<double>(*reinterpret_cast<const float*>(t));
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const double w = pow(10000.0, -static_cast<double>(i) / static_cast<double>(half));
    const double a = pos * w;
    out[i] = static_cast<float>(cos(a));
    out[half + i] = static_cast<float>(sin(a));
  }
}

}  // namespace mv

using namespace mv;

extern "C" int mv_ln_modulate(const float* x, int64_t ldx, const float* shift, const float* scale, const float* w,
                              const float* b, void* out_bf16, int64_t ldo, int M, int C, float eps, int round_ln,
                              mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(M > 0 && C > 0, "mv_ln_modulate: empty problem");
  MV_REQUIRE(C % 4 == 0 && C <= kRowThreads * 4 * kLnVec, "mv_ln_modulate: C=%d must be a multiple of 4 and <= %d", C,
             kRowThreads * 4 * kLnVec);
  MV_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "mv_ln_modulate: ldx/ldo must be multiples of 4");
  MV_REQUIRE((shift == nullptr) == (scale == nullptr) && (w == nullptr) == (b == nullptr),
             "mv_ln_modulate: shift/scale and w/b must be given in pairs");
  ln_modulate_kernel<<<M, kRowThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, shift, scale, w, b, reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, C, eps, round_ln);
  MV_CHECK_LAUNCH("ln_modulate_kernel");
  return MV_OK;
}

extern "C" int mv_rmsnorm_rope(void* x_bf16, int64_t ld, const float* weight, const float* cs, int M, int C,
                               int head_dim, float eps, mv_stream_t stream) {
  return mv_qkv_prepare(x_bf16, ld, weight, cs, nullptr, 1, M, C, head_dim, eps, stream);
}

static int qkv_prepare_impl(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* out_bf16,
                            void* const* dst_ptrs, int src_rank, int sp_world, int M, int C, int head_dim, float eps,
                            mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(M > 0 && C > 0, "mv_rmsnorm_rope: empty problem");
  MV_REQUIRE(sp_world >= 1 && sp_world <= 8 && C % sp_world == 0 && (C / sp_world) % head_dim == 0,
             "mv_qkv_prepare: C=%d is not divisible into %d head groups", C, sp_world);
  MV_REQUIRE(out_bf16 != nullptr || dst_ptrs != nullptr || sp_world == 1, "mv_qkv_prepare: scatter needs an output buffer");
  MV_REQUIRE((reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0, "mv_qkv_prepare: out must be 16B aligned");
  MV_REQUIRE(C % 8 == 0 && C <= kRowThreads * 8 * kRmsVec, "mv_rmsnorm_rope: C=%d must be a multiple of 8 and <= %d", C,
             kRowThreads * 8 * kRmsVec);
  MV_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, "mv_rmsnorm_rope: rows must be 16B aligned");
  MV_REQUIRE(head_dim > 0 && head_dim % 2 == 0 && C % head_dim == 0, "mv_rmsnorm_rope: bad head_dim %d", head_dim);
  ScatterTable tab;
  tab.n = 0;
  tab.src_rank = src_rank;
  for (int i = 0; i < 8; ++i) tab.dst[i] = nullptr;
  if (dst_ptrs != nullptr) {
    MV_REQUIRE(src_rank >= 0 && src_rank < sp_world, "mv_qkv_prepare_p2p: bad source rank");
    tab.n = sp_world;
    for (int i = 0; i < sp_world; ++i) {
      MV_REQUIRE(dst_ptrs[i] != nullptr && (reinterpret_cast<uintptr_t>(dst_ptrs[i]) & 15) == 0,
                 "mv_qkv_prepare_p2p: destination %d is null or misaligned", i);
      tab.dst[i] = reinterpret_cast<__nv_bfloat16*>(dst_ptrs[i]);
    }
  }
  rmsnorm_rope_kernel<<<M, kRowThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(x_bf16), ld, weight, cs, C, head_dim, eps,
      reinterpret_cast<__nv_bfloat16*>(out_bf16), sp_world, tab);
  MV_CHECK_LAUNCH("rmsnorm_rope_kernel");
  return MV_OK;
}

extern "C" int mv_qkv_prepare(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* out_bf16,
                              int sp_world, int M, int C, int head_dim, float eps, mv_stream_t stream) {
  return qkv_prepare_impl(x_bf16, ld, weight, cs, out_bf16, nullptr, 0, sp_world, M, C, head_dim, eps, stream);
}

extern "C" int mv_qkv_prepare_p2p(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* const* dst_ptrs,
                                  int src_rank, int sp_world, int M, int C, int head_dim, float eps,
                                  mv_stream_t stream) {
  return qkv_prepare_impl(x_bf16, ld, weight, cs, nullptr, dst_ptrs, src_rank, sp_world, M, C, head_dim, eps, stream);
}

extern "C" int mv_patchify(const float* latent, void* a_bf16, int C, int F, int H, int W, int ph, int pw,
                           mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(C > 0 && F > 0 && H > 0 && W > 0 && ph > 0 && pw > 0 && H % ph == 0 && W % pw == 0,
             "mv_patchify: bad shape C=%d F=%d H=%d W=%d patch=(%d,%d)", C, F, H, W, ph, pw);
  const int64_t total = static_cast<int64_t>(C) * F * H * W;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  patchify_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      latent, reinterpret_cast<__nv_bfloat16*>(a_bf16), C, F, H, W, ph, pw);
  MV_CHECK_LAUNCH("patchify_kernel");
  return MV_OK;
}

extern "C" int mv_head_unpatchify(const float* x, int64_t ldx, const float* shift, const float* scale,
                                  const float* Wh, const float* bh, float* out, int F, int Hp, int Wp, int ph, int pw,
                                  int Cout, int C, float eps, mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(F > 0 && Hp > 0 && Wp > 0 && C > 0, "mv_head_unpatchify: empty problem");
  MV_REQUIRE(ph * pw * Cout <= kHeadNOut, "mv_head_unpatchify: ph*pw*Cout=%d exceeds %d", ph * pw * Cout, kHeadNOut);
  const int L = F * Hp * Wp;
  const int blocks = (L + kHeadRows - 1) / kHeadRows;
  head_unpatchify_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, shift, scale, Wh, bh, out, L,
                                                                               F, Hp, Wp, ph, pw, Cout, C, eps);
  MV_CHECK_LAUNCH("head_unpatchify_kernel");
  return MV_OK;
}

extern "C" int mv_linear_f32_vec(const float* x, const float* W, const float* b, float* out, int N, int K, int act_in,
                                 mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(N > 0 && K > 0 && K * 4 <= 160 * 1024, "mv_linear_f32_vec: bad shape N=%d K=%d", N, K);
  MV_REQUIRE((reinterpret_cast<uintptr_t>(W) & 15) == 0, "mv_linear_f32_vec: W must be 16B aligned");
  const size_t smem = static_cast<size_t>(K) * sizeof(float);
  if (smem > 48 * 1024) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(linear_f32_vec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  }
  linear_f32_vec_kernel<<<(N + 7) / 8, 256, smem, static_cast<cudaStream_t>(stream)>>>(x, W, b, out, N, K, act_in);
  MV_CHECK_LAUNCH("linear_f32_vec_kernel");
  return MV_OK;
}

extern "C" int mv_sinusoid_embed(const void* t, int t_is_int64, float* out, int dim, mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(dim > 0 && dim % 2 == 0, "mv_sinusoid_embed: dim must be even");
  sinusoid_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, t_is_int64, out, dim);
  MV_CHECK_LAUNCH("sinusoid_kernel");
  return MV_OK;
}

extern "C" int mv_head_tokens(const float* x, int64_t ldx, const float* shift, const float* scale, const float* Wh,
                              const float* bh, float* out_tokens, int L, int nout, int C, float eps,
                              mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(L > 0 && C > 0 && nout > 0 && nout <= kHeadNOut, "mv_head_tokens: bad shape L=%d nout=%d", L, nout);
  const int blocks = (L + kHeadRows - 1) / kHeadRows;
  // F = 0 selects the token-major store; (ph, pw, Cout) = (1, 1, nout)
  head_unpatchify_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, shift, scale, Wh, bh,
                                                                               out_tokens, L, 0, 1, 1, 1, 1, nout, C, eps);
  MV_CHECK_LAUNCH("head_tokens_kernel");
  return MV_OK;
}

extern "C" int mv_unpatchify(const float* tokens, float* out, int F, int Hp, int Wp, int ph, int pw, int Cout,
                             mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(F > 0 && Hp > 0 && Wp > 0 && ph > 0 && pw > 0 && Cout > 0, "mv_unpatchify: bad shape");
  const int64_t total = static_cast<int64_t>(F) * Hp * Wp * ph * pw * Cout;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  unpatchify_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(tokens, out, F, Hp, Wp, ph, pw, Cout);
  MV_CHECK_LAUNCH("unpatchify_kernel");
  return MV_OK;
}
