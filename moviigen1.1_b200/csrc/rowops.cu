// HBM-bound fused row kernels of the DiT block: LayerNorm + adaLN modulation, full-row RMSNorm + RoPE,
// patchify gather, fp32 head + unpatchify, fp32 time-embedding GEMV, sinusoidal embedding.
// Each replaces a chain of ATen elementwise launches in wan/modules/model.py (cited per kernel).
#include <math.h>
#include <string.h>

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
// q / k / (v) of one fused QKV row in ONE launch: full-row WanRMSNorm (+ RoPE) per part, optional Ulysses head
// scatter of all parts (model.py:139-151; xdit_context_parallel.py:169-190).
// One WARP per (row, part): the row part (C bf16) lives in the warp's registers (NV uint4 per lane), the sum of
// squares is a shuffle reduction (no __syncthreads), every lane keeps NV independent 16-byte loads in flight.
// Replaces three launches of rmsnorm_rope_kernel (one CTA per row, 62 % lane use at C = 5120) per layer.
// --------------------------------------------------------------------------------------------
struct NormPart {
  const float* weight;        // RMSNorm gain [C] or null (plain copy: V)
  int rope;                   // apply the cos/sin table
  int col0;                   // column offset of this part inside the row (elements)
  __nv_bfloat16* dst[8];      // scatter: base of destination d's slab for THIS source rank, [rows][C / n_dst];
                              // dst[0] == null: in place
};
struct NormArgs {
  NormPart part[3];
  int nparts;
  int n_dst;                  // 0: in place; else number of head groups (= sequence-parallel ranks)
};

template <int NV>
__global__ void __launch_bounds__(256)
qkv_norm_rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ld, const float* __restrict__ cs, int M, int C,
                     int head_dim, float eps, const NormArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (item >= static_cast<int64_t>(M) * a.nparts) return;   // warp-uniform
  const int row = static_cast<int>(item / a.nparts);
  const int pi = static_cast<int>(item - static_cast<int64_t>(row) * a.nparts);
  const NormPart& pt = a.part[pi];
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ld + pt.col0);
  const int nvec = C >> 3;
  uint4 v[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (idx < nvec) v[i] = xr[idx];
  }
  const float* w = pt.weight;
  float rinv = 1.f;
  if (w != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float lo = bf16_lo(u[k]), hi = bf16_hi(u[k]);
        ss += lo * lo + hi * hi;
      }
    }
    rinv = rsqrtf(warp_sum(ss) / static_cast<float>(C) + eps);
  }
  const float* csr = (pt.rope && cs != nullptr) ? cs + static_cast<int64_t>(row) * head_dim : nullptr;  // [half][2]
  const int cols_per_dst = a.n_dst > 0 ? C / a.n_dst : C;
  // A lane's columns are lane*8 + 256*i: when head_dim divides 256 they sit at the SAME position inside every head, so
  // the lane needs one set of 4 (cos, sin) pairs per row, not one per vector (2 loads instead of 2 * NV)
  const bool cs_once = (256 % head_dim) == 0;
  float c4[8];
  if (csr != nullptr && cs_once) {
    const float4* cp = reinterpret_cast<const float4*>(csr + ((lane << 3) % head_dim));
    const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
    c4[0] = c0.x; c4[1] = c0.y; c4[2] = c0.z; c4[3] = c0.w; c4[4] = c1.x; c4[5] = c1.y; c4[6] = c1.z; c4[7] = c1.w;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = lane + 32 * i;
    if (idx >= nvec) continue;
    const int col = idx << 3;
    uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
    if (w != nullptr) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + col));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + col) + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      if (csr != nullptr && !cs_once) {
        const float4* cp = reinterpret_cast<const float4*>(csr + (col % head_dim));   // 4 (cos, sin) pairs
        const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
        c4[0] = c0.x; c4[1] = c0.y; c4[2] = c0.z; c4[3] = c0.w; c4[4] = c1.x; c4[5] = c1.y; c4[6] = c1.z; c4[7] = c1.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // (x.float() * rsqrt(..)).type_as(x) * weight     -> bf16 rounding before the fp32 weight (model.py:83-86)
        float p0 = bf16_round(bf16_lo(u[k]) * rinv) * wv[2 * k];
        float p1 = bf16_round(bf16_hi(u[k]) * rinv) * wv[2 * k + 1];
        if (csr != nullptr) {
          const float r0 = p0 * c4[2 * k] - p1 * c4[2 * k + 1];
          const float r1 = p0 * c4[2 * k + 1] + p1 * c4[2 * k];
          p0 = r0;
          p1 = r1;
        }
        u[k] = pack_bf16(p0, p1);
      }
    }
    if (a.n_dst == 0) {
      if (w != nullptr) xr[idx] = make_uint4(u[0], u[1], u[2], u[3]);
    } else {
      const int d = col / cols_per_dst, cc = col - d * cols_per_dst;
      uint4* o = reinterpret_cast<uint4*>(pt.dst[d] + static_cast<int64_t>(row) * cols_per_dst + cc);
      *o = make_uint4(u[0], u[1], u[2], u[3]);
    }
  }
}

// --------------------------------------------------------------------------------------------
// Classifier-free guidance + one UniPC multistep update, all latent-sized arithmetic of a sampling step in one
// pass (text2video.py:245-254; fm_solvers_unipc.py:319-332 convert_model_output, :487-627 corrector, :351-485
// predictor).  The scalar coefficients are computed on the host exactly as the reference
// does (its sigmas live on the CPU); the element-wise operations are issued in the reference's order with
// explicitly rounded fp32 operations (no FMA contraction), so the result equals the chain of ATen kernels.
// --------------------------------------------------------------------------------------------
struct UniPCArgs {
  float guide;                 // noise = uncond + guide * (cond - uncond)
  float sigma;                 // x0 = sample - sigma * noise
  int use_corrector, c_order;  // corrector (UniC) of order c_order on (last_sample, history, x0)
  float c_ratio, c_a, c_b, c_rho_last, c_rk[2], c_rho[2];
  int p_order;                 // predictor (UniP) of order p_order
  float p_ratio, p_a, p_b, p_rk[2], p_rho[2];
};

__global__ void __launch_bounds__(256)
unipc_cfg_step_kernel(const float* cond, const float* uncond, const float* sample, const float* last_sample,
                      const float* h0, const float* h1, const float* h2, float* x0_out, float* sample_out,
                      float* prev_out, int64_t n, const UniPCArgs a) {   // no __restrict__: outputs may alias inputs
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float c = cond[i], u = uncond[i];
    const float noise = __fadd_rn(u, __fmul_rn(a.guide, __fsub_rn(c, u)));
    float s = sample[i];
    const float x0 = __fsub_rn(s, __fmul_rn(a.sigma, noise));
    const float m0 = h0 != nullptr ? h0[i] : 0.f;      // model_outputs[-1] before this step
    const float m1 = h1 != nullptr ? h1[i] : 0.f;      // model_outputs[-2]
    const float m2 = h2 != nullptr ? h2[i] : 0.f;      // model_outputs[-3]
    if (a.use_corrector) {
      const float xt = __fsub_rn(__fmul_rn(a.c_ratio, last_sample[i]), __fmul_rn(a.c_a, m0));
      float res = __fmul_rn(__fsub_rn(x0, m0), a.c_rho_last);
      if (a.c_order >= 2) res = __fadd_rn(res, __fmul_rn(__fdiv_rn(__fsub_rn(m1, m0), a.c_rk[0]), a.c_rho[0]));
      if (a.c_order >= 3) res = __fadd_rn(res, __fmul_rn(__fdiv_rn(__fsub_rn(m2, m0), a.c_rk[1]), a.c_rho[1]));
      s = __fsub_rn(xt, __fmul_rn(a.c_b, res));
    }
    // predictor: history is now (x0, m0, m1)
    float xn = __fsub_rn(__fmul_rn(a.p_ratio, s), __fmul_rn(a.p_a, x0));
    if (a.p_order >= 2) {
      float res = __fmul_rn(__fdiv_rn(__fsub_rn(m0, x0), a.p_rk[0]), a.p_rho[0]);
      if (a.p_order >= 3) res = __fadd_rn(res, __fmul_rn(__fdiv_rn(__fsub_rn(m1, x0), a.p_rk[1]), a.p_rho[1]));
      xn = __fsub_rn(xn, __fmul_rn(a.p_b, res));
    }
    x0_out[i] = x0;
    sample_out[i] = s;
    prev_out[i] = xn;
  }
}

// out[l, :] = mods[l, :] + e0[:]  (per-layer adaLN table `modulation + e`, model.py:292-295 / :341), fp32
__global__ void __launch_bounds__(256)
modulation_table_kernel(const float* __restrict__ mods, const float* __restrict__ e0, float* __restrict__ out,
                        int layers, int len) {
  const int64_t total = static_cast<int64_t>(layers) * len;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = mods[i] + e0[i % len];
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

// Launches qkv_norm_rope_kernel for up to three row parts.
static int launch_norm_parts(void* x_bf16, int64_t ld, const float* cs, int M, int C, int head_dim, float eps,
                             const NormArgs& a, mv_stream_t stream, const char* who) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(M > 0 && C > 0, "%s: empty problem", who);
  MV_REQUIRE(C % 8 == 0 && C <= 8192, "%s: C=%d must be a multiple of 8 and <= 8192", who, C);
  MV_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, "%s: rows must be 16B aligned", who);
  MV_REQUIRE(head_dim >= 8 && head_dim % 8 == 0 && C % head_dim == 0, "%s: bad head_dim %d", who, head_dim);
  MV_REQUIRE((reinterpret_cast<uintptr_t>(cs) & 15) == 0, "%s: cos/sin table must be 16B aligned", who);
  MV_REQUIRE(a.n_dst >= 0 && a.n_dst <= 8, "%s: at most 8 destinations", who);
  if (a.n_dst > 0) {
    MV_REQUIRE(C % a.n_dst == 0 && (C / a.n_dst) % head_dim == 0, "%s: C=%d is not divisible into %d head groups", who, C,
               a.n_dst);
    for (int p = 0; p < a.nparts; ++p)
      for (int d = 0; d < a.n_dst; ++d)
        MV_REQUIRE(a.part[p].dst[d] != nullptr && (reinterpret_cast<uintptr_t>(a.part[p].dst[d]) & 15) == 0,
                   "%s: destination %d of part %d is null or misaligned", who, d, p);
  }
  for (int p = 0; p < a.nparts; ++p)
    MV_REQUIRE(a.part[p].col0 % 8 == 0 && (reinterpret_cast<uintptr_t>(a.part[p].weight) & 15) == 0,
               "%s: part %d misaligned", who, p);
  const int64_t warps = static_cast<int64_t>(M) * a.nparts;
  const int64_t blocks = (warps + 7) / 8;
  MV_REQUIRE(blocks < (1ll << 31), "%s: too many rows", who);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* x = reinterpret_cast<__nv_bfloat16*>(x_bf16);
  const int nv = (C / 8 + 31) / 32;
  const int g = static_cast<int>(blocks);
  if (nv <= 4) qkv_norm_rope_kernel<4><<<g, 256, 0, st>>>(x, ld, cs, M, C, head_dim, eps, a);
  else if (nv <= 8) qkv_norm_rope_kernel<8><<<g, 256, 0, st>>>(x, ld, cs, M, C, head_dim, eps, a);
  else if (nv <= 20) qkv_norm_rope_kernel<20><<<g, 256, 0, st>>>(x, ld, cs, M, C, head_dim, eps, a);
  else qkv_norm_rope_kernel<32><<<g, 256, 0, st>>>(x, ld, cs, M, C, head_dim, eps, a);
  MV_CHECK_LAUNCH("qkv_norm_rope_kernel");
  return MV_OK;
}

static NormArgs one_part(const float* weight, const float* cs) {
  NormArgs a;
  memset(&a, 0, sizeof(a));
  a.nparts = 1;
  a.part[0].weight = weight;
  a.part[0].rope = cs != nullptr ? 1 : 0;
  return a;
}

extern "C" int mv_rmsnorm_rope(void* x_bf16, int64_t ld, const float* weight, const float* cs, int M, int C,
                               int head_dim, float eps, mv_stream_t stream) {
  if (weight == nullptr) {
    set_error("mv_rmsnorm_rope: weight is required");
    return MV_E_SHAPE;
  }
  return launch_norm_parts(x_bf16, ld, cs, M, C, head_dim, eps, one_part(weight, cs), stream, "mv_rmsnorm_rope");
}

extern "C" int mv_qkv_prepare(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* out_bf16,
                              int sp_world, int M, int C, int head_dim, float eps, mv_stream_t stream) {
  NormArgs a = one_part(weight, cs);
  if (out_bf16 != nullptr) {
    if (sp_world < 1 || sp_world > 8 || C % sp_world != 0) {
      set_error("mv_qkv_prepare: bad sp_world %d for C=%d", sp_world, C);
      return MV_E_SHAPE;
    }
    a.n_dst = sp_world;   // out[dst][row][C / sp_world]
    for (int d = 0; d < sp_world; ++d)
      a.part[0].dst[d] = reinterpret_cast<__nv_bfloat16*>(out_bf16) + static_cast<int64_t>(d) * M * (C / sp_world);
  } else if (weight == nullptr) {
    set_error("mv_qkv_prepare: nothing to do (no weight and no output)");
    return MV_E_SHAPE;
  }
  return launch_norm_parts(x_bf16, ld, cs, M, C, head_dim, eps, a, stream, "mv_qkv_prepare");
}

extern "C" int mv_qkv_prepare_p2p(void* x_bf16, int64_t ld, const float* weight, const float* cs, void* const* dst_ptrs,
                                  int src_rank, int sp_world, int M, int C, int head_dim, float eps,
                                  mv_stream_t stream) {
  if (dst_ptrs == nullptr || sp_world < 1 || sp_world > 8 || src_rank < 0 || src_rank >= sp_world || C % sp_world != 0) {
    set_error("mv_qkv_prepare_p2p: bad destination table / rank %d of %d", src_rank, sp_world);
    return MV_E_SHAPE;
  }
  NormArgs a = one_part(weight, cs);
  a.n_dst = sp_world;     // dst_ptrs[d][src_rank][row][C / sp_world]
  for (int d = 0; d < sp_world; ++d)
    a.part[0].dst[d] = dst_ptrs[d] == nullptr ? nullptr
                       : reinterpret_cast<__nv_bfloat16*>(dst_ptrs[d]) + static_cast<int64_t>(src_rank) * M * (C / sp_world);
  return launch_norm_parts(x_bf16, ld, cs, M, C, head_dim, eps, a, stream, "mv_qkv_prepare_p2p");
}

extern "C" int mv_qkv_norm_rope(void* qkv_bf16, int64_t ld, const float* gain_q, const float* gain_k, const float* cs,
                                int M, int C, int head_dim, float eps, void* const* dst_q, void* const* dst_k,
                                void* const* dst_v, int n_dst, int src_slot, mv_stream_t stream) {
  if (gain_q == nullptr || gain_k == nullptr) {
    set_error("mv_qkv_norm_rope: both gains are required");
    return MV_E_SHAPE;
  }
  NormArgs a;
  memset(&a, 0, sizeof(a));
  a.nparts = n_dst > 0 ? 3 : 2;
  a.n_dst = n_dst;
  a.part[0].weight = gain_q;
  a.part[1].weight = gain_k;
  a.part[0].rope = a.part[1].rope = cs != nullptr ? 1 : 0;
  for (int p = 0; p < 3; ++p) a.part[p].col0 = p * C;
  if (n_dst > 0) {
    if (n_dst > 8 || dst_q == nullptr || dst_k == nullptr || dst_v == nullptr || src_slot < 0 || C % n_dst != 0) {
      set_error("mv_qkv_norm_rope: bad destination tables (n_dst=%d)", n_dst);
      return MV_E_SHAPE;
    }
    void* const* tabs[3] = {dst_q, dst_k, dst_v};
    for (int p = 0; p < 3; ++p)
      for (int d = 0; d < n_dst; ++d)
        a.part[p].dst[d] = tabs[p][d] == nullptr ? nullptr
                           : reinterpret_cast<__nv_bfloat16*>(tabs[p][d]) + static_cast<int64_t>(src_slot) * M * (C / n_dst);
  }
  return launch_norm_parts(qkv_bf16, ld, cs, M, C, head_dim, eps, a, stream, "mv_qkv_norm_rope");
}

extern "C" int mv_unipc_cfg_step(const float* cond, const float* uncond, const float* sample, const float* last_sample,
                                 const float* hist0, const float* hist1, const float* hist2, float* x0_out,
                                 float* sample_out, float* prev_out, int64_t n, const float* coef, int ncoef,
                                 mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(n > 0 && cond != nullptr && uncond != nullptr && sample != nullptr && x0_out != nullptr &&
             sample_out != nullptr && prev_out != nullptr, "mv_unipc_cfg_step: null tensor / empty problem");
  MV_REQUIRE(coef != nullptr && ncoef == MV_UNIPC_NCOEF, "mv_unipc_cfg_step: expected %d coefficients, got %d",
             MV_UNIPC_NCOEF, ncoef);
  UniPCArgs a;
  a.guide = coef[0];
  a.sigma = coef[1];
  a.use_corrector = coef[2] != 0.f ? 1 : 0;
  a.c_order = static_cast<int>(coef[3]);
  a.c_ratio = coef[4];
  a.c_a = coef[5];
  a.c_b = coef[6];
  a.c_rho_last = coef[7];
  a.c_rk[0] = coef[8];
  a.c_rk[1] = coef[9];
  a.c_rho[0] = coef[10];
  a.c_rho[1] = coef[11];
  a.p_order = static_cast<int>(coef[12]);
  a.p_ratio = coef[13];
  a.p_a = coef[14];
  a.p_b = coef[15];
  a.p_rk[0] = coef[16];
  a.p_rk[1] = coef[17];
  a.p_rho[0] = coef[18];
  a.p_rho[1] = coef[19];
  MV_REQUIRE(a.c_order >= 0 && a.c_order <= 3 && a.p_order >= 1 && a.p_order <= 3, "mv_unipc_cfg_step: orders out of range");
  MV_REQUIRE(!a.use_corrector || (last_sample != nullptr && hist0 != nullptr && (a.c_order < 2 || hist1 != nullptr) &&
                                  (a.c_order < 3 || hist2 != nullptr)), "mv_unipc_cfg_step: corrector history missing");
  MV_REQUIRE((a.p_order < 2 || hist0 != nullptr) && (a.p_order < 3 || hist1 != nullptr),
             "mv_unipc_cfg_step: predictor history missing");
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  unipc_cfg_step_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cond, uncond, sample, last_sample, hist0, hist1, hist2, x0_out, sample_out, prev_out, n, a);
  MV_CHECK_LAUNCH("unipc_cfg_step_kernel");
  return MV_OK;
}

extern "C" int mv_modulation_table(const float* mods, const float* e0, float* out, int layers, int len,
                                   mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(layers > 0 && len > 0 && mods != nullptr && e0 != nullptr && out != nullptr, "mv_modulation_table: bad arguments");
  const int64_t total = static_cast<int64_t>(layers) * len;
  int64_t blocks = (total + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  modulation_table_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(mods, e0, out, layers, len);
  MV_CHECK_LAUNCH("modulation_table_kernel");
  return MV_OK;
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
