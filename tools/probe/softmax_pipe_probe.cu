// Pipe-overlap probe for the softmax inner loop of the attention kernel (diagnostics, not product code).
// Each warp runs the per-step exponential pass of attention_fwd_k128_kernel on 128 register-resident scores, REPS times,
// and reports clock64 cycles per pass.  Variants isolate what a lone warp per SM sub-partition can overlap:
//   0 MUFU + 1 FADD per element          1 classic pass (FFMA2, MUFU, FADD2, F2FP)      2 = 1 + separate row-max pre-pass
//   3 = 1 + shadow max on x (4 chains)    4 = 1 with 1/4 of the exponentials emulated      5 = 1 + bf16x2 max on packed P
//   6 = 1 + shadow max, 8 rotating chains 7 = pass without the sum (FFMA2, MUFU, F2FP)     8 = pass without the pack
// launched with 1 and 2 warps per sub-partition (128 / 256 threads per CTA, one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_pipe_probe softmax_pipe_probe.cu && ./softmax_pipe_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float exp2_emu(float x) {
  x = fmaxf(x, -125.f);
  const float xr = x + 12582912.f;
  const float f = x - (xr - 12582912.f);
  float pz = fmaf(0.05360212177038193f, f, 0.24237291514873505f);
  pz = fmaf(pz, f, 0.6935023665428162f);
  pz = fmaf(pz, f, 0.9999481439590454f);
  return __int_as_float(__float_as_int(pz) + (__float_as_int(xr) << 23));
}
__device__ __forceinline__ uint32_t bf162_max(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

constexpr int REPS = 64;

template <int V, int T>
__global__ void __launch_bounds__(T, 1) probe(const float* __restrict__ in, float* __restrict__ out, long long* cycles) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) {
    s[i] = in[(threadIdx.x * 128 + i) & 4095];
    asm volatile("" : "+f"(s[i]));   // keep the scores in registers (they come from TMEM in the real kernel)
  }
  float l = 0.f, m_run = 0.f;
  uint32_t keep = 0;
  const float sl2 = in[4096];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < REPS; ++r) {
    if (V == 2) {
      float mx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        mx[c] = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx[c] = fmax3(mx[c], s[c * 32 + i], s[c * 32 + i + 1]);
      }
      m_run = fmax3(m_run, fmax3(mx[0], mx[1], mx[2]), mx[3]) * 0.5f;
    }
    const float neg_m = -m_run * sl2 - r;
    const float2 sc2 = make_float2(sl2, sl2);
    const float2 nm2 = make_float2(neg_m, neg_m);
    float2 sum2 = make_float2(0.f, 0.f);
    float mx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) mx[c] = 0.f;
    uint32_t pmax = 0;
    uint32_t pk[64];
    if (V == 0) {
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 128; ++i) sum += fast_exp2(s[i] + neg_m);
      l += sum;
    } else {
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        const float2 x01 = __ffma2_rn(make_float2(s[i], s[i + 1]), sc2, nm2);
        const float2 x23 = __ffma2_rn(make_float2(s[i + 2], s[i + 3]), sc2, nm2);
        float2 e01, e23;
        e01.x = fast_exp2(x01.x);
        e01.y = fast_exp2(x01.y);
        e23.x = fast_exp2(x23.x);
        e23.y = (V == 4) ? exp2_emu(x23.y) : fast_exp2(x23.y);
        if (V == 3) {
          mx[i >> 5] = fmax3(mx[i >> 5], x01.x, x01.y);
          mx[i >> 5] = fmax3(mx[i >> 5], x23.x, x23.y);
        }
        if (V == 6) {
          mx[(i >> 1) & 7] = fmax3(mx[(i >> 1) & 7], x01.x, x01.y);
          mx[((i >> 1) + 1) & 7] = fmax3(mx[((i >> 1) + 1) & 7], x23.x, x23.y);
        }
        if (V != 7) sum2 = __fadd2_rn(sum2, __fadd2_rn(e01, e23));
        if (V != 8) {
          pk[i >> 1] = pack_bf16(e01.x, e01.y);
          pk[(i >> 1) + 1] = pack_bf16(e23.x, e23.y);
        } else {
          pk[i >> 1] = __float_as_uint(e01.x) ^ __float_as_uint(e01.y);
          pk[(i >> 1) + 1] = __float_as_uint(e23.x) ^ __float_as_uint(e23.y);
        }
        if (V == 5) pmax = bf162_max(pmax, bf162_max(pk[i >> 1], pk[(i >> 1) + 1]));
      }
      l += sum2.x + sum2.y;
#pragma unroll
      for (int i = 0; i < 64; i += 2) keep ^= pk[i] ^ pk[i + 1];   // P would go to TMEM here (2 STTM); 32 LOP3 keep it alive
      keep ^= pmax;
      float g = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) g = fmaxf(g, mx[c]);
      if (V == 3 || V == 6) m_run += g * 1e-9f;
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = l + m_run + __uint_as_float(keep);
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = (t1 - t0) / REPS;
}

template <int V>
void run(const float* in, float* out, long long* cyc, const char* name) {
  for (int threads = 128; threads <= 256; threads *= 2) {
    long long h = 0;
    for (int it = 0; it < 2; ++it) {
      if (threads == 128) probe<V, 128><<<148, 128>>>(in, out, cyc);
      else probe<V, 256><<<148, 256>>>(in, out, cyc);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("variant %d (%s) warps/subpartition %d: %lld clk per 128-element pass per warp (%s)\n", V, name, threads / 128, h,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  float* in;
  float* out;
  long long* cyc;
  cudaMalloc(&in, 4097 * sizeof(float));
  cudaMalloc(&out, 148 * 512 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  float h[4097];
  for (int i = 0; i < 4096; ++i) h[i] = -((i * 2654435761u) % 1000) * 0.01f;
  h[4096] = 0.1275f;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>(in, out, cyc, "MUFU + FADD");
  run<1>(in, out, cyc, "classic pass");
  run<2>(in, out, cyc, "row max + pass");
  run<3>(in, out, cyc, "pass + shadow max 4 chains");
  run<6>(in, out, cyc, "pass + shadow max 8 chains");
  run<5>(in, out, cyc, "pass + bf16x2 max on P");
  run<4>(in, out, cyc, "pass, 1/4 emulated");
  run<7>(in, out, cyc, "pass without sum");
  run<8>(in, out, cyc, "pass without bf16 pack");
  return 0;
}
