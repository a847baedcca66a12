// WanVAE decoder convolutions as an implicit GEMM on tcgen05 (channels-last FP16 activations and weights: 10 mantissa
// bits like the TF32 convolutions the reference runs, fp32 accumulation; max-abs 4e-3 vs the fp32 reference where bf16
// storage measured 3e-2 — tests/test_vae_gpu.py).
//
//   out[t,h,w,:] = bias + sum_taps  in[t+dt, h+dh, w+dw, :] . Wtap^T        (+ residual)
//
// M tile = 128 output voxels = an 8 x 16 (h x w) patch of one frame.  For every tap and every block of BK input
// channels the A operand is ONE 4-D TMA box {BK, 16, 8, 1} of the input tensor at the shifted coordinate — the
// hardware's out-of-bounds zero fill implements the spatial zero padding and the causal temporal padding (frames
// before the first are zero, vae.py:17-36), and the box lands in shared memory already in the 128-row K-major
// swizzled layout tcgen05.mma wants ("im2col staging" without an im2col buffer).  B = the tap's weight slab
// [Cout_tile, BK] of the pre-packed matrix W[Cout][tap][Cin].  fp32 accumulation in TMEM (2 stages), persistent
// CTAs, warp-specialised exactly like gemm_sm100.cu.
//
// Serves CausalConv3d 3x3x3 / 1x1x1 (vae.py:17-36), the Conv2d 3x3 after nearest-2x upsampling as four 2x2
// sub-pixel convolutions on the low-resolution input (vae.py:74-79: the nearest-exact x2 + 3x3 kernel collapses to
// a 2x2 kernel per output parity, 2.25x fewer FLOPs and no upsampled intermediate), the temporal time_conv
// (3,1,1) with its frame interleave (vae.py:84-85,128-137) and the 96->3 head conv with the final clamp.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int kConvBM = 128;
constexpr int kConvTW = 16;   // tile width  (w)
constexpr int kConvTH = 8;    // tile height (h)
constexpr int kConvMaxStages = 16;
constexpr int kConvThreads = 192;
constexpr int kConvMaxTaps = 27;
constexpr uint32_t kConvRingBytes = 200 * 1024;         // operand ring; the stage count adapts to the tile shape
constexpr uint32_t kConvSmem = kConvRingBytes + 1024 + 512;

struct ConvParams {
  const float* bias;            // [Cout] or null
  const __half* res;     // residual, same addressing as out (fp16 channels-last) or null
  void* out;
  // output addressing (elements): off = base + t*os_t + h*os_h + w*os_w (+ parity remap) + channel
  int64_t o_base, os_t, os_h, os_w;
  int nsplit;                   // > 0: output channel block n0 >= nsplit goes to (n0 - nsplit) with + nsplit_off
  int64_t nsplit_off;
  int T, H, W;                  // output grid (stride-1 "same" convolutions: input H, W equal; input T = t_off + T)
  int t_off;                    // output frame t reads input frames t + t_off + dt: the first t_off input frames are the
                                // cached tail of the previous temporal chunk (CausalConv3d feature cache, vae.py:28-36,205-217)
  int st, sh, sw;               // convolution strides (encoder downsampling, vae.py:92-104): output voxel (t,h,w) reads input
                                // (st*t + t_off + dt, sh*h + dh, sw*w + dw); the TMA map carries sh / sw as traversal strides
  int Cin, Cout;
  int BN;                       // N tile (multiple of 16, <= 256)
  int ntaps;
  int8_t dt[kConvMaxTaps], dh[kConvMaxTaps], dw[kConvMaxTaps];
  int num_n, tiles_h, tiles_w, num_tiles, kblocks_per_tap;
  int out_mode;                 // 0: fp16 channels-last; 1: fp32 channel-first video [Cout_real,T,H,W] clamped to [-1,1]
  int cout_real;                // out_mode 1: number of real output channels (3)
  int stages;                   // ring depth (<= kConvMaxStages)
  uint32_t a_stage_bytes, b_stage_bytes;  // 1024-aligned slot sizes (kps blocks each)
  const float* norm_gamma;      // fused RMS_norm + SiLU of the output row (needs BN == Cout): gamma [Cout] or null
  __half* norm_out;      // where silu(rms_norm(out)) goes (same addressing as out); `out` may then be null
  int epi_regs;                 // fused norm epilogue: 1 = single TMEM pass, row kept in registers (BN = 96)
  int kps;                      // k-blocks per ring slot (3 for Cin = 96: one whole tap per slot, 6 MMAs per barrier trip)
  uint32_t a_block_bytes, b_block_bytes;
};

template <int BK>
struct ConvCfg {
  static constexpr uint32_t kRowBytes = BK * 2;
  static constexpr uint32_t kLayout = BK == 64 ? 2u : (BK == 32 ? 4u : 6u);  // 128B / 64B / 32B swizzle
  static constexpr uint32_t kSBO = 8 * kRowBytes;
  static constexpr uint32_t kABytes = kConvBM * kRowBytes;
};

// Fused RMS_norm + SiLU of the output row (vae.py:39-54,194-199): the whole channel row of this voxel sits in this
// thread's TMEM lane, so the consumer's normalised input is produced here and the stand-alone normalisation pass
// over HBM disappears.  Pass 1: bias / residual, sum of squares of the value AS STORED (fp16-rounded), optional raw
// store; pass 2: re-read TMEM, normalise, SiLU, store.  FULL: BN is a multiple of 32 (every chunk has 32 columns):
// vector loads of bias / gamma, no per-element predicates.
template <bool FULL>
__device__ __forceinline__ void conv_epilogue_fused(const ConvParams& p, uint32_t taddr, int64_t off, bool ok) {
  float ss = 0.f;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const float inv = pass == 0 ? 0.f : sqrtf(static_cast<float>(p.Cout)) / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll 1
    for (int c = 0; c < p.BN; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(taddr + c, v);
      tc_wait_ld();
      const int ncols = FULL ? 32 : min(32, p.BN - c);
      if (!ok) continue;
      float f[32];
      if (FULL && p.bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c) + i);
          f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b4.x;
          f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
          f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
          f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float b = (p.bias != nullptr && i < ncols) ? __ldg(p.bias + c + i) : 0.f;
          f[i] = i < ncols ? __uint_as_float(v[i]) + b : 0.f;
        }
      }
      if (p.res != nullptr) {
        const uint4* r4 = reinterpret_cast<const uint4*>(p.res + off + c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (8 * i < ncols) {
            const uint4 q = r4[i];
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              f[8 * i + 2 * k] += f16_lo(u[k]);
              f[8 * i + 2 * k + 1] += f16_hi(u[k]);
            }
          }
        }
      }
      __half* dst = nullptr;
      if (pass == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          f[i] = f16_round(f[i]);      // statistics of the value as it is stored
          ss += f[i] * f[i];
        }
        if (p.out != nullptr) dst = reinterpret_cast<__half*>(p.out) + off + c;
      } else {
        if (FULL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.norm_gamma + c) + i);
            const float gg[4] = {g4.x * inv, g4.y * inv, g4.z * inv, g4.w * inv};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float y = f16_round(f[4 * i + k]) * gg[k];
              f[4 * i + k] = __fdividef(y, 1.f + __expf(-y));
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float gmm = i < ncols ? __ldg(p.norm_gamma + c + i) : 0.f;
            const float y = f16_round(f[i]) * (gmm * inv);
            f[i] = __fdividef(y, 1.f + __expf(-y));
          }
        }
        dst = p.norm_out + off + c;
      }
      if (dst != nullptr) {
        uint4* o4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (8 * i < ncols) {
            uint4 q;
            q.x = pack_f16(f[8 * i + 0], f[8 * i + 1]);
            q.y = pack_f16(f[8 * i + 2], f[8 * i + 3]);
            q.z = pack_f16(f[8 * i + 4], f[8 * i + 5]);
            q.w = pack_f16(f[8 * i + 6], f[8 * i + 7]);
            o4[i] = q;
          }
        }
      }
    }
  }
}

// Hands an accumulator tile back to the MMA warp: local barrier (rel_cta < 0) or the barrier of CTA rel_cta of the pair.
__device__ __forceinline__ void conv_release_tile(uint64_t* rel_bar, int rel_cta) {
  tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    if (rel_cta < 0) mbar_arrive(rel_bar);
    else mbar_arrive_cluster(rel_bar, static_cast<uint32_t>(rel_cta));
  }
}

// The same fused epilogue with ONE pass over TMEM for BN = 96 (stage D, where the two-pass epilogue costs the most): the
// row is kept in 48 registers as packed fp16 — exactly the values that are stored, which is what the statistics are
// defined on — so the accumulator buffer goes back to the MMA warp before the normalisation / SiLU / store pass
// starts, TMEM is read once and bias / residual are applied once.  A real call (__noinline__): its register allocation
// must not leak into the MMA-issue path of the calling kernel.
struct FusedRowArgs {
  const float* bias;
  const __half* res;
  __half* out;
  __half* norm_out;
  const float* gamma;
  float sqrt_c;
};

// Column split (pair kernel, one tile per CTA): the two epilogue warp sets each take half of a 192-channel row; the
// partial sums of squares meet in shared memory (ss_mine / ss_other, named barrier 1 over the 8 epilogue warps).
template <int NCH>
__device__ __noinline__ void conv_epilogue_fused_regs(const FusedRowArgs a, uint32_t taddr, int64_t off, bool ok,
                                                      uint64_t* rel_bar, int rel_cta, float* ss_mine,
                                                      const float* ss_other) {
  uint32_t h[NCH * 16];
  float ss = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int c = ch * 32;
    uint32_t vc[32];
    tmem_ld_x32(taddr + c, vc);
    tc_wait_ld();
    float f[32];
    if (a.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + c) + i);
        f[4 * i + 0] = __uint_as_float(vc[4 * i + 0]) + b4.x;
        f[4 * i + 1] = __uint_as_float(vc[4 * i + 1]) + b4.y;
        f[4 * i + 2] = __uint_as_float(vc[4 * i + 2]) + b4.z;
        f[4 * i + 3] = __uint_as_float(vc[4 * i + 3]) + b4.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(vc[i]);
    }
    if (a.res != nullptr && ok) {
      const uint4* r4 = reinterpret_cast<const uint4*>(a.res + off + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 q = r4[i];
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          f[8 * i + 2 * k] += f16_lo(u[k]);
          f[8 * i + 2 * k + 1] += f16_hi(u[k]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t pk = pack_f16(f[2 * i], f[2 * i + 1]);
      h[ch * 16 + i] = pk;
      const float lo = f16_lo(pk), hi = f16_hi(pk);      // statistics of the value as it is stored
      ss = fmaf(lo, lo, ss);
      ss = fmaf(hi, hi, ss);
    }
  }
  conv_release_tile(rel_bar, rel_cta);                   // the accumulator tile is free: the next MMAs may overwrite it
  if (ss_mine != nullptr) {                              // the other half of the row lives in the other warp set
    *ss_mine = ss;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    ss += *ss_other;
  }
  if (!ok) return;
  if (a.out != nullptr) {
    uint4* o4 = reinterpret_cast<uint4*>(a.out + off);
#pragma unroll
    for (int i = 0; i < NCH * 4; ++i) o4[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
  }
  const float inv = a.sqrt_c / fmaxf(sqrtf(ss), 1e-12f);
  uint4* n4 = reinterpret_cast<uint4*>(a.norm_out + off);
#pragma unroll
  for (int i = 0; i < NCH * 4; ++i) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma) + 2 * i);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma) + 2 * i + 1);
    const float gg[8] = {g0.x * inv, g0.y * inv, g0.z * inv, g0.w * inv, g1.x * inv, g1.y * inv, g1.z * inv, g1.w * inv};
    float y[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float u0 = f16_lo(h[4 * i + k]) * gg[2 * k], u1 = f16_hi(h[4 * i + k]) * gg[2 * k + 1];
      y[2 * k] = __fdividef(u0, 1.f + __expf(-u0));
      y[2 * k + 1] = __fdividef(u1, 1.f + __expf(-u1));
    }
    n4[i] = make_uint4(pack_f16(y[0], y[1]), pack_f16(y[2], y[3]), pack_f16(y[4], y[5]), pack_f16(y[6], y[7]));
  }
}

// The residual row of this thread's voxel is last-touched a whole conv ago (DRAM): ask for it while the tile's MMAs are
// still running, so that the epilogue's loads find it in L2 instead of serialising three DRAM round trips per tile
// (ncu: tensor pipe 63 % with the residual + two-output epilogue vs 75 % without, at a HIGHER clock).
__device__ __forceinline__ void conv_prefetch_res(const ConvParams& p, int n_blk, int t, int h, int w, int c0, int ncol) {
  if (p.res == nullptr || h >= p.H || w >= p.W) return;
  int n0 = n_blk * p.BN;
  int64_t off = p.o_base + t * p.os_t + h * p.os_h + w * p.os_w;
  if (p.nsplit > 0 && n0 >= p.nsplit) {
    n0 -= p.nsplit;
    off += p.nsplit_off;
  }
  const char* a = reinterpret_cast<const char*>(p.res + off + (p.norm_out != nullptr ? 0 : n0) + c0);
  for (int b = 0; b < ncol * 2; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + b));
}

// Epilogue of one accumulator tile for ONE output voxel (t, h, w) = this thread's TMEM lane: bias, residual, store —
// or, with norm_out, the consumer's RMS_norm + SiLU fused in (two passes over the TMEM row).  taddr = this warp's lane
// quadrant + the tile's first accumulator column.
// (rel_bar, rel_cta) hand the accumulator tile back to the MMA warp (conv_release_tile) as soon as TMEM has been read for
// the last time.
// [c0, c0 + ncol) = the accumulator columns this thread drains (the whole tile, or one half of it when the two warp sets
// of the pair kernel share a tile; then ss_mine / ss_other carry the fused norm's partial statistics).
__device__ __forceinline__ void conv_epilogue_tile(const ConvParams& p, uint32_t taddr, int n_blk, int t, int h, int w,
                                                   uint64_t* rel_bar, int rel_cta, int c0, int ncol, float* ss_mine,
                                                   const float* ss_other) {
  const bool ok = (h < p.H) && (w < p.W);
  int n0 = n_blk * p.BN;
  int64_t off = p.o_base + t * p.os_t + h * p.os_h + w * p.os_w;
  int nb = n0;  // channel offset inside the destination row
  if (p.nsplit > 0 && n0 >= p.nsplit) {
    nb = n0 - p.nsplit;
    off += p.nsplit_off;
  }
  if (p.norm_out != nullptr) {
    if (p.epi_regs && ncol == 96) {    // a 96-channel row, or half of a 192-channel one (a whole 192-channel row would need
                                       // 96 + 64 registers: ptxas spills ~900 B — measured, dropped)
      FusedRowArgs a;
      a.bias = p.bias != nullptr ? p.bias + c0 : nullptr;
      a.res = p.res;
      a.out = reinterpret_cast<__half*>(p.out);
      a.norm_out = p.norm_out;
      a.gamma = p.norm_gamma + c0;
      a.sqrt_c = sqrtf(static_cast<float>(p.Cout));
      conv_epilogue_fused_regs<3>(a, taddr + c0, off + c0, ok, rel_bar, rel_cta, ss_mine, ss_other);
      return;
    }
    if ((p.BN & 31) == 0) conv_epilogue_fused<true>(p, taddr, off, ok);
    else conv_epilogue_fused<false>(p, taddr, off, ok);
  } else
  for (int c = c0; c < c0 + ncol; c += 32) {   // BN multiple of 16: last chunk may be half valid
    uint32_t v[32];
    tmem_ld_x32(taddr + c, v);
    tc_wait_ld();
    const int ncols = min(32, p.BN - c);
    if (ok) {
      if (p.out_mode == 0) {
        float f[32];
        if (ncols == 32 && p.bias != nullptr) {       // whole chunk (n0, c multiples of 16/32, Cout % 16 == 0): vector loads
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c) + i);
            f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b4.x;
            f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b4.y;
            f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b4.z;
            f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b4.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float b = (p.bias != nullptr && i < ncols && n0 + c + i < p.Cout) ? __ldg(p.bias + n0 + c + i) : 0.f;
            f[i] = __uint_as_float(v[i]) + b;
          }
        }
        __half* o = reinterpret_cast<__half*>(p.out) + off + nb + c;
        if (p.res != nullptr) {
          const uint4* r4 = reinterpret_cast<const uint4*>(p.res + off + nb + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (8 * i < ncols) {
              const uint4 q = r4[i];
              const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                f[8 * i + 2 * k] += f16_lo(u[k]);
                f[8 * i + 2 * k + 1] += f16_hi(u[k]);
              }
            }
          }
        }
        uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (8 * i < ncols) {
            uint4 q;
            q.x = pack_f16(f[8 * i + 0], f[8 * i + 1]);
            q.y = pack_f16(f[8 * i + 2], f[8 * i + 3]);
            q.z = pack_f16(f[8 * i + 4], f[8 * i + 5]);
            q.w = pack_f16(f[8 * i + 6], f[8 * i + 7]);
            o4[i] = q;
          }
        }
      } else {
        // head conv: fp32 channel-first video, clamp(-1, 1)                       vae.py:660-661
        float* o = reinterpret_cast<float*>(p.out);
        const int64_t plane = p.os_t > 0 ? p.os_t : static_cast<int64_t>(p.T) * p.H * p.W;   // channel stride
        const int64_t pos = p.o_base + (static_cast<int64_t>(t) * p.H + h) * p.W + w;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (c == 0 && i < p.cout_real) {
            float b = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
            float y = __uint_as_float(v[i]) + b;
            y = fminf(1.f, fmaxf(-1.f, y));
            o[i * plane + pos] = y;
          }
        }
      }
    }
  }
  conv_release_tile(rel_bar, rel_cta);
}

template <int BK, bool HI>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvParams p) {
  using Cfg = ConvCfg<BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kConvStages = p.stages;
  const uint32_t kConvABytesMax = p.a_stage_bytes, kConvBBytesMax = p.b_stage_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + kConvStages * kConvABytesMax;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kConvRingBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kConvMaxStages;
  uint64_t* tfull = bars + 2 * kConvMaxStages;
  uint64_t* tempty = bars + 2 * kConvMaxStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kConvMaxStages + 4);

  const int pw = threadIdx.x >> 5;                      // physical warp: TMEM lane quadrant = pw & 3
  const int warp = HI ? (pw + 2) % 6 : pw;              // role id: 0 TMA, 1 MMA, 2-5 epilogue (HI: roles on warps 4, 5)
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * Cfg::kRowBytes;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kConvStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_total = p.ntaps * p.kblocks_per_tap;

  // tile -> (n block fastest, then w, h, t): CTAs running together share the same input patch through L2
  auto decode = [&](int tile, int& n_blk, int& t, int& h0, int& w0) {
    n_blk = tile % p.num_n;
    int r = tile / p.num_n;
    w0 = (r % p.tiles_w) * kConvTW;
    r /= p.tiles_w;
    h0 = (r % p.tiles_h) * kConvTH;
    t = r / p.tiles_h;
  };

  if (warp == 0) {
    // TMA producer: warp-uniform loop, elect.sync around the issue only
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int n_blk, t, h0, w0;
      decode(tile, n_blk, t, h0, w0);
      for (int kb = 0; kb < kb_total; kb += p.kps) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], static_cast<uint32_t>(p.kps) * (Cfg::kABytes + b_bytes));
          for (int i = 0; i < p.kps; ++i) {
            const int tap = (kb + i) / p.kblocks_per_tap;
            const int cb = (kb + i) - tap * p.kblocks_per_tap;
            tma_load_4d(sA + stage * kConvABytesMax + i * p.a_block_bytes, &tmA, &full[stage], cb * BK,
                        p.sw * w0 + p.dw[tap], p.sh * h0 + p.dh[tap], p.st * t + p.t_off + p.dt[tap]);
            tma_load_2d(sB + stage * kConvBBytesMax + i * p.b_block_bytes, &tmB, &full[stage],
                        tap * p.Cin + cb * BK, n_blk * p.BN);
          }
        }
        __syncwarp();
        if (++stage == kConvStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(kConvBM, static_cast<uint32_t>(p.BN), 0, 0);
    const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, Cfg::kSBO, Cfg::kLayout);
    const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, Cfg::kSBO, Cfg::kLayout);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * 256;
      for (int kb = 0; kb < kb_total; kb += p.kps) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          for (int i = 0; i < p.kps; ++i) {
            const uint64_t adesc = adesc0 + ((stage * kConvABytesMax + i * p.a_block_bytes) >> 4);
            const uint64_t bdesc = bdesc0 + ((stage * kConvBBytesMax + i * p.b_block_bytes) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | i | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb + p.kps >= kb_total) umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == kConvStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    const int quad = pw & 3;
    int as = 0;
    uint32_t aphase = 0;
    const int r = quad * 32 + lane;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int n_blk, t, h0, w0;
      decode(tile, n_blk, t, h0, w0);
      conv_prefetch_res(p, n_blk, t, h0 + (r >> 4), w0 + (r & 15), 0, p.BN);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 256;
      conv_epilogue_tile(p, taddr, n_blk, t, h0 + (r >> 4), w0 + (r & 15), &tempty[as], -1, 0, p.BN, nullptr, nullptr);
      as ^= 1;
      if (as == 0) aphase ^= 1;
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
// conv_igemm_pair_kernel — the same implicit GEMM restructured around its real limit.  ncu of the kernel above
// (profiles/r02_ncu_vae_conv.txt): every (tap, channel block) re-fetches its 128-voxel A box and its weight slab from
// L2 — 104 B per tensor-pipe clock per SM for Cout = 192, 146 B/clk for Cout = 96 — and the L2 -> SM path delivers
// ~14 TB/s (51 % of its peak, latency-bound with ~200 KB in flight per SM): tensor pipe 60 % / 29 % active.
// Here the bytes per FLOP are cut 3-4x:
//   * a CTA owns NT vertically stacked 8 x 16 voxel tiles and loads, per (channel block, dt, dw), ONE box of
//     (8 NT + 2) x 16 voxels; the three vertical taps dh = -1, 0, +1 of all NT tiles are row-offset views of that box
//     (16 rows = whole swizzle atoms, so a view is just a descriptor start address) — 2.4-2.9x fewer A bytes;
//   * two CTAs (a cluster on one TPC) issue ONE tcgen05.mma.cta_group::2 over M = 256 = tile j of both CTAs; each CTA
//     stages HALF of the weight slab, and a slab serves all NT tile pairs — 2 NT x fewer B bytes per voxel.
// One pipeline stage = one A box + the (<= 3) weight slabs of its vertical taps = 3 x NT x BK/16 MMAs (24 for the
// decoder's heavy stages).  Accumulators: NT x BN columns per buffer; two buffers when they fit in 512 TMEM columns,
// otherwise one buffer released tile by tile so the next super-tile's MMAs follow the epilogue.
//   warp 0: TMA producer (both CTAs)   warp 1: MMA issuer (leader)   warps 2-9: two epilogue warp sets (tiles j = 0, 2 /
//   1, 3 ...), each thread one voxel row of its tile (conv_epilogue_tile above).
// ================================================================================================================
constexpr int kConv2Threads = 320;
constexpr int kDefaultConvPair = 1;   // measured: 1080P decode 47 -> 73 fps (profiles/r02_vae_*); MV_CONV_PAIR=0 = single-CTA kernel
constexpr int kConv2MaxGroups = 9;
constexpr int kConv2MaxStages = 6;
constexpr uint32_t kConv2RingBytes = 216 * 1024;
constexpr uint32_t kConv2Smem = kConv2RingBytes + 1024 + 512 + 2048;   // + [2][2][128] floats: row statistics of a split tile

struct Conv2Params {
  ConvParams c;
  int ngroups;                              // distinct (dt, dw) pairs of the tap set
  int8_t g_dt[kConv2MaxGroups], g_dw[kConv2MaxGroups], g_ntap[kConv2MaxGroups];
  int8_t g_tap[kConv2MaxGroups][3];         // index of the tap in the packed weight matrix
  int8_t g_dhoff[kConv2MaxGroups][3];       // dh - dh_min: row-group offset of the tap's view inside the A box
  int dh_min, box_h;                        // A box: h from (tile origin + dh_min), box_h = 8 NT + dh_max - dh_min rows
  uint32_t a_box_bytes, b_slab_bytes, stage_bytes;   // 1024-aligned
  int stages, nbuf;
  int split_cols;                           // NT = 1: both epilogue warp sets drain the one tile, half of its columns each
  int regular;                              // n > 0: every group = n vertical taps dh = dh_min, dh_min + 1, ... in order
  int sup_h, sup_w, num_super;              // super-tiles (pair = 16 NT x 16 voxels) per frame in h / w; total incl. t, n
};

// One pipeline stage of the pair kernel for a regular tap group: NTAP vertical taps x NT tiles x BK/16 k-steps, issued by
// the elected thread as straight-line code (every offset a compile-time constant), then the stage / accumulator commits.
template <int BK, int NT, int NTAP>
__device__ __forceinline__ void conv_issue_stage(uint64_t a0, uint64_t b0, uint32_t bstep, uint32_t dbase, int BN,
                                                 uint32_t idesc, uint32_t acc0, uint64_t* empty_bar, uint64_t* tfull_bar) {
  constexpr uint32_t kRowGroupBytes = kConvTW * ConvCfg<BK>::kRowBytes;   // one h row of the box = 16 voxels
  if (elect_one()) {
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const uint32_t d_tmem = dbase + static_cast<uint32_t>(j * BN);
#pragma unroll
      for (int i = 0; i < NTAP; ++i) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_ss_pair(d_tmem, a0 + (((j * kConvTH + i) * kRowGroupBytes) >> 4) + 2 * k, b0 + i * bstep + 2 * k, idesc,
                       (i | k) != 0 ? 1u : acc0);
      }
    }
    umma_commit_pair(empty_bar, 3);
    if (tfull_bar != nullptr) umma_commit_pair(tfull_bar, 3);
  }
  __syncwarp();
}

template <int BK, int NT, bool HI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConv2Threads, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const Conv2Params q) {
  using Cfg = ConvCfg<BK>;
  const ConvParams& p = q.c;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kConv2RingBytes);
  uint64_t* full = bars;                                  // leader only
  uint64_t* empty = bars + kConv2MaxStages;               // both CTAs (multicast commit)
  uint64_t* tfull = bars + 2 * kConv2MaxStages;           // [nbuf], both CTAs (multicast commit)
  uint64_t* tempty = bars + 2 * kConv2MaxStages + 2;      // [nbuf][NT], leader only: 4 warps x 2 CTAs arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kConv2MaxStages + 2 + 2 * NT);
  float* ssx = reinterpret_cast<float*>(smem + kConv2RingBytes + 512);   // [parity][warp set][row]
  const bool split = (NT == 1) && q.split_cols != 0;

  const int pw = threadIdx.x >> 5;                      // physical warp: TMEM lane quadrant = pw & 3
  const int warp = HI ? (pw + 2) % 10 : pw;             // role id: 0 TMA, 1 MMA, 2-9 epilogue (HI: roles on warps 8, 9)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int nstages = q.stages;
  const int nbuf = q.nbuf;
  const int half_n = p.BN >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < nstages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&tfull[i], 1);
    for (int i = 0; i < 2 * NT; ++i) mbar_init(&tempty[i], split ? 16 : 8);   // arriving warps: 4 (8 if split) x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int steps_per_super = p.kblocks_per_tap * q.ngroups;

  // super-tile -> (n block fastest, then w, then t, then h): a band of rows is swept through ALL frames before the
  // next band starts, so the two history frames a temporal tap re-reads are still in L2 (ncu on the (n, w, h, t)
  // order: every input frame came from DRAM three times, 21 GB read for 10.4 GB of operands on the stage-D conv)
  auto decode = [&](int st, int& n_blk, int& t, int& h0, int& w0) {
    n_blk = st % p.num_n;
    int r = st / p.num_n;
    w0 = (r % q.sup_w) * kConvTW;
    r /= q.sup_w;
    t = r % p.T;
    h0 = (r / p.T) * (2 * NT * kConvTH) + static_cast<int>(rank) * (NT * kConvTH);   // this CTA's first row
  };

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int st = pair_id; st < q.num_super; st += num_pairs) {
      int n_blk, t, h0, w0;
      decode(st, n_blk, t, h0, w0);
      for (int cb = 0; cb < p.kblocks_per_tap; ++cb) {
        for (int g = 0; g < q.ngroups; ++g) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            const int ntap = q.g_ntap[g];
            uint8_t* base = smem + stage * q.stage_bytes;
            if (leader) mbar_expect_tx(&full[stage], 2u * (static_cast<uint32_t>(q.box_h) * kConvTW * Cfg::kRowBytes +
                                                           static_cast<uint32_t>(ntap) * half_n * Cfg::kRowBytes));
            tma_load_4d_pair(base, &tmA, &full[stage], cb * BK, w0 + q.g_dw[g], h0 + q.dh_min,
                             t + p.t_off + q.g_dt[g]);
            for (int i = 0; i < ntap; ++i)
              tma_load_2d_pair(base + q.a_box_bytes + i * q.b_slab_bytes, &tmB, &full[stage],
                               q.g_tap[g][i] * p.Cin + cb * BK, n_blk * p.BN + static_cast<int>(rank) * half_n);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA) ------------------------------
    if (leader) {
      const uint32_t idesc = make_idesc_f16(2 * kConvBM, static_cast<uint32_t>(p.BN), 0, 0);
      const uint64_t desc0 = make_smem_desc(smem_u32(smem), 16, Cfg::kSBO, Cfg::kLayout);
      constexpr uint32_t kRowGroupBytes = kConvTW * Cfg::kRowBytes;   // one h row of the box = 16 voxels
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int st = pair_id; st < q.num_super; st += num_pairs, ++it) {
        const int buf = it % nbuf;
        const uint32_t use = static_cast<uint32_t>(it / nbuf);
        uint32_t started = 0;    // bit j: tile j has received its first MMA of this super-tile
        if (q.regular > 0) {
          // Regular tap sets (3x3x3: 3 vertical taps per (dt, dw) group; sub-pixel 2x2: 2; time_conv: 1 — always
          // dh = dh_min, dh_min + 1, ... in order): all operand offsets inside a stage are compile-time constants and
          // the issue sequence is straight-line code on uniform registers (NTAP x NT x BK/16 back-to-back UTCHMMAs per
          // stage).  With 48-clock MMAs (N = 96) the general loop below (indexed constant loads + per-thread address
          // math + R2UR per MMA) could not keep the tensor pipe fed: 36 % -> 70 % tensor-pipe active on stage D.
          const uint32_t dbase = tmem_base + static_cast<uint32_t>(buf * NT * p.BN);
          for (int step = 0; step < steps_per_super; ++step) {
            mbar_wait(&full[stage], phase);
            if (step == 0) {
              for (int j = 0; j < NT; ++j) mbar_wait(&tempty[buf * NT + j], (use & 1u) ^ 1u);   // accumulators drained
            }
            tc_fence_after();
            const uint64_t a0 = desc0 + ((static_cast<uint32_t>(stage) * q.stage_bytes) >> 4);
            const uint64_t b0 = a0 + (q.a_box_bytes >> 4);
            const uint32_t acc0 = step > 0 ? 1u : 0u;
            const bool last = step == steps_per_super - 1;
            if (q.regular == 3) conv_issue_stage<BK, NT, 3>(a0, b0, q.b_slab_bytes >> 4, dbase, p.BN, idesc, acc0, &empty[stage], last ? &tfull[buf] : nullptr);
            else if (q.regular == 2) conv_issue_stage<BK, NT, 2>(a0, b0, q.b_slab_bytes >> 4, dbase, p.BN, idesc, acc0, &empty[stage], last ? &tfull[buf] : nullptr);
            else conv_issue_stage<BK, NT, 1>(a0, b0, q.b_slab_bytes >> 4, dbase, p.BN, idesc, acc0, &empty[stage], last ? &tfull[buf] : nullptr);
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
          }
          continue;
        }
        for (int step = 0; step < steps_per_super; ++step) {
          const int g = step % q.ngroups;
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const int ntap = q.g_ntap[g];
          const uint32_t sbase = static_cast<uint32_t>(stage) * q.stage_bytes;
#pragma unroll 1
          for (int j = 0; j < NT; ++j) {
            if (!((started >> j) & 1u)) {
              // the accumulator of tile j must have been drained by the epilogue of its previous use
              mbar_wait(&tempty[buf * NT + j], (use & 1u) ^ 1u);
              tc_fence_after();
            }
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((buf * NT + j) * p.BN);
              for (int i = 0; i < ntap; ++i) {
                const uint64_t adesc = desc0 + ((sbase + static_cast<uint32_t>(j * kConvTH + q.g_dhoff[g][i]) * kRowGroupBytes) >> 4);
                const uint64_t bdesc = desc0 + ((sbase + q.a_box_bytes + static_cast<uint32_t>(i) * q.b_slab_bytes) >> 4);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)
                  umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc,
                               (((started >> j) & 1u) | static_cast<uint32_t>(i | k)) != 0 ? 1u : 0u);
              }
            }
            __syncwarp();
            started |= 1u << j;
          }
          if (elect_one()) {
            umma_commit_pair(&empty[stage], 3);
            if (step == steps_per_super - 1) umma_commit_pair(&tfull[buf], 3);
          }
          __syncwarp();
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ------------------------------ epilogue: two warp sets, tiles j = set, set + 2, ... ------------------------------
    const int wset = (warp - 2) >> 2;
    const int quad = pw & 3;
    const int r = quad * 32 + lane;
    const int half0 = ((p.BN + 63) >> 6) << 5;     // split tiles: set 0 takes whole 32-column chunks covering >= half the row
    int it = 0;
    for (int st = pair_id; st < q.num_super; st += num_pairs, ++it) {
      const int buf = it % nbuf;
      const uint32_t use = static_cast<uint32_t>(it / nbuf);
      int n_blk, t, h0, w0;
      decode(st, n_blk, t, h0, w0);
      if (split) conv_prefetch_res(p, n_blk, t, h0 + (r >> 4), w0 + (r & 15), wset * half0, wset == 0 ? half0 : p.BN - half0);
      else
        for (int j = wset; j < NT; j += 2) conv_prefetch_res(p, n_blk, t, h0 + j * kConvTH + (r >> 4), w0 + (r & 15), 0, p.BN);
      mbar_wait(&tfull[buf], use & 1u);
      tc_fence_after();
      if (split) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(buf * p.BN);
        float* mine = ssx + ((it & 1) * 2 + wset) * 128 + r;
        const float* other = ssx + ((it & 1) * 2 + (wset ^ 1)) * 128 + r;
        conv_epilogue_tile(p, taddr, n_blk, t, h0 + (r >> 4), w0 + (r & 15), &tempty[buf], 0, wset * half0,
                           wset == 0 ? half0 : p.BN - half0, p.norm_out != nullptr ? mine : nullptr, other);
        continue;
      }
      for (int j = wset; j < NT; j += 2) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>((buf * NT + j) * p.BN);
        conv_epilogue_tile(p, taddr, n_blk, t, h0 + j * kConvTH + (r >> 4), w0 + (r & 15), &tempty[buf * NT + j], 0, 0,
                           p.BN, nullptr, nullptr);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// --------------------------------------------------------------------------------------------
// RMS_norm over channels + SiLU, channels-last fp16 -> fp16 (in place allowed).   vae.py:39-54 + nn.SiLU
// y = silu( x / max(||x||_2, 1e-12) * sqrt(C) * gamma (+ beta) );  one warp per voxel, C <= 512, C % 8 == 0.
// silu == 0 skips the activation (AttentionBlock.norm, vae.py:233,246).
// --------------------------------------------------------------------------------------------
// G lanes cooperate on one voxel (G = 16 for C <= 128, else 32), so a warp keeps 32/G voxels and every lane in
// flight; two voxels per group are processed per iteration for memory-level parallelism.
template <int G, int VPL>  // VPL = uint4 vectors per lane (C <= G * VPL * 8)
__global__ void __launch_bounds__(256)
rmsnorm_silu_cl_kernel(const __half* __restrict__ x, __half* __restrict__ y, const float* __restrict__ gamma,
                       int64_t nvox, int C, int silu) {
  constexpr int kUnroll = 2;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);                  // lane inside the group
  // the loop bound is WARP-uniform (full-mask shuffles inside): a warp covers (32/G)*kUnroll consecutive voxels
  constexpr int kPerWarp = (32 / G) * kUnroll;
  const int64_t warp_id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int gw = lane / G;                        // group inside the warp
  const int nvec = C >> 3;
  const float scale = sqrtf(static_cast<float>(C));
  float g[VPL][8];
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    const int idx = gl + G * j;
    if (idx < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx);
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * idx + 1);
      g[j][0] = g0.x; g[j][1] = g0.y; g[j][2] = g0.z; g[j][3] = g0.w;
      g[j][4] = g1.x; g[j][5] = g1.y; g[j][6] = g1.z; g[j][7] = g1.w;
    }
  }
  for (int64_t vw = warp_id * kPerWarp; vw < nvox; vw += nwarps * kPerWarp) {
    const int64_t v0 = vw + gw * kUnroll;
    uint4 a[kUnroll][VPL];
    float ss[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      ss[u] = 0.f;
      const int64_t vox = v0 + u;
      const uint4* xr = reinterpret_cast<const uint4*>(x + vox * C);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        a[u][j] = make_uint4(0, 0, 0, 0);
        if (vox < nvox && idx < nvec) a[u][j] = xr[idx];
        const uint32_t w[4] = {a[u][j].x, a[u][j].y, a[u][j].z, a[u][j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = f16_lo(w[k]), hi = f16_hi(w[k]);
          ss[u] += lo * lo + hi * hi;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], o);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t vox = v0 + u;
      const float inv = scale / fmaxf(sqrtf(ss[u]), 1e-12f);
      uint4* yr = reinterpret_cast<uint4*>(y + vox * C);
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        const int idx = gl + G * j;
        if (vox < nvox && idx < nvec) {
          uint32_t w[4] = {a[u][j].x, a[u][j].y, a[u][j].z, a[u][j].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float lo = f16_lo(w[k]) * inv * g[j][2 * k];
            float hi = f16_hi(w[k]) * inv * g[j][2 * k + 1];
            if (silu) {
              lo = __fdividef(lo, 1.f + __expf(-lo));   // as in the fused conv epilogue (an IEEE divide made the kernel
              hi = __fdividef(hi, 1.f + __expf(-hi));   // issue-bound: ncu 2.7 TB/s at 71 % issue-slot use)
            }
            w[k] = pack_f16(lo, hi);
          }
          yr[idx] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// Latent de-normalisation + conv2 (1x1x1, 16 -> 16) + layout change to channels-last fp16.
// x[t,h,w,o] = sum_c W2[o,c] * (z[c,t,h,w] * std[c] + mean[c]) + b2[o]                vae.py:547-553
// --------------------------------------------------------------------------------------------
__global__ void vae_latent_in_kernel(const float* __restrict__ z, const float* __restrict__ W2, const float* __restrict__ b2,
                                     const float* __restrict__ mean, const float* __restrict__ stdv,
                                     __half* __restrict__ out, int Z, int64_t nvox) {
  __shared__ float sw[32 * 32 + 96];
  float* sb = sw + Z * Z;
  float* sm = sb + Z;
  float* ss = sm + Z;
  for (int i = threadIdx.x; i < Z * Z; i += blockDim.x) sw[i] = W2[i];
  for (int i = threadIdx.x; i < Z; i += blockDim.x) {
    sb[i] = b2[i];
    sm[i] = mean[i];
    ss[i] = stdv[i];
  }
  __syncthreads();
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvox;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float zin[32];
    for (int c = 0; c < Z; ++c) zin[c] = z[c * nvox + v] * ss[c] + sm[c];
    for (int o = 0; o < Z; ++o) {
      float acc = sb[o];
      for (int c = 0; c < Z; ++c) acc = fmaf(sw[o * Z + c], zin[c], acc);
      out[v * Z + o] = __float2half_rn(f16_sat(acc));
    }
  }
}

// --------------------------------------------------------------------------------------------
// Encoder edges.  video_in: frames [t0, t0+n) of a channel-first fp32 video [3, T, H, W] -> channels-last fp16
// [n, H, W, 16] (channels 3..15 zero: the 16-channel K block of the stem conv, vae.py:286).
// latent_out: conv1 (1x1x1, 2Z -> 2Z; only the mu half is computed) on the fp16 head output [nvox, 2Z], then
// mu = (mu - mean) * (1 / std), written channel-first fp32                                   vae.py:531-537
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vae_video_in_kernel(const float* __restrict__ video, int64_t chan_stride, int64_t base, int64_t nvox,
                    uint4* __restrict__ out) {
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvox;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float r = video[base + v], g = video[chan_stride + base + v], b = video[2 * chan_stride + base + v];
    uint4 lo;
    lo.x = pack_f16(r, g);
    lo.y = pack_f16(b, 0.f);
    lo.z = 0u;
    lo.w = 0u;
    out[2 * v] = lo;
    out[2 * v + 1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

__global__ void __launch_bounds__(256)
vae_latent_out_kernel(const __half* __restrict__ head, const float* __restrict__ W1, const float* __restrict__ b1,
                      const float* __restrict__ mean, const float* __restrict__ inv_std, float* __restrict__ mu, int Z,
                      int64_t nvox, int64_t mu_plane, int64_t mu_off) {
  __shared__ float sw[16 * 32 + 48];
  const int C = 2 * Z;
  float* sb = sw + Z * C;
  float* sm = sb + Z;
  float* ss = sm + Z;
  for (int i = threadIdx.x; i < Z * C; i += blockDim.x) sw[i] = W1[i];      // rows 0..Z-1 of the [2Z, 2Z] matrix = mu
  for (int i = threadIdx.x; i < Z; i += blockDim.x) {
    sb[i] = b1[i];
    sm[i] = mean[i];
    ss[i] = inv_std[i];
  }
  __syncthreads();
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvox;
       v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float x[32];
    for (int c = 0; c < C; ++c) x[c] = __half2float(head[v * C + c]);
    for (int o = 0; o < Z; ++o) {
      float acc = sb[o];
      for (int c = 0; c < C; ++c) acc = fmaf(sw[o * C + c], x[c], acc);
      mu[o * mu_plane + mu_off + v] = __fmul_rn(__fsub_rn(acc, sm[o]), ss[o]);
    }
  }
}

// --------------------------------------------------------------------------------------------
// Head conv (96 -> 3, 3x3x3) without 162 sixteen-column MMAs per voxel tile: the conv is linear, so
//   out[co, v] = bias[co] + sum_tap D[v + tap][tap][co]      with      D[u][tap][co] = in[u, :] . W[co][tap][:]
// D is ONE 1x1x1 convolution 96 -> 27 taps x 4 (3 real channels + pad) = 112 columns on the implicit-GEMM kernel; this
// kernel is the gather: per output voxel 27 8-byte loads of its neighbours' partial sums (fp16, fp32 accumulation, fixed
// order), bias, clamp(-1, 1), fp32 channel-first store (vae.py:660-661).  Frames before the chunk come from d_prev
// (the last kprev <= 2 frames of D of the previous chunk); frames before the sequence and voxels outside the image
// contribute zero (causal / spatial zero padding, vae.py:17-36).
// --------------------------------------------------------------------------------------------
constexpr int kHeadLd = 112;
__global__ void __launch_bounds__(256)
vae_head_gather_kernel(const __half* __restrict__ d_cur, const __half* __restrict__ d_prev, int kprev, int H, int W,
                       float b0, float b1, float b2, float* __restrict__ video, int64_t plane, int64_t frame_off) {
  const int w = blockIdx.x * 32 + (threadIdx.x & 31);
  const int h = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int t = blockIdx.z;
  if (w >= W || h >= H) return;
  const int64_t fvox = static_cast<int64_t>(H) * W;
  float a0 = b0, a1 = b1, a2 = b2;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int tt = t + it - 2;
    const __half* frame;
    if (tt >= 0) frame = d_cur + tt * fvox * kHeadLd;
    else if (kprev + tt >= 0) frame = d_prev + (kprev + tt) * fvox * kHeadLd;
    else continue;
#pragma unroll
    for (int ih = 0; ih < 3; ++ih) {
      const int hh = h + ih - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int iw = 0; iw < 3; ++iw) {
        const int ww = w + iw - 1;
        if (ww < 0 || ww >= W) continue;
        const int tap = (it * 3 + ih) * 3 + iw;
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(frame + (static_cast<int64_t>(hh) * W + ww) * kHeadLd + tap * 4));
        a0 += f16_lo(q.x);
        a1 += f16_hi(q.x);
        a2 += f16_lo(q.y);
      }
    }
  }
  const int64_t pos = frame_off + (static_cast<int64_t>(t) * H + h) * W + w;
  video[pos] = fminf(1.f, fmaxf(-1.f, a0));
  video[plane + pos] = fminf(1.f, fmaxf(-1.f, a1));
  video[2 * plane + pos] = fminf(1.f, fmaxf(-1.f, a2));
}

// --------------------------------------------------------------------------------------------
// Row softmax for the VAE's single-head attention: P = softmax(S * scale), S fp32 [M, N] -> P fp16 [M, ldp].
// One CTA per row.                                                                 vae.py:246-257
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, int64_t lds, __half* __restrict__ P, int64_t ldp, int N,
                    float scale) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float* s = S + row * lds;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, s[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sum += __expf((s[i] - mx) * scale);
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  __half* pr = P + row * ldp;
  for (int i = threadIdx.x; i < N; i += blockDim.x) pr[i] = __float2half_rn(__expf((s[i] - mx) * scale) * inv);
}

// One pass: the whole row (N <= 1024 * 4 * kSmxVec floats) lives in the registers of a 1024-thread CTA — S is read
// once (float4), P written once; the three-pass kernel above stays for longer rows.
constexpr int kSmxVec = 8;   // float4 per thread -> N <= 32768
__global__ void __launch_bounds__(1024)
softmax_rows_reg_kernel(const float* __restrict__ S, int64_t lds, __half* __restrict__ P, int64_t ldp, int N, float scale) {
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const float4* s4 = reinterpret_cast<const float4*>(S + row * lds);
  const int nvec = N >> 2;
  float4 v[kSmxVec];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kSmxVec; ++i) {
    const int idx = threadIdx.x + i * 1024;
    v[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (idx < nvec) v[i] = s4[idx];
    mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  const float sl2 = scale * 1.4426950408889634f;
  const float nm = -mx * sl2;
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kSmxVec; ++i) {
    v[i].x = fast_exp2(fmaf(v[i].x, sl2, nm));
    v[i].y = fast_exp2(fmaf(v[i].y, sl2, nm));
    v[i].z = fast_exp2(fmaf(v[i].z, sl2, nm));
    v[i].w = fast_exp2(fmaf(v[i].w, sl2, nm));
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);      // exp2(-inf) = 0 for the padding lanes
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = warp_sum(red[lane]);
  const float inv = 1.f / sum;
  uint2* p2 = reinterpret_cast<uint2*>(P + row * ldp);
#pragma unroll
  for (int i = 0; i < kSmxVec; ++i) {
    const int idx = threadIdx.x + i * 1024;
    if (idx < nvec) p2[idx] = make_uint2(pack_f16(v[i].x * inv, v[i].y * inv), pack_f16(v[i].z * inv, v[i].w * inv));
  }
}

template <int BK>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kConvSmem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kConvSmem)));
    attr_set = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (roles_hi()) conv_igemm_kernel<BK, true><<<grid, kConvThreads, kConvSmem, st>>>(tmA, tmB, p);
  else conv_igemm_kernel<BK, false><<<grid, kConvThreads, kConvSmem, st>>>(tmA, tmB, p);
  MV_CHECK_LAUNCH("conv_igemm_kernel");
  return MV_OK;
}

template <int BK, int NT>
static int launch_conv_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const Conv2Params& q, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_pair_kernel<BK, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kConv2Smem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_pair_kernel<BK, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kConv2Smem)));
    attr_set = true;
  }
  int pairs = sm_count() / 2;
  if (q.num_super < pairs) pairs = q.num_super;
  if (roles_hi()) conv_igemm_pair_kernel<BK, NT, true><<<2 * pairs, kConv2Threads, kConv2Smem, st>>>(tmA, tmB, q);
  else conv_igemm_pair_kernel<BK, NT, false><<<2 * pairs, kConv2Threads, kConv2Smem, st>>>(tmA, tmB, q);
  MV_CHECK_LAUNCH("conv_igemm_pair_kernel");
  return MV_OK;
}

}  // namespace mv

using namespace mv;

namespace {
int g_conv_pair = -1;   // -1: read MV_CONV_PAIR on first use
int g_conv_nt = -1;     // -1: read MV_CONV_NT on first use; 0 = automatic
bool conv_pair_enabled() {
  if (g_conv_pair < 0) {
    const char* e = getenv("MV_CONV_PAIR");
    g_conv_pair = (e != nullptr && e[0] != 0) ? (atoi(e) != 0 ? 1 : 0) : kDefaultConvPair;
  }
  return g_conv_pair == 1;
}
}  // namespace

constexpr int kDefaultConvEpiRegs = 1;   // fused norm epilogue: 1 = single TMEM pass with the row in registers (stage D fused
                                         // conv 894 -> 1193 TF/s, profiles/r02_vae_conv_epilogue_ab.jsonl); MV_CONV_EPI=0 = two passes
int g_conv_epi = -1;
static int conv_epi_regs() {
  if (g_conv_epi < 0) {
    const char* e = getenv("MV_CONV_EPI");
    g_conv_epi = (e != nullptr && e[0] != 0) ? (atoi(e) != 0 ? 1 : 0) : kDefaultConvEpiRegs;
  }
  return g_conv_epi;
}

constexpr int kDefaultConvSplit = 1;
static int conv_split_enabled() {        // MV_CONV_SPLIT=0: A/B against one warp set per tile
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MV_CONV_SPLIT");
    v = (e != nullptr && e[0] != 0) ? (atoi(e) != 0 ? 1 : 0) : kDefaultConvSplit;
  }
  return v;
}

extern "C" int mv_vae_conv_config(int pair, int tiles_per_cta, int epi_regs) {
  if (pair >= 0) g_conv_pair = pair != 0 ? 1 : 0;
  else if (pair == -2) g_conv_pair = -1;   // back to MV_CONV_PAIR / the built-in default
  if (tiles_per_cta >= 0) g_conv_nt = tiles_per_cta;
  else if (tiles_per_cta == -2) g_conv_nt = -1;
  if (epi_regs >= 0) g_conv_epi = epi_regs != 0 ? 1 : 0;
  else if (epi_regs == -2) g_conv_epi = -1;
  return MV_OK;
}

static int vae_conv_impl(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed,
                         const float* bias, const void* res_cl, void* out, int out_mode, int out_T, int out_H,
                         int out_W, int Cout, int cout_real, int ntaps, const int8_t* taps_dt_dh_dw, int64_t o_base,
                         int64_t os_t, int64_t os_h, int64_t os_w, int nsplit, int64_t nsplit_off,
                         const float* norm_gamma, void* norm_out, int t_off, mv_stream_t stream, int stride_t = 1,
                         int stride_h = 1, int stride_w = 1) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(ntaps >= 1 && ntaps <= kConvMaxTaps, "mv_vae_conv: ntaps=%d out of range", ntaps);
  MV_REQUIRE(Cin % 16 == 0 && Cin >= 16, "mv_vae_conv: Cin=%d must be a multiple of 16", Cin);
  MV_REQUIRE(Cout % 16 == 0, "mv_vae_conv: (padded) Cout=%d must be a multiple of 16", Cout);
  MV_REQUIRE(out_T > 0 && out_H > 0 && out_W > 0, "mv_vae_conv: empty output grid");
  const bool strided = stride_t != 1 || stride_h != 1 || stride_w != 1;
  if (!strided) {
    MV_REQUIRE(t_off >= 0 && in_T == out_T + t_off && in_H == out_H && in_W == out_W,
               "mv_vae_conv: input grid %dx%dx%d does not match output %dx%dx%d + %d cached frames", in_T, in_H, in_W,
               out_T, out_H, out_W, t_off);
  } else {
    // strided: the last output voxel's first tap must lie inside the input; taps that run past the far edge read the
    // TMA zero fill (= ZeroPad2d((0,1,0,1)), vae.py:93-94)
    MV_REQUIRE(stride_t >= 1 && stride_t <= 2 && stride_h >= 1 && stride_h <= 2 && stride_w >= 1 && stride_w <= 2,
               "mv_vae_conv_strided: strides %d,%d,%d out of range (1 or 2)", stride_t, stride_h, stride_w);
    MV_REQUIRE(t_off >= 0 && in_T == stride_t * (out_T - 1) + t_off + 1,
               "mv_vae_conv_strided: %d input frames != %d * (%d - 1) + %d + 1", in_T, stride_t, out_T, t_off);
    MV_REQUIRE(stride_h * (out_H - 1) < in_H && stride_w * (out_W - 1) < in_W && stride_h * out_H + 1 >= in_H &&
                   stride_w * out_W + 1 >= in_W,
               "mv_vae_conv_strided: output %dx%d does not cover input %dx%d at stride %d,%d", out_H, out_W, in_H, in_W,
               stride_h, stride_w);
    MV_REQUIRE(norm_out == nullptr && out_mode == 0, "mv_vae_conv_strided: plain fp16 output only");
  }
  int BK = (Cin % 64 == 0) ? 64 : ((Cin % 32 == 0) ? 32 : 16);
  int BN;
  if (Cout <= 256) BN = Cout;
  else if (Cout % 192 == 0) BN = 192;
  else if (Cout % 256 == 0) BN = 256;
  else if (Cout % 128 == 0) BN = 128;
  else {
    set_error("mv_vae_conv: cannot tile Cout=%d", Cout);
    return MV_E_SHAPE;
  }
  MV_REQUIRE(nsplit == 0 || nsplit % BN == 0, "mv_vae_conv: nsplit must be a multiple of the N tile");
  MV_REQUIRE((norm_out == nullptr) == (norm_gamma == nullptr), "mv_vae_conv_fused: norm_gamma and norm_out go together");
  MV_REQUIRE(norm_out == nullptr || (BN == Cout && out_mode == 0 && nsplit == 0),
             "mv_vae_conv_fused: the fused norm needs the whole channel row in one N tile (Cout=%d <= 256)", Cout);
  MV_REQUIRE(out != nullptr || norm_out != nullptr, "mv_vae_conv: no output");

  // ---- CTA-pair kernel (conv_igemm_pair_kernel) for the tensor-bound convolutions -----------------------------------
  // (1x1x1 convs too: their tiles are all epilogue — 6 MMAs for 128 x 224 B of stores in the decoder's head — and the
  // pair kernel has two epilogue warp sets per CTA where the single-CTA kernel has one: ncu showed that one warp per
  // scheduler, not DRAM (2.6 TB/s), bounds them)
  if (conv_pair_enabled() && !strided && (BK == 64 || BK == 32) && BN % 16 == 0 && (ntaps >= 2 || Cin >= 64)) {
    Conv2Params q;
    memset(&q, 0, sizeof(q));
    bool ok = true;
    int dh_min = 127, dh_max = -127;
    for (int i = 0; i < ntaps; ++i) {
      const int dt = taps_dt_dh_dw[3 * i], dh = taps_dt_dh_dw[3 * i + 1], dw = taps_dt_dh_dw[3 * i + 2];
      dh_min = dh < dh_min ? dh : dh_min;
      dh_max = dh > dh_max ? dh : dh_max;
      int g = 0;
      while (g < q.ngroups && !(q.g_dt[g] == dt && q.g_dw[g] == dw)) ++g;
      if (g == q.ngroups) {
        if (q.ngroups == kConv2MaxGroups) { ok = false; break; }
        q.g_dt[g] = static_cast<int8_t>(dt);
        q.g_dw[g] = static_cast<int8_t>(dw);
        q.g_ntap[g] = 0;
        ++q.ngroups;
      }
      if (q.g_ntap[g] == 3) { ok = false; break; }
      q.g_tap[g][q.g_ntap[g]] = static_cast<int8_t>(i);
      q.g_dhoff[g][q.g_ntap[g]] = static_cast<int8_t>(dh);     // rebased below
      ++q.g_ntap[g];
    }
    if (ok && dh_max - dh_min <= 2) {
      q.regular = q.g_ntap[0];
      for (int g = 0; g < q.ngroups; ++g) {
        for (int i = 0; i < q.g_ntap[g]; ++i) {
          q.g_dhoff[g][i] = static_cast<int8_t>(q.g_dhoff[g][i] - dh_min);
          if (q.g_dhoff[g][i] != i) q.regular = 0;
        }
        if (q.g_ntap[g] != q.g_ntap[0]) q.regular = 0;
      }
      // tiles per CTA: measured (profiles/r02_vae_conv_shapes_pair.jsonl) — what pays is a DOUBLE-BUFFERED accumulator
      // (2 NT BN <= 512 columns) so that the epilogue of one super-tile overlaps the MMAs of the next: NT = 1 for the
      // 192-wide tiles (1559 vs 1062 TF/s at NT = 2 on stage C), NT = 2 for the 96-wide ones (797 vs 660 / 702 at 1 / 4)
      int nt = (BK == 64) ? 1 : 2;
      if (g_conv_nt < 0) {             // MV_CONV_NT=1|2|4 / mv_vae_conv_config: tiles per CTA (A/B measurements)
        const char* e = getenv("MV_CONV_NT");
        g_conv_nt = e ? atoi(e) : 0;
      }
      if (g_conv_nt == 1 || g_conv_nt == 2 || g_conv_nt == 4) nt = g_conv_nt;
      while (nt > 1 && nt * BN > 512) nt >>= 1;
      if (BK == 64 && nt == 4) nt = 2;                 // instantiated: (64, 1|2), (32, 1|2|4)
      const uint32_t rowb = static_cast<uint32_t>(BK) * 2;
      q.dh_min = dh_min;
      q.box_h = 8 * nt + (dh_max - dh_min);
      q.a_box_bytes = (static_cast<uint32_t>(q.box_h) * kConvTW * rowb + 1023u) & ~1023u;
      q.b_slab_bytes = (static_cast<uint32_t>(BN / 2) * rowb + 1023u) & ~1023u;
      q.stage_bytes = q.a_box_bytes + 3 * q.b_slab_bytes;
      q.stages = static_cast<int>(kConv2RingBytes / q.stage_bytes);
      if (q.stages > kConv2MaxStages) q.stages = kConv2MaxStages;
      q.nbuf = (2 * nt * BN <= 512) ? 2 : 1;
      // one tile per CTA leaves the second epilogue warp set idle: let both drain the tile, half of the columns each
      // (plain epilogue: any BN that halves into 32-column chunks; fused norm: 192 = 2 x 96 register-resident halves)
      q.split_cols = (nt == 1 && BN >= 64 && out_mode == 0 && conv_split_enabled() &&
                      (norm_out == nullptr ? true : (conv_epi_regs() && BN == 192))) ? 1 : 0;
      q.sup_h = (out_H + 2 * nt * kConvTH - 1) / (2 * nt * kConvTH);
      q.sup_w = (out_W + kConvTW - 1) / kConvTW;
      const int num_n = Cout / BN;
      const int64_t ns = static_cast<int64_t>(num_n) * q.sup_h * q.sup_w * out_T;
      if (q.stages >= 2 && ns < (1ll << 31) && q.box_h <= 256) {
        q.num_super = static_cast<int>(ns);
        ConvParams& p = q.c;
        p.bias = bias;
        p.res = reinterpret_cast<const __half*>(res_cl);
        p.out = out;
        p.o_base = o_base;
        p.os_t = os_t;
        p.os_h = os_h;
        p.os_w = os_w;
        p.nsplit = nsplit;
        p.nsplit_off = nsplit_off;
        p.T = out_T;
        p.t_off = t_off;
        p.st = p.sh = p.sw = 1;
        p.H = out_H;
        p.W = out_W;
        p.Cin = Cin;
        p.Cout = Cout;
        p.BN = BN;
        p.ntaps = ntaps;
        p.num_n = num_n;
        p.kblocks_per_tap = Cin / BK;
        p.out_mode = out_mode;
        p.cout_real = cout_real;
        p.norm_gamma = norm_gamma;
        p.norm_out = reinterpret_cast<__half*>(norm_out);
        p.epi_regs = conv_epi_regs();
        CUtensorMap tmA2, tmB2;
        {
          uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)in_W, (uint64_t)in_H, (uint64_t)in_T};
          uint64_t str[4] = {2, (uint64_t)Cin * 2, (uint64_t)Cin * in_W * 2, (uint64_t)Cin * in_W * in_H * 2};
          uint32_t box[4] = {(uint32_t)BK, kConvTW, (uint32_t)q.box_h, 1};
          rc = make_tmap_bf16_sw(&tmA2, in_cl, 4, dims, str, box, BK * 2);
          if (rc != MV_OK) return rc;
        }
        {
          uint64_t dims[2] = {(uint64_t)ntaps * Cin, (uint64_t)Cout};
          uint64_t str[2] = {2, (uint64_t)ntaps * Cin * 2};
          uint32_t box[2] = {(uint32_t)BK, (uint32_t)(BN / 2)};
          rc = make_tmap_bf16_sw(&tmB2, w_packed, 2, dims, str, box, BK * 2);
          if (rc != MV_OK) return rc;
        }
        cudaStream_t st2 = static_cast<cudaStream_t>(stream);
        if (BK == 64) return nt == 2 ? launch_conv_pair<64, 2>(tmA2, tmB2, q, st2) : launch_conv_pair<64, 1>(tmA2, tmB2, q, st2);
        if (nt == 4) return launch_conv_pair<32, 4>(tmA2, tmB2, q, st2);
        return nt == 2 ? launch_conv_pair<32, 2>(tmA2, tmB2, q, st2) : launch_conv_pair<32, 1>(tmA2, tmB2, q, st2);
      }
    }
  }

  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)in_W, (uint64_t)in_H, (uint64_t)in_T};
    uint64_t str[4] = {2, (uint64_t)Cin * 2, (uint64_t)Cin * in_W * 2, (uint64_t)Cin * in_W * in_H * 2};
    uint32_t box[4] = {(uint32_t)BK, (uint32_t)(kConvTW * stride_w), (uint32_t)(kConvTH * stride_h), 1};
    uint32_t es[4] = {1, (uint32_t)stride_w, (uint32_t)stride_h, 1};
    // (a 16-bit tensor map only moves bytes: the bf16 encoder serves fp16 data unchanged)
    rc = make_tmap_bf16_es(&tmA, in_cl, 4, dims, str, box, BK * 2, es);
    if (rc != MV_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)ntaps * Cin, (uint64_t)Cout};
    uint64_t str[2] = {2, (uint64_t)ntaps * Cin * 2};
    uint32_t box[2] = {(uint32_t)BK, (uint32_t)BN};
    rc = make_tmap_bf16_sw(&tmB, w_packed, 2, dims, str, box, BK * 2);
    if (rc != MV_OK) return rc;
  }
  ConvParams p;
  p.bias = bias;
  p.res = reinterpret_cast<const __half*>(res_cl);
  p.out = out;
  p.o_base = o_base;
  p.os_t = os_t;
  p.os_h = os_h;
  p.os_w = os_w;
  p.nsplit = nsplit;
  p.nsplit_off = nsplit_off;
  p.T = out_T;
  p.t_off = t_off;
  p.st = stride_t;
  p.sh = stride_h;
  p.sw = stride_w;
  p.H = out_H;
  p.W = out_W;
  p.Cin = Cin;
  p.Cout = Cout;
  p.BN = BN;
  p.ntaps = ntaps;
  for (int i = 0; i < ntaps; ++i) {
    p.dt[i] = taps_dt_dh_dw[3 * i];
    p.dh[i] = taps_dt_dh_dw[3 * i + 1];
    p.dw[i] = taps_dt_dh_dw[3 * i + 2];
  }
  p.num_n = Cout / BN;
  p.tiles_h = (out_H + kConvTH - 1) / kConvTH;
  p.tiles_w = (out_W + kConvTW - 1) / kConvTW;
  const int64_t nt = static_cast<int64_t>(p.num_n) * p.tiles_h * p.tiles_w * out_T;
  MV_REQUIRE(nt < (1ll << 31), "mv_vae_conv: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.kblocks_per_tap = Cin / BK;
  p.out_mode = out_mode;
  p.cout_real = cout_real;
  p.norm_gamma = norm_gamma;
  p.norm_out = reinterpret_cast<__half*>(norm_out);
  p.epi_regs = conv_epi_regs();
  p.kps = (BK == 32 && p.kblocks_per_tap % 3 == 0) ? 3 : 1;
  p.a_block_bytes = (static_cast<uint32_t>(kConvBM * BK * 2) + 1023u) & ~1023u;
  p.b_block_bytes = (static_cast<uint32_t>(BN * BK * 2) + 1023u) & ~1023u;
  p.a_stage_bytes = p.a_block_bytes * p.kps;
  p.b_stage_bytes = p.b_block_bytes * p.kps;
  p.stages = static_cast<int>(kConvRingBytes / (p.a_stage_bytes + p.b_stage_bytes));
  if (p.stages > kConvMaxStages) p.stages = kConvMaxStages;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (BK == 64) return launch_conv<64>(tmA, tmB, p, st);
  if (BK == 32) return launch_conv<32>(tmA, tmB, p, st);
  return launch_conv<16>(tmA, tmB, p, st);
}

extern "C" int mv_vae_conv(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed,
                           const float* bias, const void* res_cl, void* out, int out_mode, int out_T, int out_H,
                           int out_W, int Cout, int cout_real, int ntaps, const int8_t* taps_dt_dh_dw, int64_t o_base,
                           int64_t os_t, int64_t os_h, int64_t os_w, int nsplit, int64_t nsplit_off, int t_off,
                           mv_stream_t stream) {
  return vae_conv_impl(in_cl, in_T, in_H, in_W, Cin, w_packed, bias, res_cl, out, out_mode, out_T, out_H, out_W, Cout,
                       cout_real, ntaps, taps_dt_dh_dw, o_base, os_t, os_h, os_w, nsplit, nsplit_off, nullptr, nullptr,
                       t_off, stream);
}

extern "C" int mv_vae_conv_fused(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed,
                                 const float* bias, const void* res_cl, void* out, int out_T, int out_H, int out_W,
                                 int Cout, int ntaps, const int8_t* taps_dt_dh_dw, int64_t o_base, int64_t os_t,
                                 int64_t os_h, int64_t os_w, const float* norm_gamma, void* norm_out, int t_off,
                                 mv_stream_t stream) {
  return vae_conv_impl(in_cl, in_T, in_H, in_W, Cin, w_packed, bias, res_cl, out, 0, out_T, out_H, out_W, Cout, Cout,
                       ntaps, taps_dt_dh_dw, o_base, os_t, os_h, os_w, 0, 0, norm_gamma, norm_out, t_off, stream);
}

extern "C" int mv_vae_conv_strided(const void* in_cl, int in_T, int in_H, int in_W, int Cin, const void* w_packed,
                                   const float* bias, void* out_cl, int out_T, int out_H, int out_W, int Cout, int ntaps,
                                   const int8_t* taps_dt_dh_dw, int t_off, int stride_t, int stride_h, int stride_w,
                                   mv_stream_t stream) {
  return vae_conv_impl(in_cl, in_T, in_H, in_W, Cin, w_packed, bias, nullptr, out_cl, 0, out_T, out_H, out_W, Cout, Cout,
                       ntaps, taps_dt_dh_dw, 0, static_cast<int64_t>(out_H) * out_W * Cout,
                       static_cast<int64_t>(out_W) * Cout, Cout, 0, 0, nullptr, nullptr, t_off, stream, stride_t, stride_h,
                       stride_w);
}

extern "C" int mv_vae_video_in(const float* video, int T_total, int t0, int n, int H, int W, void* out_cl,
                               mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(n > 0 && t0 >= 0 && t0 + n <= T_total && H > 0 && W > 0, "mv_vae_video_in: bad frame range / shape");
  const int64_t plane = static_cast<int64_t>(H) * W;
  const int64_t nvox = plane * n;
  int64_t blocks = (nvox + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  vae_video_in_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      video, static_cast<int64_t>(T_total) * plane, static_cast<int64_t>(t0) * plane, nvox,
      reinterpret_cast<uint4*>(out_cl));
  MV_CHECK_LAUNCH("vae_video_in_kernel");
  return MV_OK;
}

extern "C" int mv_vae_latent_out(const void* head_cl, const float* W1, const float* b1, const float* mean,
                                 const float* inv_std, float* mu, int Z, int64_t nvox, int64_t mu_plane, int64_t mu_off,
                                 mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(Z > 0 && Z <= 16 && nvox > 0 && mu_off >= 0 && mu_off + nvox <= mu_plane, "mv_vae_latent_out: bad shape");
  int64_t blocks = (nvox + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  vae_latent_out_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(head_cl), W1, b1, mean, inv_std, mu, Z, nvox, mu_plane, mu_off);
  MV_CHECK_LAUNCH("vae_latent_out_kernel");
  return MV_OK;
}

extern "C" int mv_vae_head_gather(const void* d_cur, const void* d_prev, int kprev, int n, int H, int W,
                                  const float* bias3_host, float* video, int64_t plane, int64_t frame_off,
                                  mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0 && kprev >= 0 && kprev <= 2 && (kprev == 0 || d_prev != nullptr) &&
                 frame_off >= 0 && frame_off + static_cast<int64_t>(n) * H * W <= plane,
             "mv_vae_head_gather: bad shape (n=%d H=%d W=%d kprev=%d)", n, H, W, kprev);
  dim3 grid((W + 31) / 32, (H + 7) / 8, n);
  vae_head_gather_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __half*>(d_cur), reinterpret_cast<const __half*>(d_prev), kprev, H, W, bias3_host[0],
      bias3_host[1], bias3_host[2], video, plane, frame_off);
  MV_CHECK_LAUNCH("vae_head_gather_kernel");
  return MV_OK;
}

extern "C" int mv_vae_rmsnorm_silu(const void* x_cl, void* y_cl, const float* gamma, int64_t nvox, int C, int silu,
                                   mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(nvox > 0 && C > 0 && C % 8 == 0 && C <= 512, "mv_vae_rmsnorm_silu: bad shape nvox=%lld C=%d", (long long)nvox, C);
  const int G = C <= 128 ? 16 : 32;
  const int64_t per_warp = (32 / G) * 2;
  const int64_t warps_needed = (nvox + per_warp - 1) / per_warp;
  int64_t blocks = (warps_needed + 7) / 8;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* x = reinterpret_cast<const __half*>(x_cl);
  __half* y = reinterpret_cast<__half*>(y_cl);
  if (C <= 128) rmsnorm_silu_cl_kernel<16, 1><<<static_cast<int>(blocks), 256, 0, st>>>(x, y, gamma, nvox, C, silu);
  else if (C <= 256) rmsnorm_silu_cl_kernel<32, 1><<<static_cast<int>(blocks), 256, 0, st>>>(x, y, gamma, nvox, C, silu);
  else rmsnorm_silu_cl_kernel<32, 2><<<static_cast<int>(blocks), 256, 0, st>>>(x, y, gamma, nvox, C, silu);
  MV_CHECK_LAUNCH("rmsnorm_silu_cl_kernel");
  return MV_OK;
}

extern "C" int mv_vae_latent_in(const float* z, const float* W2, const float* b2, const float* mean, const float* stdv,
                                void* out_cl, int Z, int64_t nvox, mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(Z > 0 && Z <= 32 && nvox > 0, "mv_vae_latent_in: bad shape");
  int64_t blocks = (nvox + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  vae_latent_in_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      z, W2, b2, mean, stdv, reinterpret_cast<__half*>(out_cl), Z, nvox);
  MV_CHECK_LAUNCH("vae_latent_in_kernel");
  return MV_OK;
}

extern "C" int mv_softmax_rows(const float* S, int64_t lds, void* P_f16, int64_t ldp, int M, int N, float scale,
                               mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(M > 0 && N > 0, "mv_softmax_rows: empty problem");
  if ((N & 3) == 0 && (lds & 3) == 0 && (ldp & 3) == 0 && N <= 1024 * 4 * kSmxVec &&
      (reinterpret_cast<uintptr_t>(S) & 15) == 0 && (reinterpret_cast<uintptr_t>(P_f16) & 7) == 0) {
    softmax_rows_reg_kernel<<<M, 1024, 0, static_cast<cudaStream_t>(stream)>>>(S, lds, reinterpret_cast<__half*>(P_f16),
                                                                              ldp, N, scale);
    MV_CHECK_LAUNCH("softmax_rows_reg_kernel");
    return MV_OK;
  }
  softmax_rows_kernel<<<M, 256, 0, static_cast<cudaStream_t>(stream)>>>(S, lds, reinterpret_cast<__half*>(P_f16),
                                                                       ldp, N, scale);
  MV_CHECK_LAUNCH("softmax_rows_kernel");
  return MV_OK;
}
