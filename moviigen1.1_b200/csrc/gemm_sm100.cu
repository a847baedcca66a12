// tcgen05 GEMM for the DiT linears: out = epilogue(A[M,K] . W[N,K]^T + bias).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   (A tile 128x64, W tile BNx64 per stage, 128B-swizzled; BN = 256 with a 4-stage ring,
//                               or BN = 64 with an 8-stage ring for skinny problems)
//   warp 1      MMA issuer     (one elected thread, tcgen05.mma cta_group::1 M=128 N=BN K=16,
//                               fp32 accumulators in TMEM, 2 accumulator stages)
//   warps 2..5  epilogue       (tcgen05.ld -> bias / GELU / gate+residual -> global), overlapped
//                               with the next tile's main loop through the 2 TMEM stages
// Tile order is grouped (16 M-tiles x all N-tiles) so that a wave's working set stays in L2.
//
// Replaces nn.Linear under bf16 autocast (cuBLASLt) — wan/modules/model.py:139-141,155,171-173,180,
// 267-269,451-453 — and the patch-embedding Conv3d (:445-450,529).
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGroupM = 16;
constexpr int kDefaultGemmPair = 0;   // flipped to 1 once measured faster in the step (profiles/)
constexpr int kGemmThreads = 192;
constexpr uint32_t kABytes = BM * BK * 2;  // 16 KB
constexpr uint32_t kEpiStageBytes = 4 * 4096;  // one 32x32 fp32 transpose buffer per epilogue warp

// Two tile widths: N = 256 (4-stage ring) for the DiT linears, and N = 64 (8-stage ring) for single-M-tile problems
// (M <= 128: umT5 encoder on a typical prompt) where a 256-wide tiling would leave most SMs without a tile — those
// are weight-streaming bound, so what matters is how many SMs pull weights and how many bytes are in flight
// (umT5 encode of 128 tokens: 7.9 -> 5.8 ms).
template <int BN>
struct GemmCfg {
  static constexpr int kStages = BN == 256 ? 4 : 8;
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kSmem = kStages * kStageBytes + kEpiStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
  const float* bias;
  const float* gate;
  void* out;
  int64_t ldo;
  int M, N, K;
  int num_m, num_n, num_tiles, num_kb;
  int group_m;   // M tiles per raster group (tile order: all N tiles of a group of group_m M tiles, M fastest)
  int a_kblock;  // > 0: A is split along K into blocks of a_kblock columns (3-D tensor map {k, m, block})
  int f16;         // 1: operands (and 16-bit outputs) are IEEE fp16 instead of bf16 (mv_gemm_f16: WanVAE attention)
  int stream_out;  // 1 (MV_GEMM_STREAM=1): ld/st.global.cs (evict-first) for the output and the fp32 residual, so that
                   // this one-touch traffic does not push the re-used A / W tiles out of L2 (ncu: 4.4 GB DRAM reads
                   // for 2.4 GB algorithmic on the 75600 x 5120 x 5120 residual GEMM).  Measured neutral: off.
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_blk, int& n_blk) {
  const int per_group = p.group_m * p.num_n;
  const int g = tile / per_group;
  const int within = tile - g * per_group;
  const int gm = min(p.group_m, p.num_m - g * p.group_m);
  m_blk = g * p.group_m + within % gm;
  n_blk = within / gm;
}

// Drains one accumulator tile (this warp's 32 TMEM lanes x tile_cols fp32 columns at taddr) through the epilogue:
// TMEM -> registers (one accumulator row per thread) -> per-warp smem transpose -> global, so that every global access
// of a warp covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes.  row_base = first of the warp's 32
// output rows, col_base = first output column of the tile; stg = this warp's 4 KB staging buffer.
template <int EPI>
__device__ __forceinline__ void epi_drain_tile(const GemmParams& p, uint32_t taddr, int row_base, int col_base,
                                               int tile_cols, uint8_t* stg, int lane) {
  const int rr0 = lane >> 3;   // row inside a 4-row group
  const int cc = lane & 7;     // 16-byte column chunk = 4 fp32 accumulator columns
#pragma unroll 1
  for (int c = 0; c < tile_cols / 32; ++c) {
    const int col0 = col_base + c * 32;
    if (col0 >= p.N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld_x32(taddr + c * 32, r);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
          make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    __syncwarp();
    const int col = col0 + cc * 4;
    const bool col_full = (col + 4 <= p.N);
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    float g4[4] = {1.f, 1.f, 1.f, 1.f};
    if (col_full) {
      if (p.bias != nullptr) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
      }
      if (EPI == MV_EPI_RESID_F32 && p.gate != nullptr) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.gate + col));
        g4[0] = t.x; g4[1] = t.y; g4[2] = t.z; g4[3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (col + i < p.N) {
          if (p.bias != nullptr) b4[i] = __ldg(p.bias + col + i);
          if (EPI == MV_EPI_RESID_F32 && p.gate != nullptr) g4[i] = __ldg(p.gate + col + i);
        }
    }
    // residual epilogue: issue all eight row loads first so their latencies overlap (a load placed after the
    // previous row's store could not be hoisted by the compiler: same pointer, possible aliasing)
    float4 xres[8];
    if constexpr (EPI == MV_EPI_RESID_F32) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int row = row_base + rr0 + 4 * k;
        xres[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < p.M && col_full) {
          const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) +
                                                              static_cast<int64_t>(row) * p.ldo + col);
          xres[k] = p.stream_out ? __ldcs(src) : *src;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int rr = rr0 + 4 * k;
      const int row = row_base + rr;
      const uint4 a = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((cc ^ (rr & 7)) << 4));
      if (row < p.M && col < p.N) {
        float v[4] = {__uint_as_float(a.x) + b4[0], __uint_as_float(a.y) + b4[1], __uint_as_float(a.z) + b4[2],
                      __uint_as_float(a.w) + b4[3]};
        if constexpr (EPI != MV_EPI_F32) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = p.f16 ? f16_round(v[i]) : bf16_round(v[i]);
        }
        if constexpr (EPI == MV_EPI_BF16_GELU) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = gelu_tanh(v[i]);
        }
        if constexpr (EPI == MV_EPI_BF16 || EPI == MV_EPI_BF16_GELU) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<int64_t>(row) * p.ldo + col;
          if (col_full) {
            const uint2 w2 = p.f16 ? make_uint2(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]))
                                   : make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            if (p.stream_out) __stcs(reinterpret_cast<uint2*>(o), w2);
            else *reinterpret_cast<uint2*>(o) = w2;
          } else {
            for (int i = 0; i < 4 && col + i < p.N; ++i) {
              if (p.f16) reinterpret_cast<__half*>(o)[i] = __float2half_rn(f16_sat(v[i]));
              else o[i] = __float2bfloat16_rn(v[i]);
            }
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + static_cast<int64_t>(row) * p.ldo + col;
          if (col_full) {
            float4 w;
            if constexpr (EPI == MV_EPI_RESID_F32) {
              const float4 x = xres[k];
              w.x = x.x + v[0] * g4[0]; w.y = x.y + v[1] * g4[1]; w.z = x.z + v[2] * g4[2]; w.w = x.w + v[3] * g4[3];
            } else {
              w.x = v[0]; w.y = v[1]; w.z = v[2]; w.w = v[3];
            }
            if (p.stream_out) __stcs(reinterpret_cast<float4*>(o), w);
            else *reinterpret_cast<float4*>(o) = w;
          } else {
            for (int i = 0; i < 4 && col + i < p.N; ++i) {
              if constexpr (EPI == MV_EPI_RESID_F32) o[i] = o[i] + v[i] * g4[i];
              else o[i] = v[i];
            }
          }
        }
      }
    }
    __syncwarp();  // staging buffer is reused by the next chunk
  }
}

template <int EPI, int BN, bool HI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  constexpr int kStages = GemmCfg<BN>::kStages;
  constexpr uint32_t kBBytes = GemmCfg<BN>::kBBytes;
  constexpr uint32_t kStageBytes = GemmCfg<BN>::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  uint8_t* sEpi = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kEpiStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = bars + 2 * kStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int pw = threadIdx.x >> 5;                      // physical warp: TMEM lane quadrant = pw & 3
  const int warp = HI ? (pw + 2) % 6 : pw;              // role id: 0 TMA, 1 MMA, 2-5 epilogue (HI: roles on warps 4, 5)
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // (warp-uniform loop, elect.sync only around the TMA issue: keeps addresses in uniform registers)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full[stage], kStageBytes);
          if (p.a_kblock > 0) {
            const int k0 = kb * BK;
            const int blk = k0 / p.a_kblock;
            tma_load_3d(sA + stage * kABytes, &tmA, &full[stage], k0 - blk * p.a_kblock, m_blk * BM, blk);
          } else {
            tma_load_2d(sA + stage * kABytes, &tmA, &full[stage], kb * BK, m_blk * BM);
          }
          tma_load_2d(sB + stage * kBBytes, &tmB, &full[stage], kb * BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    const uint32_t idesc = p.f16 ? make_idesc_f16(BM, BN, 0, 0) : make_idesc_bf16(BM, BN, 0, 0);
    const uint64_t adesc0 = make_desc_kmajor_sw128(smem_u32(sA));
    const uint64_t bdesc0 = make_desc_kmajor_sw128(smem_u32(sB));
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = adesc0 + ((stage * kABytes) >> 4);
          const uint64_t bdesc = bdesc0 + ((stage * kBBytes) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes (>>4 = 2) per K=16 step inside the 128B swizzle atom
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == p.num_kb - 1) umma_commit(&tfull[as]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  } else {
    // ------------------------------ epilogue ----------------------------------
    // TMEM -> registers (one accumulator row per thread) -> per-warp smem transpose -> global, so that every
    // global access of a warp covers 4 rows x 128 contiguous bytes instead of 32 rows x 16 bytes.
    const int quad = pw & 3;  // TMEM lane quadrant this warp may access
    uint8_t* stg = sEpi + quad * 4096;  // 32 rows x 128 B, 16-byte chunks XOR-swizzled by (row & 7)
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const int row_base = m_blk * BM + quad * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      epi_drain_tile<EPI>(p, taddr, row_base, n_blk * BN, BN, stg, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
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
// CTA-pair variant (cta_group::2): a cluster of two CTAs (the two SMs of a TPC) computes one 256 x 256 output tile.
// CTA r of the pair owns output rows [m0 + 128 r, +128): it stages its own A rows and the W rows [n0 + 128 r, +128)
// (HALF of the B tile) per stage; the leader's single MMA thread issues tcgen05.mma.cta_group::2 (M 256, N 256, K 16)
// which reads A from both CTAs and each half of B from the CTA that loaded it, and accumulates into each CTA's own
// TMEM.  Per CTA and K block: 16 KB + 16 KB instead of 16 KB + 32 KB of TMA traffic and shared-memory operand reads
// (64 instead of 96 B/clk), 6 pipeline stages instead of 4 in the same shared memory.
//   both CTAs   warp 0  TMA producer: waits on its own `empty`, loads into its own smem, transaction bytes are
//                       reported to the LEADER's `full` barrier
//   leader      warp 1  MMA issuer; tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to the
//                       same barrier in both CTAs
//   both CTAs   warps 2-5 epilogue over their own 128 accumulator rows; "accumulator drained" arrives at the LEADER's
//                       `tempty` (8 arrivals: 4 warps x 2 CTAs)
// ================================================================================================================
constexpr int kPairStages = 6;
constexpr int kPairBN = 256;                         // N of the pair's tile; each CTA stages kPairBN / 2 rows of W
constexpr uint32_t kPairBBytes = (kPairBN / 2) * BK * 2;   // 16 KB
constexpr uint32_t kPairStageBytes = kABytes + kPairBBytes;
constexpr uint32_t kPairSmem = kPairStages * kPairStageBytes + kEpiStageBytes + 1024 + 256;

template <int EPI, bool HI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kPairStages * kABytes;
  uint8_t* sEpi = smem + kPairStages * kPairStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes + kEpiStageBytes);
  uint64_t* full = bars;                        // used in the leader only
  uint64_t* empty = bars + kPairStages;         // one set per CTA (multicast commit)
  uint64_t* tfull = bars + 2 * kPairStages;     // one set per CTA (multicast commit)
  uint64_t* tempty = bars + 2 * kPairStages + 2;  // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

  const int pw = threadIdx.x >> 5;                      // physical warp: TMEM lane quadrant = pw & 3
  const int warp = HI ? (pw + 2) % 6 : pw;              // role id: 0 TMA, 1 MMA, 2-5 epilogue (HI: roles on warps 4, 5)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kPairStages; ++i) {
      mbar_init(&full[i], 1);     // the leader producer's arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(&empty[i], 1);    // one multicast commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);   // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();             // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      const int row0 = m_blk * (2 * BM) + static_cast<int>(rank) * BM;
      const int wrow0 = n_blk * kPairBN + static_cast<int>(rank) * (kPairBN / 2);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[stage], 2 * kPairStageBytes);
          if (p.a_kblock > 0) {
            const int k0 = kb * BK;
            const int blk = k0 / p.a_kblock;
            tma_load_3d_pair(sA + stage * kABytes, &tmA, &full[stage], k0 - blk * p.a_kblock, row0, blk);
          } else {
            tma_load_2d_pair(sA + stage * kABytes, &tmA, &full[stage], kb * BK, row0);
          }
          tma_load_2d_pair(sB + stage * kPairBBytes, &tmB, &full[stage], kb * BK, wrow0);
        }
        __syncwarp();
        if (++stage == kPairStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) --------------------------------
    if (leader) {
      const uint32_t idesc = p.f16 ? make_idesc_f16(2 * BM, kPairBN, 0, 0) : make_idesc_bf16(2 * BM, kPairBN, 0, 0);
      const uint64_t adesc0 = make_desc_kmajor_sw128(smem_u32(sA));
      const uint64_t bdesc0 = make_desc_kmajor_sw128(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * kPairBN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = adesc0 + ((stage * kABytes) >> 4);
            const uint64_t bdesc = bdesc0 + ((stage * kPairBBytes) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_pair(&empty[stage], 3);                       // stage free in both CTAs
            if (kb == p.num_kb - 1) umma_commit_pair(&tfull[as], 3);  // accumulator ready in both CTAs
          }
          __syncwarp();
          if (++stage == kPairStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ----------------------------------
    const int quad = pw & 3;
    uint8_t* stg = sEpi + quad * 4096;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair_id; tile < p.num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const int row_base = m_blk * (2 * BM) + static_cast<int>(rank) * BM + quad * 32;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * kPairBN;
      epi_drain_tile<EPI>(p, taddr, row_base, n_blk * kPairBN, kPairBN, stg, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty[as], 0);   // the leader's barrier counts both CTAs' epilogue warps
      as ^= 1;
      if (as == 0) aphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();             // nobody frees TMEM / exits while the peer may still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int EPI>
static int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_pair_kernel<EPI, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kPairSmem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_pair_kernel<EPI, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kPairSmem)));
    attr_set = true;
  }
  int pairs = sm_count() / 2;
  if (p.num_tiles < pairs) pairs = p.num_tiles;
  if (roles_hi()) gemm_bf16_pair_kernel<EPI, true><<<2 * pairs, kGemmThreads, kPairSmem, st>>>(tmA, tmB, p);
  else gemm_bf16_pair_kernel<EPI, false><<<2 * pairs, kGemmThreads, kPairSmem, st>>>(tmA, tmB, p);
  MV_CHECK_LAUNCH("gemm_bf16_pair_kernel");
  return MV_OK;
}

static int dispatch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int epilogue,
                              cudaStream_t st) {
  switch (epilogue) {
    case MV_EPI_BF16: return launch_gemm_pair<MV_EPI_BF16>(tmA, tmB, p, st);
    case MV_EPI_BF16_GELU: return launch_gemm_pair<MV_EPI_BF16_GELU>(tmA, tmB, p, st);
    case MV_EPI_RESID_F32: return launch_gemm_pair<MV_EPI_RESID_F32>(tmA, tmB, p, st);
    case MV_EPI_F32: return launch_gemm_pair<MV_EPI_F32>(tmA, tmB, p, st);
    default: return launch_gemm_pair<MV_EPI_F32_ROUND>(tmA, tmB, p, st);
  }
}

template <int EPI, int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MV_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, BN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(GemmCfg<BN>::kSmem)));
    MV_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, BN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(GemmCfg<BN>::kSmem)));
    attr_set = true;
  }
  const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  if (roles_hi()) gemm_bf16_kernel<EPI, BN, true><<<grid, kGemmThreads, GemmCfg<BN>::kSmem, st>>>(tmA, tmB, p);
  else gemm_bf16_kernel<EPI, BN, false><<<grid, kGemmThreads, GemmCfg<BN>::kSmem, st>>>(tmA, tmB, p);
  MV_CHECK_LAUNCH("gemm_bf16_kernel");
  return MV_OK;
}

template <int BN>
static int dispatch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int epilogue,
                         cudaStream_t st) {
  switch (epilogue) {
    case MV_EPI_BF16: return launch_gemm<MV_EPI_BF16, BN>(tmA, tmB, p, st);
    case MV_EPI_BF16_GELU: return launch_gemm<MV_EPI_BF16_GELU, BN>(tmA, tmB, p, st);
    case MV_EPI_RESID_F32: return launch_gemm<MV_EPI_RESID_F32, BN>(tmA, tmB, p, st);
    case MV_EPI_F32: return launch_gemm<MV_EPI_F32, BN>(tmA, tmB, p, st);
    default: return launch_gemm<MV_EPI_F32_ROUND, BN>(tmA, tmB, p, st);
  }
}

}  // namespace mv

namespace {
int g_gemm_pair = -1;   // -1: read MV_GEMM_PAIR on first use
bool gemm_pair_enabled() {
  if (g_gemm_pair < 0) {
    const char* e = getenv("MV_GEMM_PAIR");
    g_gemm_pair = (e != nullptr && e[0] != 0) ? (atoi(e) != 0 ? 1 : 0) : mv::kDefaultGemmPair;
  }
  return g_gemm_pair == 1;
}
}  // namespace

extern "C" int mv_gemm_config(int pair) {
  if (pair >= 0) g_gemm_pair = pair != 0 ? 1 : 0;
  else if (pair == -2) g_gemm_pair = -1;   // back to MV_GEMM_PAIR / the built-in default
  return MV_OK;
}

static int gemm_impl(const void* A, int64_t lda, int64_t a_block_stride, int a_kblock, const void* W, int64_t ldw,
                     const float* bias, void* out, int64_t ldo, const float* gate, int M, int N, int K, int epilogue,
                     mv_stream_t stream, int f16 = 0) {
  using namespace mv;
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(M > 0 && N > 0 && K > 0, "mv_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  MV_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "mv_gemm_bf16: K/lda/ldw must be multiples of 8 (K=%d lda=%lld ldw=%lld)",
             K, (long long)lda, (long long)ldw);
  MV_REQUIRE(lda >= (a_kblock > 0 ? a_kblock : K) && ldw >= K && ldo >= N, "mv_gemm_bf16: leading dimensions too small");
  MV_REQUIRE(epilogue >= 0 && epilogue <= 4, "mv_gemm_bf16: unknown epilogue %d", epilogue);
  const bool f32_out = (epilogue == MV_EPI_RESID_F32 || epilogue == MV_EPI_F32_ROUND || epilogue == MV_EPI_F32);
  MV_REQUIRE(ldo % (f32_out ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
             "mv_gemm_bf16: output must be 16-byte aligned with 16-byte aligned rows");
  MV_REQUIRE(gate == nullptr || (reinterpret_cast<uintptr_t>(gate) & 15) == 0, "mv_gemm_bf16: gate must be 16B aligned");
  MV_REQUIRE(a_kblock == 0 || (a_kblock % BK == 0 && K % a_kblock == 0 && a_block_stride % 8 == 0),
             "mv_gemm_bf16_ksplit: k block %d must be a multiple of %d dividing K=%d", a_kblock, BK, K);

  CUtensorMap tmA, tmB;
  if (a_kblock == 0) {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t str[2] = {2, static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {BK, BM};
    rc = make_tmap_bf16(&tmA, A, 2, dims, str, box, true);
    if (rc != MV_OK) return rc;
  } else {
    uint64_t dims[3] = {static_cast<uint64_t>(a_kblock), static_cast<uint64_t>(M), static_cast<uint64_t>(K / a_kblock)};
    uint64_t str[3] = {2, static_cast<uint64_t>(lda) * 2, static_cast<uint64_t>(a_block_stride) * 2};
    uint32_t box[3] = {BK, BM, 1};
    rc = make_tmap_bf16(&tmA, A, 3, dims, str, box, true);
    if (rc != MV_OK) return rc;
  }
  // CTA-pair kernel (256 x 256 tiles) for the big DiT linears; the single-CTA kernels for everything else.
  // MV_GEMM_PAIR=0|1 / mv_gemm_config() override (A/B measurements, tests).
  const bool use_pair = gemm_pair_enabled() && M >= 1024 && N >= 512;
  if (use_pair) {
    CUtensorMap tmBp;
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[2] = {2, static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(kPairBN / 2)};
    rc = make_tmap_bf16(&tmBp, W, 2, dims, str, box, true);
    if (rc != MV_OK) return rc;
    GemmParams p;
    p.bias = bias;
    p.gate = gate;
    p.out = out;
    p.ldo = ldo;
    p.M = M;
    p.N = N;
    p.K = K;
    p.num_m = (M + 2 * BM - 1) / (2 * BM);
    p.num_n = (N + kPairBN - 1) / kPairBN;
    p.num_tiles = p.num_m * p.num_n;
    p.num_kb = (K + BK - 1) / BK;
    p.a_kblock = a_kblock;
    p.f16 = f16;
    p.stream_out = 0;
    {
      static int grp = -1;     // MV_GEMM_PAIR_GROUP: 256-row M tiles per raster group (default 8 = 2048 rows, the same
      if (grp < 0) {           // A working set as the single-CTA kernel's 16 x 128)
        const char* e = getenv("MV_GEMM_PAIR_GROUP");
        grp = (e != nullptr && atoi(e) > 0) ? atoi(e) : 8;
      }
      p.group_m = grp;
    }
    return dispatch_gemm_pair(tmA, tmBp, p, epilogue, static_cast<cudaStream_t>(stream));
  }
  // tile width: 256 unless that leaves SMs idle on a skinny problem (MV_GEMM_BN=64|256 forces one, for tests)
  const int num_m = (M + BM - 1) / BM;
  int bn = (num_m == 1 && (N + 255) / 256 < sm_count()) ? 64 : 256;   // measured: M = 512 is faster on 256-wide tiles
  {
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("MV_GEMM_BN");
      forced = e ? atoi(e) : 0;
    }
    if (forced == 64 || forced == 256) bn = forced;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t str[2] = {2, static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(bn)};
    rc = make_tmap_bf16(&tmB, W, 2, dims, str, box, true);
    if (rc != MV_OK) return rc;
  }
  GemmParams p;
  p.bias = bias;
  p.gate = gate;
  p.out = out;
  p.ldo = ldo;
  p.M = M;
  p.N = N;
  p.K = K;
  p.num_m = num_m;
  p.num_n = (N + bn - 1) / bn;
  p.num_tiles = p.num_m * p.num_n;
  p.num_kb = (K + BK - 1) / BK;
  p.a_kblock = a_kblock;
  p.f16 = f16;
  p.group_m = kGroupM;
  {
    static int stream = -1;   // MV_GEMM_STREAM=1 turns the evict-first accesses on; measured neutral (+-3 %), default off
    if (stream < 0) {
      const char* e = getenv("MV_GEMM_STREAM");
      stream = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    }
    p.stream_out = stream;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return bn == 64 ? dispatch_gemm<64>(tmA, tmB, p, epilogue, st) : dispatch_gemm<256>(tmA, tmB, p, epilogue, st);
}

extern "C" int mv_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out,
                            int64_t ldo, const float* gate, int M, int N, int K, int epilogue,
                            mv_stream_t stream) {
  return gemm_impl(A, lda, 0, 0, W, ldw, bias, out, ldo, gate, M, N, K, epilogue, stream);
}

extern "C" int mv_gemm_bf16_ksplit(const void* A, int64_t lda, int64_t a_block_stride, int a_kblock, const void* W,
                                   int64_t ldw, const float* bias, void* out, int64_t ldo, const float* gate, int M,
                                   int N, int K, int epilogue, mv_stream_t stream) {
  return gemm_impl(A, lda, a_block_stride, a_kblock, W, ldw, bias, out, ldo, gate, M, N, K, epilogue, stream);
}

extern "C" int mv_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* out,
                           int64_t ldo, const float* gate, int M, int N, int K, int epilogue, mv_stream_t stream) {
  return gemm_impl(A, lda, 0, 0, W, ldw, bias, out, ldo, gate, M, N, K, epilogue, stream, 1);
}
