// Host-side helpers: thread-local error message, CUDA error mapping, TMA tensor-map encoding
// (driver entry point fetched at run time so the library has no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/movii_b200.h"

namespace mv {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
// MV_OK if the current device is CC 10.x; caches the answer per device.
int require_sm100();
int sm_count();
// 1: role warps (TMA producer, MMA issuer) sit at the HIGHEST warp ids of the CTA — the SM sub-partition arbiter favours
// the highest warp id, so a pending TMA / tcgen05.mma issue is not queued behind the epilogue / softmax instruction
// streams that share its sub-partition.  MV_ROLES_HI / mv_roles_config(); default = the measured winner.
bool roles_hi();
void set_roles_hi(int v);

// Encodes a tiled bf16 tensor map.  dims/strides innermost first; strides in BYTES for dims 1..rank-1
// (dim 0 is contiguous).  Swizzle 128B requires box[0] * 2 bytes == 128.
int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);
// Same with an explicit swizzle span: 128, 64, 32 (bytes; must equal box[0] * 2) or 0 (none).
int make_tmap_bf16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);
// + traversal strides per dimension (elem_strides[0] ignored): dimension i delivers box[i] / elem_strides[i] elements,
// every elem_strides[i]-th one — the A operand of a strided convolution without a gather pass
int make_tmap_bf16_es(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes,
                      const uint32_t* elem_strides);

#define MV_CHECK_CUDA(expr)                                \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return mv::cuda_fail(_e, #expr); \
  } while (0)

#define MV_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      mv::set_error(__VA_ARGS__);    \
      return MV_E_SHAPE;             \
    }                                \
  } while (0)

#define MV_CHECK_LAUNCH(name)                                    \
  do {                                                           \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return mv::cuda_fail(_e, name);       \
  } while (0)

}  // namespace mv
