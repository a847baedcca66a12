#include "host_util.h"

#include <stdlib.h>
#include <string.h>

namespace mv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return MV_E_CUDA;
}

int require_sm100() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = MV_E_CUDA;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cached_dev = dev;
  if (major != 10) {
    set_error("movii_b200 requires an sm_100 (B200) device; device %d is sm_%d%d — no fallback path exists",
              dev, major, minor);
    cached_rc = MV_E_ARCH;
  } else {
    cached_rc = MV_OK;
  }
  return cached_rc;
}

static int g_roles_hi = -1;
constexpr int kDefaultRolesHi = 0;
bool roles_hi() {
  if (g_roles_hi < 0) {
    const char* e = getenv("MV_ROLES_HI");
    g_roles_hi = (e != nullptr && e[0] != 0) ? (atoi(e) != 0 ? 1 : 0) : kDefaultRolesHi;
  }
  return g_roles_hi == 1;
}
void set_roles_hi(int v) { g_roles_hi = v; }

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128) {
  return make_tmap_bf16_sw(map, base, rank, dims, strides_bytes, box, swizzle128 ? 128 : 0);
}

int make_tmap_bf16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  return make_tmap_bf16_es(map, base, rank, dims, strides_bytes, box, swizzle_bytes, nullptr);
}

int make_tmap_bf16_es(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes,
                      const uint32_t* elem_strides) {
  CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 128) swz = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (swizzle_bytes == 64) swz = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 32) swz = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes != 0) {
    set_error("unsupported TMA swizzle span %d", swizzle_bytes);
    return MV_E_SHAPE;
  }
  if (swizzle_bytes != 0 && static_cast<int>(box[0]) * 2 != swizzle_bytes) {
    set_error("TMA box inner extent %u B does not match the swizzle span %d B", box[0] * 2, swizzle_bytes);
    return MV_E_SHAPE;
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return MV_E_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return MV_E_SHAPE;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = (elem_strides != nullptr && i > 0) ? elem_strides[i] : 1;   // traversal stride: box[i] / es[i] elements land
    if (es[i] < 1 || es[i] > 8 || box[i] % es[i] != 0) {
      set_error("TMA traversal stride %u (dim %d, box %u) unsupported", es[i], i, box[i]);
      return MV_E_SHAPE;
    }
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i];
      if (strides_bytes[i] % 16 != 0) {
        set_error("TMA stride %llu (dim %d) is not a multiple of 16 bytes",
                  static_cast<unsigned long long>(strides_bytes[i]), i);
        return MV_E_SHAPE;
      }
    }
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                   gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu box %u,%u)",
              static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]),
              static_cast<unsigned long long>(rank > 1 ? dims[1] : 0), box[0], rank > 1 ? box[1] : 0);
    return MV_E_CUDA;
  }
  return MV_OK;
}

}  // namespace mv

extern "C" const char* mv_last_error(void) { return mv::g_err; }
extern "C" int mv_version(void) { return 100; }
extern "C" int mv_device_check(void) { return mv::require_sm100(); }
extern "C" int mv_roles_config(int hi) {
  if (hi >= 0) mv::set_roles_hi(hi != 0 ? 1 : 0);
  else if (hi == -2) mv::set_roles_hi(-1);
  return MV_OK;
}
