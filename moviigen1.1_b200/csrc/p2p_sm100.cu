// NVLink peer-to-peer plumbing for the fused Ulysses exchange: IPC export/open of caller-owned device buffers and a
// flag-based cross-GPU barrier kernel.  The data movement itself is fused into the producing kernels
// (rmsnorm_rope_kernel / attention_fwd_kernel store straight into the destination rank's HBM through these mapped
// pointers) — there is no collective call on the data path.
#include "common.cuh"
#include "host_util.h"

namespace mv {

constexpr int kMaxPeers = 8;

struct BarrierParams {
  unsigned int* peer_flags[kMaxPeers];  // peer_flags[i] = base of rank i's flag array (mapped), [kMaxPeers] slots
  unsigned int* local_flags;
  int rank, world;
  unsigned int epoch;
};

// One thread per peer: publish "rank reached epoch" into every peer's flag array (release, system scope), then wait
// until every peer has published the same epoch locally (acquire).  Bounded wait -> trap instead of a hung GPU.
__global__ void sp_barrier_kernel(const BarrierParams p) {
  const int i = threadIdx.x;
  if (i >= p.world) return;
  __threadfence_system();
  unsigned int* dst = p.peer_flags[i] + p.rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(p.epoch) : "memory");
  const unsigned int* src = p.local_flags + i;
  const uint64_t t0 = globaltimer_ns();
  unsigned int v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    if (static_cast<int>(v - p.epoch) >= 0) break;
    if (globaltimer_ns() - t0 > 20000000000ull) {
      printf("mv: sp_barrier timeout rank %d waiting for rank %d epoch %u (have %u)\n", p.rank, i, p.epoch, v);
      __trap();
    }
  } while (true);
  __threadfence_system();
}

typedef CUresult (*PFN_getrange)(CUdeviceptr*, size_t*, CUdeviceptr);

}  // namespace mv

using namespace mv;

extern "C" int mv_ipc_export(const void* dptr, void* handle64, int64_t* offset, int64_t* alloc_bytes) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  static PFN_getrange fn = nullptr;
  if (!fn) {
    void* pfn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &pfn, cudaEnableDefault, &q) != cudaSuccess || pfn == nullptr) {
      set_error("cuMemGetAddressRange entry point unavailable");
      return MV_E_CUDA;
    }
    fn = reinterpret_cast<PFN_getrange>(pfn);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = fn(&base, &size, reinterpret_cast<CUdeviceptr>(dptr));
  if (r != CUDA_SUCCESS) {
    set_error("cuMemGetAddressRange failed (%d)", static_cast<int>(r));
    return MV_E_CUDA;
  }
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaIpcGetMemHandle failed (%s): the buffer must come from a cudaMalloc segment "
              "(PyTorch default allocator without expandable_segments)", cudaGetErrorString(e));
    return MV_E_CUDA;
  }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(dptr) - base);
  if (alloc_bytes) *alloc_bytes = static_cast<int64_t>(size);
  return MV_OK;
}

extern "C" int mv_ipc_open(const void* handle64, void** base_out) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(e, "cudaIpcOpenMemHandle");
  }
  *base_out = p;
  return MV_OK;
}

extern "C" int mv_ipc_close(void* base) {
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return MV_OK;
}

extern "C" int mv_sp_barrier(void* const* peer_flag_ptrs, void* local_flags, int rank, int world, unsigned int epoch,
                             mv_stream_t stream) {
  int rc = require_sm100();
  if (rc != MV_OK) return rc;
  MV_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "mv_sp_barrier: bad rank/world %d/%d", rank, world);
  BarrierParams p;
  for (int i = 0; i < kMaxPeers; ++i) p.peer_flags[i] = i < world ? reinterpret_cast<unsigned int*>(peer_flag_ptrs[i]) : nullptr;
  p.local_flags = reinterpret_cast<unsigned int*>(local_flags);
  p.rank = rank;
  p.world = world;
  p.epoch = epoch;
  sp_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(p);
  MV_CHECK_LAUNCH("sp_barrier_kernel");
  return MV_OK;
}
