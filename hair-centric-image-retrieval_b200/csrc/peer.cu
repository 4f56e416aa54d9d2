// Candidate / result exchange over NVLink peer memory (SURVEY.md section 8e "later fusion").
//
// Every rank owns one REGION of device memory (cudaMalloc, exported with CUDA IPC and mapped by
// every other rank of the box), all regions laid out identically:
//
//   header (512 B, int64 words)   [0..15]  arrivals[r]   monotone count of "rank r's block landed"
//                                 [16..31] meta[0][r]    one int64 per rank riding along with a block
//                                 [32..47] meta[1][r]      (the sender's uncertified-query count)
//                                 [48]     step_seen     step of the last completed wait
//                                 [49]     error         != 0: a wait timed out
//   data   2 parities x world slots x slot_stride bytes:  slot (p, r) = rank r's block of a step
//                                 with (step & 1) == p
//
// push:  rank r stores its block into slot (step & 1, r) of EVERY region (its own included) with
//        16-byte stores over NVLink, then -- per CTA: barrier, system-scope fence -- bumps
//        arrivals[r] in every region.  No collective library call, no staging copy.
// wait:  spins until arrivals[r] >= step * ctas_per_push for every r.
// Two parities are enough: a rank can only be one step ahead of the slowest peer (its wait for
// step n+1 needs that peer's push n+1, which that peer enqueues after its own reads of step n).
// `step` lives in device memory and is incremented on the stream, so a captured CUDA graph
// replays the exchange unchanged.
#include <string.h>

#include "hcir_common.cuh"

namespace hcir {

constexpr int kPeerMax = 16;
constexpr int kHdrArrivals = 0, kHdrMeta = 16, kHdrStep = 48, kHdrError = 49;
constexpr size_t kHdrBytes = 512;
constexpr int kPushThreads = 256;

struct PeerPtrs {
  char* region[kPeerMax];
};

__device__ __forceinline__ int64_t ld_acquire_sys(const int64_t* p) {
  int64_t v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPushThreads)
peer_push_kernel(const uint4* __restrict__ src, size_t n16, PeerPtrs peers, int world, int rank, size_t slot_stride,
                 const int64_t* __restrict__ step, const int32_t* __restrict__ meta_src) {
  const int64_t st = *step;
  const size_t par = static_cast<size_t>(st & 1);
  const size_t slot_off = kHdrBytes + (par * world + rank) * slot_stride;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kPushThreads + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * kPushThreads) {
    const uint4 v = src[i];
    for (int g = 0; g < world; ++g) reinterpret_cast<uint4*>(peers.region[g] + slot_off)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < world) {
    const int64_t m = meta_src ? static_cast<int64_t>(*meta_src) : 0;
    reinterpret_cast<int64_t*>(peers.region[threadIdx.x])[kHdrMeta + par * kPeerMax + rank] = m;
  }
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();  // this CTA's stores (observed through the barrier) before the arrival
    atomicAdd_system(reinterpret_cast<unsigned long long*>(peers.region[threadIdx.x]) + kHdrArrivals + rank, 1ull);
  }
}

__global__ void __launch_bounds__(32)
peer_wait_kernel(int64_t* __restrict__ hdr, int world, const int64_t* __restrict__ step, int64_t ctas_per_push,
                 int64_t timeout_ns) {
  const int64_t st = *step;
  const int64_t target = st * ctas_per_push;
  bool ok = true;
  if (threadIdx.x < world) {
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(hdr + kHdrArrivals + threadIdx.x) < target) {
      __nanosleep(64);
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (static_cast<int64_t>(t1 - t0) > timeout_ns) {  // a peer never arrived: report, do not hang the GPU
        ok = false;
        break;
      }
    }
  }
  const bool all_ok = __all_sync(kFull, ok);
  if (threadIdx.x == 0) {
    hdr[kHdrStep] = st;
    if (!all_ok) hdr[kHdrError] = st;
  }
}

}  // namespace hcir

extern "C" size_t hcir_peer_region_bytes(int world, size_t slot_bytes) {
  if (world < 1 || world > hcir::kPeerMax) return 0;
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  return hcir::kHdrBytes + 2 * static_cast<size_t>(world) * stride;
}

extern "C" size_t hcir_peer_slot_offset(int world, size_t slot_bytes, int parity, int rank) {
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  return hcir::kHdrBytes + (static_cast<size_t>(parity & 1) * world + rank) * stride;
}

extern "C" int hcir_peer_push_ctas(size_t bytes) {
  const size_t n16 = (bytes + 15) / 16;
  const size_t want = (n16 + 4 * hcir::kPushThreads - 1) / (4 * hcir::kPushThreads);
  return static_cast<int>(want < 1 ? 1 : (want > 64 ? 64 : want));
}

extern "C" int hcir_peer_alloc(size_t bytes, void** ptr, void* ipc_handle_64) {
  using namespace hcir;
  HCIR_REQUIRE(ptr != nullptr && ipc_handle_64 != nullptr && bytes >= kHdrBytes, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  void* p = nullptr;
  HCIR_CUDA_TRY(cudaMalloc(&p, bytes));
  HCIR_CUDA_TRY(cudaMemset(p, 0, bytes));
  HCIR_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  memcpy(ipc_handle_64, &h, 64);
  *ptr = p;
  return HCIR_OK;
}

extern "C" int hcir_peer_open(const void* ipc_handle_64, void** ptr) {
  using namespace hcir;
  HCIR_REQUIRE(ptr != nullptr && ipc_handle_64 != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle_64, 64);
  HCIR_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HCIR_OK;
}

extern "C" int hcir_peer_close(void* ptr) {
  using namespace hcir;
  if (ptr != nullptr) HCIR_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return HCIR_OK;
}

extern "C" int hcir_peer_free(void* ptr) {
  using namespace hcir;
  if (ptr != nullptr) HCIR_CUDA_TRY(cudaFree(ptr));
  return HCIR_OK;
}

extern "C" int hcir_peer_push(const void* src, size_t bytes, void* const* regions, int world, int rank,
                              size_t slot_bytes, const int64_t* step, const int32_t* meta_src,
                              hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(src != nullptr && regions != nullptr && step != nullptr, "peer_push: null pointer");
  HCIR_REQUIRE(world >= 1 && world <= kPeerMax && rank >= 0 && rank < world, "peer_push: bad rank %d / world %d", rank,
               world);
  HCIR_REQUIRE(bytes > 0 && bytes <= slot_bytes && reinterpret_cast<uintptr_t>(src) % 16 == 0,
               "peer_push: block of %zu bytes (slot %zu) must be non-empty, fit the slot and be 16-byte aligned", bytes,
               slot_bytes);
  PeerPtrs pp{};
  for (int g = 0; g < world; ++g) {
    HCIR_REQUIRE(regions[g] != nullptr, "peer_push: region %d is null", g);
    pp.region[g] = static_cast<char*>(regions[g]);
  }
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  const int ctas = hcir_peer_push_ctas(bytes);
  peer_push_kernel<<<ctas, kPushThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), (bytes + 15) / 16, pp, world, rank, stride, step, meta_src);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

extern "C" int hcir_peer_wait(void* region_local, int world, const int64_t* step, int ctas_per_push,
                              int64_t timeout_ns, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(region_local != nullptr && step != nullptr && world >= 1 && world <= kPeerMax && ctas_per_push >= 1,
               "peer_wait: bad arguments");
  peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<int64_t*>(region_local), world, step,
                                                                    ctas_per_push, timeout_ns);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}
