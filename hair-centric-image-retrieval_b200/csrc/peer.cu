// Candidate / result exchange over NVLink peer memory (SURVEY.md section 8e "later fusion").
//
// Every rank owns one REGION of device memory (cudaMalloc, exported with CUDA IPC and mapped by
// every other rank of the box), all regions laid out identically:
//
//   header (512 B, int64 words; hcir_common.cuh)
//                                 [0..15]  arrivals[r]   steps whose block from rank r has landed here
//                                 [16..31] meta[0][r]    one int64 per rank riding along with a block
//                                 [32..47] meta[1][r]      (the sender's uncertified-query count)
//                                 [48]     step_seen     last completed step
//                                 [49]     error         != 0: a wait timed out at that step
//                                 [50/51]  done-CTA counters of the consumer / the standalone push
//   data   2 parities x world slots x slot_stride bytes:  slot (p, r) = rank r's block of a step
//                                 with (step & 1) == p
//
// Protocol.  `step` (device int64, one per channel) counts COMPLETED steps.  The producer of step
// st = *step + 1 stores rank r's block into slot (st & 1, r) of EVERY region (its own included) with
// plain stores over NVLink; its last CTA -- every CTA fences at system scope and counts itself done --
// writes the meta word and bumps arrivals[r] once in every region.  The producer is normally the TAIL
// OF K3 (select_rescore.cu: every query's CTA stores its own results, no packed block is staged
// locally); peer_push_kernel below is the standalone form for blocks that already sit in memory.
// The consumer (peer_wait_kernel, or the fused wait+merge+vote kernel in merge.cu) spins until
// arrivals[r] >= st for every r, reads the slots in place, and completes the step by writing
// *step = st.  Two parities are enough: a rank can only be one step ahead of the slowest peer (its
// wait for step n+1 needs that peer's block n+1, which that peer produces after it consumed step n).
// Because `step` lives in device memory, a captured CUDA graph replays the exchange unchanged.
#include <string.h>

#include "hcir_common.cuh"

namespace hcir {

constexpr int kPushThreads = 256;

struct PeerPtrs {
  char* region[kPeerMax];
};

__global__ void __launch_bounds__(kPushThreads)
peer_push_kernel(const uint4* __restrict__ src, size_t n16, PeerPtrs peers, int world, int rank, size_t slot_stride,
                 const int64_t* __restrict__ step, const int32_t* __restrict__ meta_src) {
  const int64_t st = *step + 1;
  const size_t par = static_cast<size_t>(st & 1);
  const size_t slot_off = kPeerHdrBytes + (par * world + rank) * slot_stride;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kPushThreads + threadIdx.x; i < n16;
       i += static_cast<size_t>(gridDim.x) * kPushThreads) {
    const uint4 v = src[i];
    for (int g = 0; g < world; ++g) reinterpret_cast<uint4*>(peers.region[g] + slot_off)[i] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's stores (observed through the barrier) before it counts itself done
    unsigned long long* done = reinterpret_cast<unsigned long long*>(peers.region[rank]) + kPeerHdrProdDone;
    if (atomicAdd(done, 1ull) == gridDim.x - 1) {  // last CTA: everybody's stores are ordered before the arrival
      __threadfence();
      *done = 0ull;
      const int64_t m = meta_src ? static_cast<int64_t>(*meta_src) : 0;
      for (int g = 0; g < world; ++g)
        reinterpret_cast<int64_t*>(peers.region[g])[kPeerHdrMeta + par * kPeerMax + rank] = m;
      __threadfence_system();
      for (int g = 0; g < world; ++g)
        atomicAdd_system(reinterpret_cast<unsigned long long*>(peers.region[g]) + kPeerHdrArrivals + rank, 1ull);
    }
  }
}

__global__ void __launch_bounds__(32)
peer_wait_kernel(int64_t* __restrict__ hdr, int world, int64_t* __restrict__ step, int64_t timeout_ns) {
  pdl_wait();
  const int64_t st = *step + 1;
  const bool all_ok = peer_wait_arrivals(hdr, world, st, timeout_ns, threadIdx.x);
  if (threadIdx.x == 0) {
    *step = st;
    hdr[kPeerHdrStep] = st;
    if (!all_ok) hdr[kPeerHdrError] = st;
  }
}

}  // namespace hcir

extern "C" size_t hcir_peer_region_bytes(int world, size_t slot_bytes) {
  if (world < 1 || world > hcir::kPeerMax) return 0;
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  return hcir::kPeerHdrBytes + 2 * static_cast<size_t>(world) * stride;
}

extern "C" size_t hcir_peer_slot_offset(int world, size_t slot_bytes, int parity, int rank) {
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  return hcir::kPeerHdrBytes + (static_cast<size_t>(parity & 1) * world + rank) * stride;
}

extern "C" int hcir_peer_push_ctas(size_t bytes) {
  const size_t n16 = (bytes + 15) / 16;
  const size_t want = (n16 + 4 * hcir::kPushThreads - 1) / (4 * hcir::kPushThreads);
  return static_cast<int>(want < 1 ? 1 : (want > 64 ? 64 : want));
}

extern "C" int hcir_peer_alloc(size_t bytes, void** ptr, void* ipc_handle_64) {
  using namespace hcir;
  HCIR_REQUIRE(ptr != nullptr && ipc_handle_64 != nullptr && bytes >= kPeerHdrBytes, "peer_alloc: bad arguments");
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

extern "C" int hcir_peer_wait(void* region_local, int world, int64_t* step, int64_t timeout_ns,
                              hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(region_local != nullptr && step != nullptr && world >= 1 && world <= kPeerMax,
               "peer_wait: bad arguments");
  HCIR_CUDA_TRY(launch_pdl(peer_wait_kernel, dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream),
                           static_cast<int64_t*>(region_local), world, step, timeout_ns));
  return HCIR_OK;
}
