// K5: merge of per-shard exact top-k lists (after the NCCL all-gather).  Every input list
// is already in canonical order, so an element's final rank is the number of better keys
// in all lists, found with one binary search per list.  No sort, payload (label) rides along.
#include "hcir_common.cuh"

namespace hcir {

// grid nq; block 128.  dynamic smem: G*k keys.
__global__ void __launch_bounds__(128)
merge_topk_kernel(const float* __restrict__ gsim, const int64_t* __restrict__ gidx,
                  const int32_t* __restrict__ glab, int G, int64_t nq, int k,
                  float* __restrict__ out_sim, int64_t* __restrict__ out_idx, int32_t* __restrict__ out_lab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [G][k], each list descending
  const int64_t q = blockIdx.x;
  const int total = G * k;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int g = i / k, j = i - g * k;
    const int64_t src = (static_cast<int64_t>(g) * nq + q) * k + j;
    const int64_t id = gidx[src];
    // id < 0 marks an empty slot (a shard with fewer than k rows): worst possible key
    keys[i] = (id < 0) ? 0ull : make_key(gsim[src], static_cast<uint32_t>(id));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const uint64_t mine = keys[i];
    if (mine == 0ull) continue;
    int rank = 0;
    for (int g = 0; g < G; ++g) {
      const uint64_t* lst = keys + g * k;
      int lo = 0, hi = k;  // first position with lst[pos] <= mine  == number of keys > mine
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] > mine) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      const int g = i / k, j = i - g * k;
      const int64_t src = (static_cast<int64_t>(g) * nq + q) * k + j;
      out_sim[q * k + rank] = gsim[src];
      out_idx[q * k + rank] = gidx[src];
      if (out_lab) out_lab[q * k + rank] = glab[src];
    }
  }
}

}  // namespace hcir

extern "C" int hcir_merge_topk(const float* gathered_sim, const int64_t* gathered_idx, const int32_t* gathered_lab,
                               int G, int64_t nq, int k, float* out_sim, int64_t* out_idx, int32_t* out_lab,
                               hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(G > 0 && nq >= 0 && k > 0, "merge_topk: bad shape G=%d nq=%lld k=%d", G, (long long)nq, k);
  HCIR_REQUIRE((gathered_sim && gathered_idx && out_sim && out_idx) || nq == 0, "merge_topk: null pointer");
  HCIR_REQUIRE((out_lab == nullptr) || (gathered_lab != nullptr), "merge_topk: out_lab without gathered_lab");
  const size_t smem = static_cast<size_t>(G) * k * sizeof(uint64_t);
  HCIR_REQUIRE(smem <= 200 * 1024, "merge_topk: G*k=%d too large", G * k);
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  HCIR_CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  merge_topk_kernel<<<static_cast<unsigned>(nq), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      gathered_sim, gathered_idx, gathered_lab, G, nq, k, out_sim, out_idx, out_lab);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}
