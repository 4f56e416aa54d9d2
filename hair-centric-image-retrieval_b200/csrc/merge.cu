// K5: merge of per-shard exact top-k lists (after the NCCL all-gather).  Every input list
// is already in canonical order, so an element's final rank is the number of better keys
// in all lists, found with one binary search per list.  No sort, payload (label) rides along.
//
// The per-rank lists may be DENSE ([G][nq][k] per array) or PACKED (rank g's block =
// [idx nq*k int64 | sims nq*k fp32 | labels nq*k int32], blocks `block_bytes` apart -- exactly what
// one all-gather of the ranks' packed result buffers produces, read in place).  The kernel takes
// one byte stride per array, which covers both.
#include "hcir_common.cuh"

namespace hcir {

// grid nq; block 128.  dynamic smem: G*k keys.
__global__ void __launch_bounds__(128)
merge_topk_kernel(const char* __restrict__ gsim, const char* __restrict__ gidx, const char* __restrict__ glab,
                  size_t stride_sim, size_t stride_idx, size_t stride_lab, int G, int64_t nq, int k,
                  float* __restrict__ out_sim, int64_t* __restrict__ out_idx, int32_t* __restrict__ out_lab,
                  const int64_t* __restrict__ step, size_t parity_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (step != nullptr) {  // peer exchange (peer.cu): this step's blocks live in parity (step & 1)
    const size_t off = static_cast<size_t>(*step & 1) * parity_stride;
    gsim += off;
    gidx += off;
    if (glab) glab += off;
  }
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [G][k], each list descending
  const int64_t q = blockIdx.x;
  const int total = G * k;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int g = i / k, j = i - g * k;
    const int64_t src = q * k + j;
    const int64_t id = reinterpret_cast<const int64_t*>(gidx + g * stride_idx)[src];
    const float sv = reinterpret_cast<const float*>(gsim + g * stride_sim)[src];
    // id < 0 marks an empty slot (a shard with fewer than k rows): worst possible key
    keys[i] = (id < 0) ? 0ull : make_key(sv, static_cast<uint32_t>(id));
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
      const int64_t src = q * k + j;
      out_sim[q * k + rank] = reinterpret_cast<const float*>(gsim + g * stride_sim)[src];
      out_idx[q * k + rank] = reinterpret_cast<const int64_t*>(gidx + g * stride_idx)[src];
      if (out_lab) out_lab[q * k + rank] = reinterpret_cast<const int32_t*>(glab + g * stride_lab)[src];
    }
  }
}

static int merge_launch(const void* gsim, const void* gidx, const void* glab, size_t stride_sim, size_t stride_idx,
                        size_t stride_lab, int G, int64_t nq, int k, float* out_sim, int64_t* out_idx,
                        int32_t* out_lab, hcir_stream_t stream, const int64_t* step = nullptr,
                        size_t parity_stride = 0) {
  HCIR_REQUIRE(G > 0 && nq >= 0 && k > 0, "merge_topk: bad shape G=%d nq=%lld k=%d", G, (long long)nq, k);
  HCIR_REQUIRE((gsim && gidx && out_sim && out_idx) || nq == 0, "merge_topk: null pointer");
  HCIR_REQUIRE((out_lab == nullptr) || (glab != nullptr), "merge_topk: out_lab without gathered labels");
  const size_t smem = static_cast<size_t>(G) * k * sizeof(uint64_t);
  HCIR_REQUIRE(smem <= 200 * 1024, "merge_topk: G*k=%d too large", G * k);
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  HCIR_CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  merge_topk_kernel<<<static_cast<unsigned>(nq), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const char*>(gsim), static_cast<const char*>(gidx), static_cast<const char*>(glab), stride_sim,
      stride_idx, stride_lab, G, nq, k, out_sim, out_idx, out_lab, step, parity_stride);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

}  // namespace hcir

extern "C" int hcir_merge_topk(const float* gathered_sim, const int64_t* gathered_idx, const int32_t* gathered_lab,
                               int G, int64_t nq, int k, float* out_sim, int64_t* out_idx, int32_t* out_lab,
                               hcir_stream_t stream) {
  const size_t e = static_cast<size_t>(nq > 0 ? nq : 0) * static_cast<size_t>(k > 0 ? k : 0);
  return hcir::merge_launch(gathered_sim, gathered_idx, gathered_lab, e * 4, e * 8, e * 4, G, nq, k, out_sim,
                            out_idx, out_lab, stream);
}

extern "C" size_t hcir_packed_block_bytes(int64_t nq, int k, int with_labels) {
  if (nq <= 0 || k <= 0) return 0;
  const size_t b = static_cast<size_t>(nq) * static_cast<size_t>(k) * (with_labels ? 16 : 12);
  return (b + 15) / 16 * 16;  // every rank's block starts 16-byte aligned in the gathered buffer
}

static int merge_packed(const void* gathered, int G, int64_t nq, int k, int with_labels, size_t rank_stride_bytes,
                        float* out_sim, int64_t* out_idx, int32_t* out_lab, hcir_stream_t stream,
                        const int64_t* step, size_t parity_stride) {
  using namespace hcir;
  HCIR_REQUIRE(gathered != nullptr || nq == 0, "merge_topk_packed: null pointer");
  const size_t e = static_cast<size_t>(nq > 0 ? nq : 0) * static_cast<size_t>(k > 0 ? k : 0);
  const size_t min_block = hcir_packed_block_bytes(nq, k, with_labels);
  const size_t block = rank_stride_bytes ? rank_stride_bytes : min_block;
  HCIR_REQUIRE(block >= min_block && block % 8 == 0, "merge_topk_packed: rank stride %zu < block %zu or misaligned",
               block, min_block);
  const char* base = static_cast<const char*>(gathered);
  // block layout: idx (8-byte aligned first) | sims | labels
  return merge_launch(base + e * 8, base, with_labels ? base + e * 12 : nullptr, block, block, block, G, nq, k,
                      out_sim, out_idx, with_labels ? out_lab : nullptr, stream, step, parity_stride);
}

extern "C" int hcir_merge_topk_packed(const void* gathered, int G, int64_t nq, int k, int with_labels,
                                      size_t rank_stride_bytes, float* out_sim, int64_t* out_idx, int32_t* out_lab,
                                      hcir_stream_t stream) {
  return merge_packed(gathered, G, nq, k, with_labels, rank_stride_bytes, out_sim, out_idx, out_lab, stream, nullptr,
                      0);
}

extern "C" int hcir_merge_topk_peer(const void* region_local, int G, int64_t nq, int k, int with_labels,
                                    size_t slot_bytes, const int64_t* step, float* out_sim, int64_t* out_idx,
                                    int32_t* out_lab, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(region_local != nullptr && step != nullptr, "merge_topk_peer: null pointer");
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  const char* data = static_cast<const char*>(region_local) + hcir_peer_slot_offset(G, slot_bytes, 0, 0);
  return merge_packed(data, G, nq, k, with_labels, stride, out_sim, out_idx, out_lab, stream, step,
                      static_cast<size_t>(G) * stride);
}
