// K5: merge of per-shard exact top-k lists.  Every input list is already in canonical order, so an
// element's final rank is the number of better keys in all lists, found with one binary search per
// list.  No sort, payload (label) rides along.
//
// The per-rank lists may be DENSE ([G][nq][k] per array) or PACKED (rank g's block =
// [idx nq*k int64 | sims nq*k fp32 | labels nq*k int32], blocks `block_bytes` apart -- exactly what
// one all-gather of the ranks' packed result buffers produces, or what K3's tail stores into the
// peer region, read in place).  The kernel takes one byte stride per array, which covers both.
//
// Multi-GPU tail in ONE kernel (hcir_peer_merge_vote): every query's CTA first waits for all ranks'
// blocks of the step (arrival counters of the local peer region, ld.acquire.sys, bounded spin), merges
// them in place, votes on the merged neighbour labels, and the last CTA completes the step (*step = st).
// That replaces four launches of the round-1 step (push, wait, merge, vote).
#include <math.h>

#include "hcir_common.cuh"

namespace hcir {

struct MergeTail {
  int64_t* hdr;        // local peer region header: wait for the step's arrivals first (null: no wait)
  int64_t* step;       // completed-step counter (read: st = *step + 1; the last CTA writes st back)
  int64_t timeout_ns;
  int world;
  int num_classes;     // vote on the merged labels when pred != null
  float T;
  const int64_t* classes;
  int64_t* pred;
};

// grid nq; block 128.  dynamic smem: G*k keys + k sims + k labels.
__global__ void __launch_bounds__(128)
merge_topk_kernel(const char* __restrict__ gsim, const char* __restrict__ gidx, const char* __restrict__ glab,
                  size_t stride_sim, size_t stride_idx, size_t stride_lab, int G, int64_t nq, int k,
                  float* __restrict__ out_sim, int64_t* __restrict__ out_idx, int32_t* __restrict__ out_lab,
                  size_t parity_stride, const MergeTail mt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_wait();
  int64_t st = 0;
  if (mt.hdr != nullptr) {  // peer exchange: this step's blocks live in parity (st & 1) once every rank arrived
    st = *mt.step + 1;
    if (threadIdx.x < kWarp) {
      const bool ok = peer_wait_arrivals(mt.hdr, mt.world, st, mt.timeout_ns, threadIdx.x);
      if (!ok && threadIdx.x == 0) mt.hdr[kPeerHdrError] = st;
    }
    __syncthreads();
    const size_t off = static_cast<size_t>(st & 1) * parity_stride;
    gsim += off;
    gidx += off;
    if (glab) glab += off;
  }
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);  // [G][k], each list descending
  float* rsim = reinterpret_cast<float*>(keys + static_cast<size_t>(G) * k);  // [k] merged similarities
  int32_t* rlab = reinterpret_cast<int32_t*>(rsim + k);                       // [k] merged labels
  const int64_t q = blockIdx.x;
  const int total = G * k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {  // fewer than k valid entries: vote on what there is
    rsim[i] = -INFINITY;
    rlab[i] = -1;
  }
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int g = i / k, j = i - g * k;
    const int64_t src = q * k + j;
    // .cg: the slots are written by peers over NVLink; never serve them from a stale L1 line
    const int64_t id = __ldcg(reinterpret_cast<const int64_t*>(gidx + g * stride_idx) + src);
    const float sv = __ldcg(reinterpret_cast<const float*>(gsim + g * stride_sim) + src);
    // id < 0 marks an empty slot (a shard with fewer than k rows): worst possible key
    keys[i] = (id < 0) ? 0ull : make_key(sv, static_cast<uint32_t>(id));
  }
  __syncthreads();
  const bool vote = (mt.pred != nullptr) && (glab != nullptr);
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
    HCIR_DEV_CHECK(rank >= 0 && rank < total);
    if (rank < k) {
      const int g = i / k, j = i - g * k;
      const int64_t src = q * k + j;
      const float sv = __ldcg(reinterpret_cast<const float*>(gsim + g * stride_sim) + src);
      out_sim[q * k + rank] = sv;
      out_idx[q * k + rank] = __ldcg(reinterpret_cast<const int64_t*>(gidx + g * stride_idx) + src);
      if (glab) {
        const int32_t l = __ldcg(reinterpret_cast<const int32_t*>(glab + g * stride_lab) + src);
        if (out_lab) out_lab[q * k + rank] = l;
        if (vote) {
          rsim[rank] = sv;
          rlab[rank] = l;
        }
      }
    }
  }
  if (vote) {
    __syncthreads();
    if (threadIdx.x < kWarp) {
      const int best_c = warp_vote(rsim, rlab, k, mt.num_classes, mt.T, threadIdx.x);
      if (threadIdx.x == 0) mt.pred[q] = mt.classes ? mt.classes[best_c] : static_cast<int64_t>(best_c);
    }
  }
  if (mt.hdr != nullptr) {  // the last CTA completes the step
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long* done = reinterpret_cast<unsigned long long*>(mt.hdr) + kPeerHdrConsDone;
      __threadfence();
      if (atomicAdd(done, 1ull) == static_cast<unsigned long long>(nq) - 1ull) {
        *done = 0ull;
        *mt.step = st;
        mt.hdr[kPeerHdrStep] = st;
      }
    }
  }
}

static int merge_launch(const void* gsim, const void* gidx, const void* glab, size_t stride_sim, size_t stride_idx,
                        size_t stride_lab, int G, int64_t nq, int k, float* out_sim, int64_t* out_idx,
                        int32_t* out_lab, hcir_stream_t stream, size_t parity_stride = 0,
                        const MergeTail* tail = nullptr) {
  HCIR_REQUIRE(G > 0 && nq >= 0 && k > 0, "merge_topk: bad shape G=%d nq=%lld k=%d", G, (long long)nq, k);
  HCIR_REQUIRE((gsim && gidx && out_sim && out_idx) || nq == 0, "merge_topk: null pointer");
  HCIR_REQUIRE((out_lab == nullptr) || (glab != nullptr), "merge_topk: out_lab without gathered labels");
  const size_t smem = static_cast<size_t>(G) * k * sizeof(uint64_t) + static_cast<size_t>(k) * 8;
  HCIR_REQUIRE(smem <= 200 * 1024, "merge_topk: G*k=%d too large", G * k);
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  MergeTail mt{};
  if (tail != nullptr) mt = *tail;
  HCIR_REQUIRE(mt.pred == nullptr || (glab != nullptr && mt.num_classes > 0), "merge_topk: a vote needs labels");
  HCIR_CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  HCIR_CUDA_TRY(launch_pdl(merge_topk_kernel, dim3(static_cast<unsigned>(nq)), dim3(128), smem,
                           static_cast<cudaStream_t>(stream), static_cast<const char*>(gsim),
                           static_cast<const char*>(gidx), static_cast<const char*>(glab), stride_sim, stride_idx,
                           stride_lab, G, nq, k, out_sim, out_idx, out_lab, parity_stride, mt));
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
                        size_t parity_stride, const hcir::MergeTail* tail) {
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
                      out_sim, out_idx, with_labels ? out_lab : nullptr, stream, parity_stride, tail);
}

extern "C" int hcir_merge_topk_packed(const void* gathered, int G, int64_t nq, int k, int with_labels,
                                      size_t rank_stride_bytes, float* out_sim, int64_t* out_idx, int32_t* out_lab,
                                      hcir_stream_t stream) {
  return merge_packed(gathered, G, nq, k, with_labels, rank_stride_bytes, out_sim, out_idx, out_lab, stream, 0,
                      nullptr);
}

extern "C" int hcir_peer_merge_vote(void* region_local, int G, int64_t nq, int k, int with_labels, size_t slot_bytes,
                                    int64_t* step, int64_t timeout_ns, float* out_sim, int64_t* out_idx,
                                    int32_t* out_lab, int num_classes, float T, const int64_t* classes, int64_t* pred,
                                    hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(region_local != nullptr && step != nullptr, "peer_merge_vote: null pointer");
  HCIR_REQUIRE(G >= 1 && G <= kPeerMax, "peer_merge_vote: bad world %d", G);
  HCIR_REQUIRE(pred == nullptr || with_labels, "peer_merge_vote: a vote needs labels in the blocks");
  const size_t stride = (slot_bytes + 255) / 256 * 256;
  char* data = static_cast<char*>(region_local) + hcir_peer_slot_offset(G, slot_bytes, 0, 0);
  MergeTail mt{};
  mt.hdr = static_cast<int64_t*>(region_local);
  mt.step = step;
  mt.timeout_ns = timeout_ns;
  mt.world = G;
  mt.num_classes = num_classes;
  mt.T = T;
  mt.classes = classes;
  mt.pred = pred;
  return merge_packed(data, G, nq, k, with_labels, stride, out_sim, out_idx, out_lab, stream,
                      static_cast<size_t>(G) * stride, &mt);
}
