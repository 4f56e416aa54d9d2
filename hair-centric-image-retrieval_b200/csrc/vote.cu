// K4: kNN vote (uniform == sklearn `_mode`, or temperature-weighted extension) and the
// neighbour-label gather that feeds it.  Tiny, latency-bound kernels: one warp per query.
#include "hcir_common.cuh"

namespace hcir {

__global__ void gather_labels_kernel(const int64_t* __restrict__ idx, int64_t count,
                                     const int32_t* __restrict__ labels, int64_t n_labels,
                                     int64_t idx_offset, int32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t r = idx[i] - idx_offset;
  out[i] = (r >= 0 && r < n_labels) ? labels[r] : -1;
}

constexpr int kVoteWarps = 4;

// Lane l owns classes l, l+32, ...; each lane scans the k neighbours in rank order, so the
// fp32 accumulation order per class is j = 0..k-1 (deterministic, matches the oracle).
__global__ void __launch_bounds__(kVoteWarps* kWarp)
vote_kernel(const float* __restrict__ sims, const int32_t* __restrict__ nbr, int64_t nq, int k,
            int num_classes, float T, int32_t* __restrict__ pred, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * kVoteWarps + (threadIdx.x >> 5);
  if (q >= nq) return;
  const float* s = sims + q * k;
  const int32_t* l = nbr + q * k;
  const bool weighted = T > 0.0f;
  const float s0 = weighted ? s[0] : 0.0f;
  const float inv_t = weighted ? 1.0f / T : 0.0f;
  float best = -1.0f;
  int best_c = 0x7FFFFFFF;
  for (int c = lane; c < num_classes; c += kWarp) {
    float acc = 0.0f;
    for (int j = 0; j < k; ++j) {
      if (l[j] == c) acc += weighted ? expf((s[j] - s0) * inv_t) : 1.0f;
    }
    if (scores) scores[q * num_classes + c] = acc;
    if (acc > best) {  // ascending c per lane: strict > keeps the smallest class on ties
      best = acc;
      best_c = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(kFull, best, o);
    const int oc = __shfl_xor_sync(kFull, best_c, o);
    if (ob > best || (ob == best && oc < best_c)) {
      best = ob;
      best_c = oc;
    }
  }
  if (lane == 0) pred[q] = best_c;
}

// Fused tail of the classification step: neighbour-label gather -> vote -> class value, one warp
// per query (the labels of its k neighbours are staged in shared memory).  Same arithmetic and
// tie rule as gather_labels_kernel + vote_kernel; replaces four launches (gather, vote, int cast,
// class-table lookup) by one.  dynamic smem: kVoteWarps * k int32.
__global__ void __launch_bounds__(kVoteWarps* kWarp)
vote_idx_kernel(const float* __restrict__ sims, const int64_t* __restrict__ idx, const int32_t* __restrict__ labels,
                int64_t n_labels, int64_t idx_offset, int64_t nq, int k, int num_classes, float T,
                const int64_t* __restrict__ classes, int64_t* __restrict__ pred, int32_t* __restrict__ nbr_out,
                const int32_t* __restrict__ nbr_in) {
  extern __shared__ int32_t vote_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * kVoteWarps + warp;
  if (q >= nq) return;  // warp-uniform; only warp-level synchronisation below
  int32_t* l = vote_smem + static_cast<size_t>(warp) * k;
  for (int j = lane; j < k; j += kWarp) {
    int32_t lab;
    if (nbr_in) {  // labels already gathered (they travelled with the candidates)
      lab = nbr_in[q * k + j];
    } else {
      const int64_t r = idx[q * k + j] - idx_offset;
      lab = (r >= 0 && r < n_labels) ? labels[r] : -1;
    }
    l[j] = lab;
    if (nbr_out) nbr_out[q * k + j] = lab;
  }
  __syncwarp();
  const int best_c = warp_vote(sims + q * k, l, k, num_classes, T, lane);
  if (lane == 0) pred[q] = classes ? classes[best_c] : static_cast<int64_t>(best_c);
}

}  // namespace hcir

extern "C" int hcir_vote_idx(const float* sims, const int64_t* idx, const int32_t* labels, int64_t n_labels,
                             int64_t idx_offset, int64_t nq, int k, int num_classes, float T, const int64_t* classes,
                             int64_t* pred, int32_t* nbr_labels_out, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(nq >= 0 && k > 0 && num_classes > 0 && n_labels >= 0, "vote_idx: bad shape nq=%lld k=%d C=%d",
               (long long)nq, k, num_classes);
  HCIR_REQUIRE((sims && idx && labels && pred) || nq == 0, "vote_idx: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  const size_t smem = static_cast<size_t>(kVoteWarps) * k * sizeof(int32_t);
  HCIR_REQUIRE(smem <= 200 * 1024, "vote_idx: k=%d too large", k);
  HCIR_CUDA_TRY(cudaFuncSetAttribute(vote_idx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  vote_idx_kernel<<<static_cast<unsigned>(ceil_div_i64(nq, kVoteWarps)), kVoteWarps * kWarp, smem,
                    static_cast<cudaStream_t>(stream)>>>(sims, idx, labels, n_labels, idx_offset, nq, k, num_classes, T,
                                                         classes, pred, nbr_labels_out, nullptr);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

extern "C" int hcir_vote_classes(const float* sims, const int32_t* nbr_labels, int64_t nq, int k, int num_classes,
                                 float T, const int64_t* classes, int64_t* pred, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(nq >= 0 && k > 0 && num_classes > 0, "vote_classes: bad shape nq=%lld k=%d C=%d", (long long)nq, k,
               num_classes);
  HCIR_REQUIRE((sims && nbr_labels && pred) || nq == 0, "vote_classes: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  const size_t smem = static_cast<size_t>(kVoteWarps) * k * sizeof(int32_t);
  HCIR_REQUIRE(smem <= 200 * 1024, "vote_classes: k=%d too large", k);
  HCIR_CUDA_TRY(cudaFuncSetAttribute(vote_idx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
  vote_idx_kernel<<<static_cast<unsigned>(ceil_div_i64(nq, kVoteWarps)), kVoteWarps * kWarp, smem,
                    static_cast<cudaStream_t>(stream)>>>(sims, nullptr, nullptr, 0, 0, nq, k, num_classes, T, classes,
                                                         pred, nullptr, nbr_labels);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

extern "C" int hcir_gather_labels(const int64_t* idx, int64_t count, const int32_t* labels, int64_t n_labels,
                                  int64_t idx_offset, int32_t* out, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(count >= 0 && n_labels >= 0, "gather_labels: bad sizes");
  HCIR_REQUIRE((idx && labels && out) || count == 0, "gather_labels: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (count == 0) return HCIR_OK;
  const int threads = 256;
  gather_labels_kernel<<<static_cast<unsigned>(ceil_div_i64(count, threads)), threads, 0,
                         static_cast<cudaStream_t>(stream)>>>(idx, count, labels, n_labels, idx_offset, out);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

extern "C" int hcir_vote(const float* sims, const int32_t* nbr_labels, int64_t nq, int k, int num_classes, float T,
                         int32_t* pred, float* scores, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(nq >= 0 && k > 0 && num_classes > 0, "vote: bad shape nq=%lld k=%d C=%d", (long long)nq, k,
               num_classes);
  HCIR_REQUIRE((sims && nbr_labels && pred) || nq == 0, "vote: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  vote_kernel<<<static_cast<unsigned>(ceil_div_i64(nq, kVoteWarps)), kVoteWarps * kWarp, 0,
                static_cast<cudaStream_t>(stream)>>>(sims, nbr_labels, nq, k, num_classes, T, pred, scores);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}
