// K3: per-query candidate selection, adaptive fp32 re-score, exact sort, certification.
// One CTA per query.  HBM-bound on the fp32 row gathers: ~ (k + 16 + near-boundary
// candidates) rows of 4*ld bytes per query; everything else lives in shared memory.
#include <math.h>

#include "hcir_common.cuh"

namespace hcir {

constexpr int kSelThreads = 256;
constexpr int kRound1Slack = 16;

struct SelSmem {  // offsets (bytes) into dynamic shared memory
  size_t keys, sel, fk, pos, qrow, hist, scratch, total;
};

static SelSmem sel_smem_layout(int tmax, int kc, int ld) {
  SelSmem L;
  size_t o = 0;
  L.keys = o; o += static_cast<size_t>(tmax) * 8;
  L.sel = o; o += static_cast<size_t>(kc) * 8;
  L.fk = o; o += static_cast<size_t>(kc) * 8;
  L.qrow = o; o += static_cast<size_t>(ld) * 4;
  L.pos = o; o += static_cast<size_t>(kc) * 4;
  L.hist = o; o += 256 * 4;
  L.scratch = o; o += 16 * 4;
  L.total = o;
  return L;
}

__global__ void __launch_bounds__(kSelThreads)
select_rescore_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int ld, int64_t nq,
                      int64_t ng, int k, int64_t idx_offset, int nsplit, int cap, int kc,
                      const int32_t* __restrict__ counts, const uint64_t* __restrict__ cand,
                      const float* __restrict__ q_delta, float g_delta_max, float eps_acc,
                      float* __restrict__ out_sim, int64_t* __restrict__ out_idx,
                      int32_t* __restrict__ uncert_list, int32_t* __restrict__ uncert_count, SelSmem L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + L.keys);
  uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw + L.sel);
  uint64_t* fk = reinterpret_cast<uint64_t*>(smem_raw + L.fk);
  int32_t* pos = reinterpret_cast<int32_t*>(smem_raw + L.pos);
  float* qrow = reinterpret_cast<float*>(smem_raw + L.qrow);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + L.hist);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw + L.scratch);  // [0..3] block_select, [4..] ours

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kSelThreads / kWarp;
  const int64_t q = blockIdx.x;
  const int ld4 = ld >> 2;

  // ---- stage the fp32 query row and the split lists --------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(q32 + q * static_cast<int64_t>(ld));
    for (int c = tid; c < ld4; c += kSelThreads) reinterpret_cast<float4*>(qrow)[c] = __ldg(src + c);
  }
  int total = 0;
  for (int s = 0; s < nsplit; ++s) {
    const int c = counts[q * nsplit + s];
    const uint64_t* src = cand + (q * nsplit + s) * static_cast<int64_t>(cap);
    for (int i = tid; i < c; i += kSelThreads) keys[total + i] = src[i];
    total += c;
  }
  __syncthreads();

  // ---- keep the kc best by bf16 score ----------------------------------------------------
  int ncand;
  float tprime = -INFINITY;
  bool all_in;
  if (total > kc) {
    const uint64_t thr_c = block_select(keys, total, kc, sel, hist, scratch);
    tprime = key_sim(thr_c);
    ncand = kc;
    all_in = false;
  } else {
    for (int i = tid; i < total; i += kSelThreads) sel[i] = keys[i];
    ncand = total;
    all_in = (static_cast<int64_t>(total) == ng);
  }
  for (int i = tid; i < ncand; i += kSelThreads) fk[i] = 0ull;
  if (tid == 0) { scratch[4] = 0; scratch[5] = 0; scratch[6] = 0; scratch[7] = 0; }
  __syncthreads();

  const float dq = q_delta ? q_delta[q] : 0.0f;
  const float eps = g_delta_max * (1.0f + dq) + dq * (1.0f + 1e-6f) + eps_acc;

  // ---- round 1: the k + slack best bf16 candidates ----------------------------------------
  const int k1 = (ncand < k + kRound1Slack) ? ncand : k + kRound1Slack;
  uint64_t thr1 = 0ull;
  if (ncand > k1) thr1 = block_select(sel, ncand, k1, nullptr, hist, scratch);
  for (int j = tid; j < ncand; j += kSelThreads) {
    if (sel[j] >= thr1) pos[atomicAdd(&scratch[4], 1u)] = j;
  }
  __syncthreads();
  const int n1 = static_cast<int>(scratch[4]);
  for (int t = warp; t < n1; t += kWarps) {
    const int j = pos[t];
    const uint32_t row = key_idx(sel[j]);
    const float s = canonical_dot(reinterpret_cast<const float4*>(qrow),
                                  reinterpret_cast<const float4*>(g32 + static_cast<int64_t>(row) * ld), ld4, lane);
    if (lane == 0) fk[j] = make_key(s, row);
  }
  __syncthreads();
  // k-th best fp32 score so far (a lower bound of the final k-th best)
  for (int t = tid; t < n1; t += kSelThreads) {
    const uint64_t mine = fk[pos[t]];
    int rank = 0;
    for (int i = 0; i < n1; ++i) rank += (fk[pos[i]] > mine) ? 1 : 0;
    if (rank == k - 1) scratch[5] = __float_as_uint(key_sim(mine));
  }
  __syncthreads();
  const float sk1 = (n1 >= k) ? __uint_as_float(scratch[5]) : -INFINITY;

  // ---- round 2: every other candidate whose bf16 score could still reach the top-k -------
  for (int j = tid; j < ncand; j += kSelThreads) {
    if (sel[j] < thr1 && key_sim(sel[j]) + eps >= sk1) pos[n1 + atomicAdd(&scratch[6], 1u)] = j;
  }
  __syncthreads();
  const int n2 = static_cast<int>(scratch[6]);
  for (int t = warp; t < n2; t += kWarps) {
    const int j = pos[n1 + t];
    const uint32_t row = key_idx(sel[j]);
    const float s = canonical_dot(reinterpret_cast<const float4*>(qrow),
                                  reinterpret_cast<const float4*>(g32 + static_cast<int64_t>(row) * ld), ld4, lane);
    if (lane == 0) fk[j] = make_key(s, row);
  }
  __syncthreads();

  // ---- exact order of the re-scored set, emit top-k ---------------------------------------
  const int nr = n1 + n2;
  for (int t = tid; t < nr; t += kSelThreads) {
    const uint64_t mine = fk[pos[t]];
    int rank = 0;
    for (int i = 0; i < nr; ++i) rank += (fk[pos[i]] > mine) ? 1 : 0;
    if (rank < k) {
      const float s = key_sim(mine);
      out_sim[q * k + rank] = s;
      out_idx[q * k + rank] = static_cast<int64_t>(key_idx(mine)) + idx_offset;
      if (rank == k - 1) scratch[7] = __float_as_uint(s);
    }
  }
  __syncthreads();
  if (tid == 0) {
    const float sk = (nr >= k) ? __uint_as_float(scratch[7]) : -INFINITY;
    const bool certified = (nr >= k) && (all_in || (tprime + eps < sk));
    if (!certified) uncert_list[atomicAdd(uncert_count, 1)] = static_cast<int32_t>(q);
  }
}

}  // namespace hcir

extern "C" int hcir_select_rescore(const float* q_f32, const float* g_f32, int ld, int64_t nq, int64_t ng, int k,
                                   int64_t idx_offset, const hcir_plan_t* plan, const void* workspace,
                                   const float* q_delta, float g_delta_max, float eps_acc, float* out_sim,
                                   int64_t* out_idx, int32_t* uncert_list, int32_t* uncert_count,
                                   hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(plan != nullptr, "select_rescore: null plan");
  HCIR_REQUIRE(ld > 0 && ld % 64 == 0, "select_rescore: ld=%d must be a positive multiple of 64", ld);
  HCIR_REQUIRE(nq >= 0 && ng > 0, "select_rescore: bad shape");
  HCIR_REQUIRE(k > 0 && k <= ng && k <= plan->kc, "select_rescore: need 1 <= k=%d <= min(ng=%lld, kc=%d)", k,
               (long long)ng, plan->kc);
  HCIR_REQUIRE(q_f32 && g_f32 && workspace && out_sim && out_idx && uncert_list && uncert_count,
               "select_rescore: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  const int tmax = plan->nsplit * plan->kc;
  const SelSmem L = sel_smem_layout(tmax, plan->kc, ld);
  HCIR_REQUIRE(L.total <= 220 * 1024, "select_rescore: nsplit*kc=%d needs %zu B of shared memory", tmax, L.total);
  HCIR_CUDA_TRY(cudaFuncSetAttribute(select_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(L.total)));
  const int32_t* counts = reinterpret_cast<const int32_t*>(static_cast<const char*>(workspace) + plan->counts_off);
  const uint64_t* cand = reinterpret_cast<const uint64_t*>(static_cast<const char*>(workspace) + plan->keys_off);
  select_rescore_kernel<<<static_cast<unsigned>(nq), kSelThreads, L.total, static_cast<cudaStream_t>(stream)>>>(
      q_f32, g_f32, ld, nq, ng, k, idx_offset, plan->nsplit, plan->cap, plan->kc, counts, cand, q_delta,
      g_delta_max, eps_acc, out_sim, out_idx, uncert_list, uncert_count, L);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}
