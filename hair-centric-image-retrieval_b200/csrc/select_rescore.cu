// K3: per-query candidate selection, adaptive fp32 re-score, exact sort, certification.
// One CTA per query.  Bandwidth-bound: one streaming read of the query's candidate lists
// (a second one, when the lists exceed the staging buffer, hits L2) plus ~ (k + 16 +
// near-boundary candidates) fp32 gallery rows of 4*ld bytes; everything else is shared memory.
#include <math.h>

#include "hcir_common.cuh"

namespace hcir {

constexpr int kSelThreads = 256;
constexpr int kSelWarps = kSelThreads / kWarp;
constexpr int kRound1Slack = 16;
constexpr int kBins = 2048;       // histogram bins of the streaming selection
constexpr int kStageExtra = 512;  // staging room beyond kc for the bin that holds the kc-th best

struct SelSmem {  // offsets (bytes) into dynamic shared memory
  size_t stage, sel, fk, qrow, pos, hist, offs, scratch, total;
  int stage_cap;
};

static SelSmem sel_smem_layout(int nsplit, int kc, int ld) {
  SelSmem L;
  L.stage_cap = kc + kStageExtra;
  size_t o = 0;
  L.stage = o; o += static_cast<size_t>(L.stage_cap) * 8;
  L.sel = o; o += static_cast<size_t>(kc) * 8;
  L.fk = o; o += static_cast<size_t>(kc) * 8;
  L.qrow = o; o += static_cast<size_t>(ld) * 4;
  L.pos = o; o += static_cast<size_t>(kc) * 4;
  L.hist = o; o += kBins * 4;
  L.offs = o; o += (static_cast<size_t>(nsplit) + 1) * 4;
  o = (o + 15) / 16 * 16;
  L.scratch = o; o += 32 * 4;
  L.total = o;
  return L;
}

// Visit every candidate key of this query: warp w walks lists w, w+8, ...; lanes stride over a
// list (coalesced 256-byte reads), kKeyBatch loads in flight per lane before any is consumed
// (the visitor has shared-memory side effects the compiler will not hoist loads across).
constexpr int kKeyBatch = 8;
template <typename F>
__device__ __forceinline__ void for_each_key(const uint64_t* __restrict__ lists, const int32_t* offs, int nsplit,
                                             int cap, int warp, int lane, F&& f) {
  for (int s = warp; s < nsplit; s += kSelWarps) {
    const int c = offs[s + 1] - offs[s];
    const uint64_t* src = lists + static_cast<int64_t>(s) * cap;
    for (int base = 0; base < c; base += kWarp * kKeyBatch) {
      uint64_t kk[kKeyBatch];
#pragma unroll
      for (int u = 0; u < kKeyBatch; ++u) {
        const int i = base + u * kWarp + lane;
        kk[u] = (i < c) ? __ldcg(src + i) : 0ull;  // lists hold RAW keys; index part is never 0
      }
#pragma unroll
      for (int u = 0; u < kKeyBatch; ++u)
        if (kk[u] != 0ull) f(raw2key(kk[u]));
    }
  }
}

__global__ void __launch_bounds__(kSelThreads)
select_rescore_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int ld, int64_t nq,
                      int64_t ng, int k, int64_t idx_offset, int nsplit, int cap, int kc,
                      const int32_t* __restrict__ counts, const uint64_t* __restrict__ cand,
                      const float* __restrict__ thr_out, const float* __restrict__ q_delta,
                      float g_delta_max, float eps_acc, float* __restrict__ out_sim,
                      int64_t* __restrict__ out_idx, int32_t* __restrict__ uncert_list,
                      int32_t* __restrict__ uncert_count, SelSmem L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw + L.stage);
  uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw + L.sel);
  uint64_t* fk = reinterpret_cast<uint64_t*>(smem_raw + L.fk);
  int32_t* pos = reinterpret_cast<int32_t*>(smem_raw + L.pos);
  float* qrow = reinterpret_cast<float*>(smem_raw + L.qrow);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + L.hist);
  int32_t* offs = reinterpret_cast<int32_t*>(smem_raw + L.offs);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw + L.scratch);  // [0..3] block_select, [4..] ours

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q = blockIdx.x;
  const int ld4 = ld >> 2;
  const uint64_t* lists = cand + q * nsplit * static_cast<int64_t>(cap);

  // ---- stage the fp32 query row; prefix-sum the list lengths; largest list threshold --------
  {
    const float4* src = reinterpret_cast<const float4*>(q32 + q * static_cast<int64_t>(ld));
    for (int c = tid; c < ld4; c += kSelThreads) reinterpret_cast<float4*>(qrow)[c] = __ldg(src + c);
  }
  if (warp == 0) {
    int run = 0;
    float tmax = -INFINITY, tmin = INFINITY;
    for (int base = 0; base < nsplit; base += kWarp) {
      const int s = base + lane;
      const int c = (s < nsplit) ? counts[q * nsplit + s] : 0;
      if (s < nsplit) {
        const float t = thr_out[q * nsplit + s];
        tmax = fmaxf(tmax, t);
        tmin = fminf(tmin, t);
      }
      int incl = c;
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      if (s < nsplit) offs[s + 1] = run + incl;
      run += __shfl_sync(kFull, incl, kWarp - 1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tmax = fmaxf(tmax, __shfl_xor_sync(kFull, tmax, o));
      tmin = fminf(tmin, __shfl_xor_sync(kFull, tmin, o));
    }
    if (lane == 0) {
      offs[0] = 0;
      scratch[8] = __float_as_uint(tmax);
      scratch[9] = 0;   // staged count
      scratch[13] = 0;  // keys above the assumed score range
      scratch[14] = __float_as_uint(tmin);
    }
  }
  __syncthreads();
  const int total = offs[nsplit];
  // rows that are in no list score <= the threshold their list ended with (bf16 contraction)
  float tprime = __uint_as_float(scratch[8]);
  const bool all_in = (static_cast<int64_t>(total) == ng);

  // ---- gather the candidates that can be among the kc best (by bf16 score) into `stage` ------
  int nstage;
  if (total <= L.stage_cap) {
    for_each_key(lists, offs, nsplit, cap, warp, lane, [&](uint64_t key) { stage[atomicAdd(&scratch[9], 1u)] = key; });
    __syncthreads();
    nstage = total;
  } else {
    // streaming selection on the ordered similarity bits: kBins linear bins over [lo, hi],
    // narrowed to the bin that holds the kc-th best until that bin fits the staging buffer.
    // Every key exceeds the smallest list threshold, and unit bf16 rows score <= ~1.008, which
    // makes the first pass effective; the full range is the fallback if either bound is moot.
    const float tmin = __uint_as_float(scratch[14]);
    uint32_t lo = (tmin > -INFINITY) ? f2ord(tmin) : 0u;
    uint32_t hi = (tmin > -INFINITY) ? f2ord(1.01f) : 0xFFFFFFFFu;
    bool check_range = (hi != 0xFFFFFFFFu);
    uint32_t above = 0u, cut = 0u;
    for (int pass = 0; pass < 6; ++pass) {
      const uint32_t span = hi - lo;  // inclusive range size - 1
      int shift = 0;
      while ((span >> shift) >= static_cast<uint32_t>(kBins)) ++shift;
      for (int b = tid; b < kBins; b += kSelThreads) hist[b] = 0;
      __syncthreads();
      for_each_key(lists, offs, nsplit, cap, warp, lane, [&](uint64_t key) {
        const uint32_t o = static_cast<uint32_t>(key >> 32);
        if (o >= lo && o <= hi) atomicAdd(&hist[(o - lo) >> shift], 1u);
        else if (check_range) scratch[13] = 1u;
      });
      __syncthreads();
      if (check_range) {
        check_range = false;
        if (scratch[13] != 0u) {  // a key outside the assumed range: start over with the full range
          lo = 0u;
          hi = 0xFFFFFFFFu;
          __syncthreads();
          continue;
        }
      }
      if (warp == 0) {  // scan the bins from the top: lane l owns the 64 bins below kBins - 64*l
        constexpr int per = kBins / kWarp;
        uint32_t s = 0;
        for (int j = 0; j < per; ++j) s += hist[kBins - 1 - (per * lane + j)];
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
          const uint32_t t = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += t;
        }
        const uint32_t need = static_cast<uint32_t>(kc) - above;  // >= 1
        const uint32_t excl = incl - s;
        if (excl < need && need <= incl) {
          uint32_t run = excl;
          for (int j = 0; j < per; ++j) {
            const int b = kBins - 1 - (per * lane + j);
            if (run + hist[b] >= need) {
              scratch[10] = static_cast<uint32_t>(b);
              scratch[11] = above + run;  // keys strictly above bin b
              scratch[12] = hist[b];
              break;
            }
            run += hist[b];
          }
        }
      }
      __syncthreads();
      const uint32_t b = scratch[10], n_above = scratch[11], n_in = scratch[12];
      cut = lo + (b << shift);
      if (n_above + n_in <= static_cast<uint32_t>(L.stage_cap) || shift == 0) break;
      above = n_above;
      lo = cut;
      hi = cut + ((1u << shift) - 1u);
      __syncthreads();
    }
    // collect everything at or above the cut (exact ties beyond the staging room are dropped:
    // they score == the kc-th best, which the certification bound below covers)
    for_each_key(lists, offs, nsplit, cap, warp, lane, [&](uint64_t key) {
      if (static_cast<uint32_t>(key >> 32) >= cut) {
        const uint32_t at = atomicAdd(&scratch[9], 1u);
        if (at < static_cast<uint32_t>(L.stage_cap)) stage[at] = key;
      }
    });
    __syncthreads();
    nstage = min(static_cast<int>(scratch[9]), L.stage_cap);
  }

  // ---- keep the kc best by bf16 score --------------------------------------------------------
  int ncand;
  if (nstage > kc) {
    const uint64_t thr_c = block_select(stage, nstage, kc, sel, hist, scratch);
    tprime = fmaxf(tprime, key_sim(thr_c));
    ncand = kc;
  } else {
    for (int i = tid; i < nstage; i += kSelThreads) sel[i] = stage[i];
    ncand = nstage;
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
  for (int t = warp; t < n1; t += kSelWarps) {
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
  for (int t = warp; t < n2; t += kSelWarps) {
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
  const SelSmem L = sel_smem_layout(plan->nlists, plan->kc, ld);
  HCIR_REQUIRE(L.total <= 220 * 1024, "select_rescore: kc=%d ld=%d needs %zu B of shared memory", plan->kc, ld,
               L.total);
  HCIR_CUDA_TRY(cudaFuncSetAttribute(select_rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(L.total)));
  const char* ws = static_cast<const char*>(workspace);
  const int32_t* counts = reinterpret_cast<const int32_t*>(ws + plan->counts_off);
  const uint64_t* cand = reinterpret_cast<const uint64_t*>(ws + plan->keys_off);
  const float* thr_out = reinterpret_cast<const float*>(ws + plan->thr_out_off);
  select_rescore_kernel<<<static_cast<unsigned>(nq), kSelThreads, L.total, static_cast<cudaStream_t>(stream)>>>(
      q_f32, g_f32, ld, nq, ng, k, idx_offset, plan->nlists, plan->cap, plan->kc, counts, cand, thr_out, q_delta,
      g_delta_max, eps_acc, out_sim, out_idx, uncert_list, uncert_count, L);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}
