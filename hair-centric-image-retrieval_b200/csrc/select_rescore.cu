// K3: per-query candidate selection, adaptive fp32 re-score, exact sort, certification -- and the
// TAIL of the step fused behind it: neighbour-label gather, kNN vote, and the store of the query's
// results straight into every rank's peer region over NVLink (multi-GPU), with one arrival signal per
// rank and step.  One CTA per query.
//
// Bandwidth / latency-bound: ONE streaming read of the query's candidate lists (the sample pass
// supplies a staging hint thr_hi[q] that ~4*kc candidates exceed, so the kc best can be picked from
// a small shared-memory stage without a second pass; a histogram-select fallback covers the cases
// where the hint is off) plus ~ (k + 6 + near-boundary candidates) fp32 gallery rows of 4*ld bytes.
//
// Measured shape of the round-1 kernel (ncu --set full, C2, profiles/): 29 % of the stall samples in
// the list read + staging loop (one shared-memory atomic per staged key: short-scoreboard), 25 % in
// the row gathers, 21 % at barriers.  Hence: (a) lists are walked as a queue of 32-key chunks so a
// lane always has kKeyBatch independent loads in flight whatever the list lengths are, (b) staging
// reserves room with ONE atomic per warp and batch (shuffle prefix over per-lane hit counts),
// (c) few queries (the streaming regime) get 1024-thread CTAs: 32 warps share the list walk, one
// fp32 row per warp in the re-score, one warp per key in the rank sort.
#include <math.h>

#include "hcir_common.cuh"

namespace hcir {

constexpr int kRound1Slack = 6;
#ifndef HCIR_K3_PREFETCH_AHEAD
#define HCIR_K3_PREFETCH_AHEAD 0
#endif
constexpr int kPrefetchAhead = HCIR_K3_PREFETCH_AHEAD;  // gather steps whose rows are prefetched into L2 (0 = off, the default)
constexpr int kBins = 2048;  // histogram bins of the fallback streaming selection

struct SelSmem {  // offsets (bytes) into dynamic shared memory
  size_t stage, sel, fk, clab, qrow, hist, offs, scratch, total;
  int stage_cap;
};

static SelSmem sel_smem_layout(int nlists, int kc, int ld) {
  SelSmem L;
  // the lists hold ~6.5 x kc keys (optimistic threshold): stage them all -- 10 x kc slots, fewer when a wide
  // kc (the 4x kc of a completion pass) would not leave room for the rest (the kernel falls back to its
  // histogram selection when the stage overflows)
  const size_t fixed = static_cast<size_t>(kc) * 16 + (static_cast<size_t>(kc) * 4 + 15) / 16 * 16 +
                       static_cast<size_t>(ld) * 4 + kBins * 4 + (static_cast<size_t>(nlists) + 1) * 4 + 16 + 32 * 4;
  size_t cap = 10 * static_cast<size_t>(kc) > 1024 ? 10 * static_cast<size_t>(kc) : 1024;
  const size_t budget = 212 * 1024;
  if (fixed + cap * 8 > budget && fixed < budget) {
    const size_t fit = (budget - fixed) / 8;
    const size_t floor_cap = 2 * static_cast<size_t>(kc) > 1024 ? 2 * static_cast<size_t>(kc) : 1024;
    cap = fit > floor_cap ? fit : floor_cap;
  }
  L.stage_cap = static_cast<int>(cap);
  size_t o = 0;
  L.stage = o; o += static_cast<size_t>(L.stage_cap) * 8;
  L.sel = o; o += static_cast<size_t>(kc) * 8;
  L.fk = o; o += static_cast<size_t>(kc) * 8;
  L.clab = o; o += (static_cast<size_t>(kc) * 4 + 15) / 16 * 16;  // keeps qrow (float4 accesses) 16-byte aligned
  L.qrow = o; o += static_cast<size_t>(ld) * 4;
  L.hist = o; o += kBins * 4;
  L.offs = o; o += (static_cast<size_t>(nlists) + 1) * 4;
  o = (o + 15) / 16 * 16;
  L.scratch = o; o += 32 * 4;
  L.total = o;
  return L;
}

// device-side form of hcir_tail_t
struct TailParams {
  const int32_t* labels;
  int64_t n_labels;
  int num_classes;
  float T;
  const int64_t* classes;
  int64_t* pred;
  int32_t* out_lab;
  int world, rank, payload;
  size_t slot_stride;
  char* region[kPeerMax];
  const int64_t* step;
  // completion launches (second pass over the uncertified queries of a step, device-driven):
  const int32_t* qmap;    // launch row r answers ORIGINAL query qmap[r]: every output goes to that row
  const int32_t* active;  // only rows r < *active are live; the others go straight to the done count
  int commit_certified_only;  // outputs are written only if this launch certifies the query
  int no_signal;          // peer rows are stored but the arrival signal is left to a later launch
  int64_t out_rows;       // rows of the output arrays / of the packed peer block (0 = nq)
};

struct K3Args {
  const float* q32;
  const float* g32;
  int ld;
  int64_t nq, ng;
  int k;
  int64_t idx_offset;
  int nlists, cap, kc;
  const int32_t* counts;
  const uint64_t* cand;
  const float* thr_out;
  const float* thr_hi;
  const float* q_delta;
  float g_delta_max, eps_acc;
  float* out_sim;
  int64_t* out_idx;
  int32_t* uncert_list;
  int32_t* state;
  SelSmem L;
  TailParams tail;
};

// Visit every candidate key of this query in batches: the lists of warp w (w, w+W, ...) form a queue
// of 32-key chunks; every round takes the next kKeyBatch chunks, issues all their loads, then hands
// the batch to the visitor (kk[u] == 0: no key; RAW keys, their index part is never 0).  With the
// optimistic main-pass threshold a query's lists hold 4-7 x kc keys in all -- a few dozen per list, a
// handful in the streaming regime -- so a warp that walked one list per round trip (the round-1 walk,
// 5 % faster on the long lists of the deterministic threshold) would pay one DRAM latency per list.
constexpr int kKeyBatch = 8;
template <int kSelThreads, typename F>
__device__ __forceinline__ void for_each_batch(const uint64_t* __restrict__ lists, const int32_t* cnts, int nlists,
                                               int cap, int warp, int lane, F&& f) {
  constexpr int kSelWarps = kSelThreads / kWarp;
  int s = warp, base = 0;
  int c = (s < nlists) ? cnts[s] : 0;
  for (;;) {
    uint64_t kk[kKeyBatch];
    bool any = false;
#pragma unroll
    for (int u = 0; u < kKeyBatch; ++u) {
      while (s < nlists && base >= c) {  // warp-uniform: next non-empty list of this warp
        s += kSelWarps;
        base = 0;
        c = (s < nlists) ? cnts[s] : 0;
      }
      kk[u] = 0ull;
      if (s < nlists) {
        const int i = base + lane;
        if (i < c) kk[u] = __ldcg(lists + static_cast<int64_t>(s) * cap + i);
        base += kWarp;
        any = true;
      }
    }
    if (!any) break;
    f(kk);
  }
}

// Append the keys of a batch whose ordered score exceeds (or reaches) `lim` to stage[]: every lane
// counts its hits, a shuffle prefix gives its offset, ONE shared-memory atomic per warp reserves
// the room.  Keys beyond stage_cap are dropped (the caller sees the count).
template <bool kInclusive>
__device__ __forceinline__ void stage_batch(const uint64_t (&kk)[kKeyBatch], uint32_t lim, uint64_t* stage,
                                            int stage_cap, uint32_t* counter, int lane) {
  uint32_t hits = 0u;
#pragma unroll
  for (int u = 0; u < kKeyBatch; ++u) {
    if (kk[u] != 0ull) {
      const uint32_t o = f2ord(__uint_as_float(static_cast<uint32_t>(kk[u] >> 32)));
      if (kInclusive ? (o >= lim) : (o > lim)) hits |= 1u << u;
    }
  }
  const int n = __popc(hits);
  int incl = n;
#pragma unroll
  for (int o = 1; o < kWarp; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(kFull, incl, kWarp - 1);
  if (total == 0) return;  // warp-uniform
  uint32_t base = 0u;
  if (lane == kWarp - 1) base = atomicAdd(counter, static_cast<uint32_t>(total));
  base = __shfl_sync(kFull, base, kWarp - 1);
  uint32_t at = base + static_cast<uint32_t>(incl - n);
#pragma unroll
  for (int u = 0; u < kKeyBatch; ++u) {
    if (hits & (1u << u)) {
      if (at < static_cast<uint32_t>(stage_cap)) stage[at] = raw2key(kk[u]);
      ++at;
      HCIR_DEV_CHECK(key_idx(raw2key(kk[u])) < 0x7FFFFFFFu);   // a listed key names a real gallery row
    }
  }
}

// rank (number of strictly greater keys) of every key[0..n) -> dst[rank] = key for rank < keep.
// Keys are unique, so ranks are a permutation.  The caller syncs.  kWarpPerKey: one warp per key,
// lanes split the scan (wide CTAs, short arrays), else one thread per key.
template <int kSelThreads, bool kWarpPerKey>
__device__ __forceinline__ void rank_scatter(const uint64_t* keys, int n, uint64_t* dst, int keep, int tid) {
  if constexpr (kWarpPerKey) {
    const int lane = tid & 31;
    for (int j = tid >> 5; j < n; j += kSelThreads / kWarp) {
      const uint64_t mine = keys[j];
      int rank = 0;
      for (int i = lane; i < n; i += kWarp) rank += (keys[i] > mine) ? 1 : 0;
      rank = __reduce_add_sync(kFull, rank);
      if (lane == 0 && rank < keep) dst[rank] = mine;
    }
  } else {
    for (int j = tid; j < n; j += kSelThreads) {
      const uint64_t mine = keys[j];
      int rank = 0;
      for (int i = 0; i < n; ++i) rank += (keys[i] > mine) ? 1 : 0;
      if (rank < keep) dst[rank] = mine;
    }
  }
}

#ifndef HCIR_K3_THREADS_PER_SM
#define HCIR_K3_THREADS_PER_SM 1024  // resident K3 threads per SM the register budget is sized for (A/B: 1280)
#endif
// Everything a query's CTA does up to and including its stores into the peers (the kernel below adds the
// launch-wide done count / arrival signal, which every CTA must reach whether its row is live or not).
template <int kSelThreads>
__device__ __forceinline__ void select_rescore_body(const K3Args& a, const int64_t q) {
  const float* __restrict__ q32 = a.q32;
  const float* __restrict__ g32 = a.g32;
  const int ld = a.ld;
  const int64_t nq = a.nq, ng = a.ng;
  const int k = a.k;
  const int64_t idx_offset = a.idx_offset;
  const int nlists = a.nlists, cap = a.cap, kc = a.kc;
  const int32_t* __restrict__ counts = a.counts;
  const uint64_t* __restrict__ cand = a.cand;
  const float* __restrict__ thr_out = a.thr_out;
  const float* __restrict__ thr_hi = a.thr_hi;
  const float* __restrict__ q_delta = a.q_delta;
  const float g_delta_max = a.g_delta_max, eps_acc = a.eps_acc;
  float* __restrict__ out_sim = a.out_sim;
  int64_t* __restrict__ out_idx = a.out_idx;
  int32_t* __restrict__ uncert_list = a.uncert_list;
  int32_t* __restrict__ state = a.state;
  const SelSmem& L = a.L;
  const TailParams& tail = a.tail;
  constexpr int kSelWarps = kSelThreads / kWarp;
  constexpr bool kWide = (kSelThreads >= 1024);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* stage = reinterpret_cast<uint64_t*>(smem_raw + L.stage);
  uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw + L.sel);   // kc best by bf16 score, DESCENDING
  uint64_t* fk = reinterpret_cast<uint64_t*>(smem_raw + L.fk);     // fp32 keys of the re-scored prefix of sel
  int32_t* clab = reinterpret_cast<int32_t*>(smem_raw + L.clab);   // class index of every re-scored candidate
  float* qrow = reinterpret_cast<float*>(smem_raw + L.qrow);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + L.hist);
  int32_t* offs = reinterpret_cast<int32_t*>(smem_raw + L.offs);
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw + L.scratch);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ld4 = ld >> 2;
  const uint64_t* lists = cand + q * nlists * static_cast<int64_t>(cap);

  // ---- start: issue the query-row loads (consumed after the key pass); every warp fetches the
  // lengths / thresholds of ITS lists (w, w+W, ...) and goes straight to reading keys -- there is
  // no serial prefix phase; totals are combined with shared-memory atomics -----------------------
  constexpr int kQv = 2;  // float4 per thread held in registers: covers ld <= 2048 (the rest is loaded late)
  float4 qv[kQv];
  {
    const float4* src = reinterpret_cast<const float4*>(q32 + q * static_cast<int64_t>(ld));
#pragma unroll
    for (int u = 0; u < kQv; ++u) {
      const int c = tid + u * kSelThreads;
      if (c < ld4) qv[u] = __ldg(src + c);
    }
  }
  if (tid == 0) {
    scratch[5] = 0;
    scratch[6] = 0;
    scratch[7] = 0;
    scratch[9] = 0;             // staged count
    scratch[10] = 0;            // fallback bin scan result: bin 0 / nothing above / nothing in it = "take every
    scratch[11] = 0;            //   key in range" -- what remains in force when fewer than kc keys exist and no
    scratch[12] = 0;            //   bin reaches the kc-th rank
    scratch[13] = 0;            // keys outside the assumed score range
    scratch[14] = 0xFFFFFFFFu;  // smallest list threshold (ordered bits)
    scratch[18] = 0;            // total candidates
    scratch[19] = 0;            // largest list threshold (ordered bits)
  }
  __syncthreads();
  {
    int tot = 0;
    uint32_t omax = 0u, omin = 0xFFFFFFFFu;
    for (int s = warp + kSelWarps * lane; s < nlists; s += kSelWarps * kWarp) {
      const int c = counts[q * nlists + s];
      offs[s] = c;
      tot += c;
      const uint32_t o = f2ord(thr_out[q * nlists + s]);
      omax = max(omax, o);
      omin = min(omin, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot += __shfl_xor_sync(kFull, tot, o);
      omax = max(omax, __shfl_xor_sync(kFull, omax, o));
      omin = min(omin, __shfl_xor_sync(kFull, omin, o));
    }
    if (lane == 0) {
      atomicAdd(&scratch[18], static_cast<uint32_t>(tot));
      atomicMax(&scratch[19], omax);
      atomicMin(&scratch[14], omin);
    }
    __syncwarp();  // this warp's offs[] entries are visible to its own lanes
  }

  // ---- gather the candidates that can be among the kc best (by bf16 score) into `stage` ------
  // fast path: one pass, keep what exceeds the staging hint (or everything, if it all fits)
  // (the hint is "stage everything" when there was no sample pass)
  const uint32_t hint = thr_hi ? f2ord(thr_hi[q]) : 0u;
  int nstage = 0;
  bool staged = false;
  {
    auto visit = [&](const uint64_t (&kk)[kKeyBatch]) {
      stage_batch<false>(kk, hint, stage, L.stage_cap, &scratch[9], lane);
    };
    for_each_batch<kSelThreads>(lists, offs, nlists, cap, warp, lane, visit);
    // the query row goes to shared memory now that the key loads have been issued
#pragma unroll
    for (int u = 0; u < kQv; ++u) {
      const int c = tid + u * kSelThreads;
      if (c < ld4) reinterpret_cast<float4*>(qrow)[c] = qv[u];
    }
    for (int c = tid + kQv * kSelThreads; c < ld4; c += kSelThreads)
      reinterpret_cast<float4*>(qrow)[c] = __ldg(reinterpret_cast<const float4*>(q32 + q * static_cast<int64_t>(ld)) + c);
    __syncthreads();
    const int got = static_cast<int>(scratch[9]);
    // usable iff nothing was dropped and the kc best are all in the stage
    staged = (got <= L.stage_cap) && (got >= kc || got == static_cast<int>(scratch[18]));
    nstage = got;
    __syncthreads();
  }
  const int total = static_cast<int>(scratch[18]);
  // rows that are in no list score <= the threshold their list ended with (bf16 contraction)
  // every gallery row is in some list (no threshold dropped anything): only the kc cut below
  // separates rows from the re-score
  const bool all_in = (static_cast<int64_t>(total) == ng);
  float tprime = all_in ? -INFINITY : ord2f(scratch[19]);
  if (!staged) {
    // fallback: streaming selection on the ordered similarity bits -- kBins linear bins over
    // [lo, hi], narrowed to the bin that holds the kc-th best until that bin fits the stage.
    // Every key exceeds the smallest list threshold and unit bf16 rows score <= ~1.008, which
    // makes the first pass effective; the full range is used if either bound is moot.
    if (tid == 0) scratch[9] = 0;
    const float tmin = ord2f(scratch[14]);
    uint32_t lo = (tmin > -INFINITY) ? scratch[14] : 0u;
    uint32_t hi = (tmin > -INFINITY) ? f2ord(1.01f) : 0xFFFFFFFFu;
    bool check_range = (hi != 0xFFFFFFFFu);
    uint32_t above = 0u, cut = 0u;
    for (int pass = 0; pass < 6; ++pass) {
      const uint32_t span = hi - lo;  // inclusive range size - 1
      int shift = 0;
      while ((span >> shift) >= static_cast<uint32_t>(kBins)) ++shift;
      for (int b = tid; b < kBins; b += kSelThreads) hist[b] = 0;
      __syncthreads();
      for_each_batch<kSelThreads>(lists, offs, nlists, cap, warp, lane, [&](const uint64_t (&kk)[kKeyBatch]) {
#pragma unroll
        for (int u = 0; u < kKeyBatch; ++u) {
          if (kk[u] == 0ull) continue;
          const uint32_t o = f2ord(__uint_as_float(static_cast<uint32_t>(kk[u] >> 32)));
          if (o >= lo && o <= hi) {
            HCIR_DEV_CHECK(((o - lo) >> shift) < static_cast<uint32_t>(kBins));
            atomicAdd(&hist[(o - lo) >> shift], 1u);
          } else if (check_range) {
            scratch[13] = 1u;
          }
        }
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
    for_each_batch<kSelThreads>(lists, offs, nlists, cap, warp, lane, [&](const uint64_t (&kk)[kKeyBatch]) {
      stage_batch<true>(kk, cut, stage, L.stage_cap, &scratch[9], lane);
    });
    __syncthreads();
    nstage = min(static_cast<int>(scratch[9]), L.stage_cap);
  }

  // ---- the kc best by bf16 score, in descending order -----------------------------------------
  // The stage holds a superset of the kc best.  Rank counting is O(n^2), so a stage much larger
  // than kc is first cut down with one shared-memory histogram over its own score range.
  const int ncand = nstage < kc ? nstage : kc;
  const uint64_t* pool = stage;
  int npool = nstage;
  // (measured, r2h: ranking the ~700 staged keys of the streaming regime directly -- one warp per key -- instead of
  //  cutting them down first doubles K3 there, 31 -> 62 us: n^2 / 32 warp steps is 11 us of issue slots)
  if (nstage > kc + 128) {
    // wide CTAs (streaming regime) cut with 512 bins: warp 0 scans the bins alone while 31 warps wait at the
    // barrier (14 % of the kernel's stall samples with 2048 bins), and ~700 staged keys need no finer cut
    constexpr int kCutBins = kWide ? 512 : kBins;
    uint64_t* compact = reinterpret_cast<uint64_t*>(hist);  // the histogram's memory, reused after the scan
    constexpr int kCompactCap = kBins * 4 / 8;
    if (tid == 0) { scratch[15] = 0xFFFFFFFFu; scratch[16] = 0u; scratch[17] = 0u; }
    __syncthreads();
    {  // score range of the stage
      uint32_t mn = 0xFFFFFFFFu, mx = 0u;
      for (int i = tid; i < nstage; i += kSelThreads) {
        const uint32_t o = static_cast<uint32_t>(stage[i] >> 32);
        mn = min(mn, o);
        mx = max(mx, o);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(kFull, mn, o));
        mx = max(mx, __shfl_xor_sync(kFull, mx, o));
      }
      if (lane == 0) { atomicMin(&scratch[15], mn); atomicMax(&scratch[16], mx); }
    }
    for (int b = tid; b < kCutBins; b += kSelThreads) hist[b] = 0;
    __syncthreads();
    const uint32_t lo = scratch[15], span = scratch[16] - lo;
    int shift = 0;
    while ((span >> shift) >= static_cast<uint32_t>(kCutBins)) ++shift;
    for (int i = tid; i < nstage; i += kSelThreads) {
      HCIR_DEV_CHECK(((static_cast<uint32_t>(stage[i] >> 32) - lo) >> shift) < static_cast<uint32_t>(kCutBins));
      atomicAdd(&hist[(static_cast<uint32_t>(stage[i] >> 32) - lo) >> shift], 1u);
    }
    __syncthreads();
    if (warp == 0) {  // bin that holds the kc-th best, scanning from the top
      constexpr int per = kCutBins / kWarp;
      uint32_t sum = 0;
      for (int j = 0; j < per; ++j) sum += hist[kCutBins - 1 - (per * lane + j)];
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t need = static_cast<uint32_t>(kc), excl = incl - sum;
      if (excl < need && need <= incl) {
        uint32_t run = excl;
        for (int j = 0; j < per; ++j) {
          const int b = kCutBins - 1 - (per * lane + j);
          if (run + hist[b] >= need) {
            scratch[10] = static_cast<uint32_t>(b);
            scratch[11] = run + hist[b];  // keys at or above the cut
            break;
          }
          run += hist[b];
        }
      }
    }
    __syncthreads();
    const uint32_t cut = lo + (scratch[10] << shift);
    const int keep = static_cast<int>(scratch[11]);
    __syncthreads();  // everyone has read the scan result: the histogram memory may be reused
    if (keep <= kCompactCap) {
      // one atomic per warp and round of 32 keys (ballot prefix), not one per kept key
      for (int i0 = warp * kWarp; i0 < nstage; i0 += kSelThreads) {
        const int i = i0 + lane;
        const uint64_t key = (i < nstage) ? stage[i] : 0ull;
        const bool hit = (i < nstage) && (static_cast<uint32_t>(key >> 32) >= cut);
        const uint32_t b = __ballot_sync(kFull, hit);
        if (b == 0u) continue;
        uint32_t base = 0u;
        if (lane == 0) base = atomicAdd(&scratch[17], static_cast<uint32_t>(__popc(b)));
        base = __shfl_sync(kFull, base, 0);
        HCIR_DEV_CHECK(!hit || base + __popc(b & ((1u << lane) - 1u)) < static_cast<uint32_t>(kCompactCap));
        if (hit) compact[base + __popc(b & ((1u << lane) - 1u))] = key;
      }
      __syncthreads();
      pool = compact;
      npool = keep;
    }
  }
  rank_scatter<kSelThreads, kWide>(pool, npool, sel, kc, tid);
  __syncthreads();
  // Listed keys that did not make it into `sel` are bounded by sel's worst score -- whether they were
  // staged and lost the kc cut (nstage > kc) or never staged at all (nstage == kc exactly: they sit at
  // or below the staging hint / the fallback cut, which every staged key exceeds).  Only when every
  // listed key is in `sel` (total <= kc) do the list thresholds alone bound the excluded rows.
  if (nstage >= kc && total > kc) tprime = fmaxf(tprime, key_sim(sel[kc - 1]));

  const float dq = q_delta ? q_delta[q] : 0.0f;
  const float eps = g_delta_max * (1.0f + dq) + dq * (1.0f + 1e-6f) + eps_acc;

  // fp32 re-score of sel[a..b): one row per warp while the warps suffice, else two rows per warp step.
  // The candidates' class indices are fetched alongside (the load hides under the row gather), so the
  // tail needs no dependent label round trip once the final ranks are known.
  auto label_of = [&](uint32_t r) -> int32_t {
    return (tail.labels != nullptr && static_cast<int64_t>(r) < tail.n_labels) ? __ldg(tail.labels + r) : -1;
  };
  auto rescore = [&](int a, int b) {
    const float4* q4 = reinterpret_cast<const float4*>(qrow);
    if (b - a <= kSelWarps) {
      const int j = a + warp;
      if (j < b) {
        const uint32_t r0 = key_idx(sel[j]);
        const int32_t l0 = (lane == 0) ? label_of(r0) : 0;
        const float s0 =
            canonical_dot(q4, reinterpret_cast<const float4*>(g32 + static_cast<int64_t>(r0) * ld), ld4, lane);
        if (lane == 0) {
          fk[j] = make_key(s0, r0);
          clab[j] = l0;
        }
      }
      return;
    }
    // EXPERIMENT, off by default (-DHCIR_K3_PREFETCH_AHEAD=n): request the rows of the step n steps ahead into
    // L2 with prefetch instructions (no registers, no scoreboard) so that later steps hit L2.  Measured (r2k):
    // it makes K3 SLOWER -- C3 0.45 -> 0.52 / 0.57 / 0.58 ms for n = 2 / 4 / 8, C5 shard 8.9 -> 13.7 ms -- i.e.
    // the row gathers are not short of requests in flight; ~4.3 TB/s is what random 3-8 KB rows get here.
    const int stride = 2 * kSelWarps;
    const int lines = ld4 >> 3;  // 128-byte lines per row
    auto prefetch_pair = [&](int jp) {
      if (jp >= b) return;
      const char* p0 = reinterpret_cast<const char*>(g32 + static_cast<int64_t>(key_idx(sel[jp])) * ld);
      for (int l = lane; l < lines; l += kWarp) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + l * 128));
      if (jp + 1 < b) {
        const char* p1 = reinterpret_cast<const char*>(g32 + static_cast<int64_t>(key_idx(sel[jp + 1])) * ld);
        for (int l = lane; l < lines; l += kWarp) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + l * 128));
      }
    };
#pragma unroll
    for (int sft = 1; sft < kPrefetchAhead; ++sft) prefetch_pair(a + 2 * warp + sft * stride);
    for (int j = a + 2 * warp; j < b; j += stride) {
      if constexpr (kPrefetchAhead > 0) prefetch_pair(j + kPrefetchAhead * stride);
      const uint32_t r0 = key_idx(sel[j]);
      const bool two = (j + 1 < b);
      const uint32_t r1 = two ? key_idx(sel[j + 1]) : r0;
      const int32_t l0 = (lane == 0) ? label_of(r0) : 0, l1 = (lane == 0 && two) ? label_of(r1) : 0;
      float s0, s1;
      canonical_dot2(q4, reinterpret_cast<const float4*>(g32 + static_cast<int64_t>(r0) * ld),
                     reinterpret_cast<const float4*>(g32 + static_cast<int64_t>(r1) * ld), ld4, lane, s0, s1);
      if (lane == 0) {
        fk[j] = make_key(s0, r0);
        clab[j] = l0;
        if (two) {
          fk[j + 1] = make_key(s1, r1);
          clab[j + 1] = l1;
        }
      }
    }
  };

  // ---- round 1: the k + slack best bf16 candidates ----------------------------------------
  const int n1 = (ncand < k + kRound1Slack) ? ncand : k + kRound1Slack;
  rescore(0, n1);
  __syncthreads();
  // k-th best fp32 score so far (a lower bound of the final k-th best)
  for (int t = tid; t < n1; t += kSelThreads) {
    const uint64_t mine = fk[t];
    int rank = 0;
    for (int i = 0; i < n1; ++i) rank += (fk[i] > mine) ? 1 : 0;
    if (rank == k - 1) scratch[5] = __float_as_uint(key_sim(mine));
  }
  __syncthreads();
  const float sk1 = (n1 >= k) ? __uint_as_float(scratch[5]) : -INFINITY;

  // ---- round 2: every other candidate whose bf16 score could still reach the top-k --------
  // sel is sorted by bf16 score, so these form a prefix [n1, n1 + n2)
  for (int j = n1 + tid; j < ncand; j += kSelThreads) {
    if (key_sim(sel[j]) + eps >= sk1) atomicAdd(&scratch[6], 1u);
  }
  __syncthreads();
  const int nr = n1 + static_cast<int>(scratch[6]);
  HCIR_DEV_CHECK(nr <= ncand && ncand <= kc && nstage <= L.stage_cap && 8 * k <= 8 * L.stage_cap);
  rescore(n1, nr);
  __syncthreads();

  // ---- exact order of the re-scored set, emit top-k ---------------------------------------
  // (sel is dead from here on: its memory receives the final keys in rank order for the tail)
  // and the stage -- dead since the rank sort -- the neighbour class indices and similarities in rank order)
  uint64_t* res = sel;
  int32_t* lab = reinterpret_cast<int32_t*>(stage);    // [k]
  float* rsim = reinterpret_cast<float*>(stage) + k;  // [k]   (k <= kc <= stage_cap / 8)
  for (int t = tid; t < nr; t += kSelThreads) {
    const uint64_t mine = fk[t];
    int rank = 0;
    for (int i = 0; i < nr; ++i) rank += (fk[i] > mine) ? 1 : 0;
    if (rank < k) {
      const float s = key_sim(mine);
      res[rank] = mine;
      rsim[rank] = s;
      lab[rank] = clab[t];
      if (rank == k - 1) scratch[7] = __float_as_uint(s);
    }
  }
  // fewer than k candidates (the query is uncertified): the missing ranks read (-inf, -1), never stale memory
  for (int j = nr + tid; j < k; j += kSelThreads) {
    res[j] = 0ull;
    rsim[j] = -INFINITY;
    lab[j] = -1;
  }
  __syncthreads();
  const float sk = (nr >= k) ? __uint_as_float(scratch[7]) : -INFINITY;
  // rows outside `sel` score <= tprime in bf16 (list thresholds, kc cut), hence <= tprime + eps in fp32
  const bool certified = (nr >= k) && (tprime + eps < sk);
  // completion launches answer ORIGINAL query qmap[q] and only commit what they certify (the first pass's
  // best-so-far answer stays in place otherwise and the query goes to the final uncertified list)
  const int64_t row = tail.qmap ? static_cast<int64_t>(tail.qmap[q]) : q;
  const int64_t out_rows = tail.out_rows > 0 ? tail.out_rows : nq;
  const bool commit = certified || !tail.commit_certified_only;
  if (tid == 0 && !certified) {
    const int at = atomicAdd(&state[0], 1);
    HCIR_DEV_CHECK(at >= 0 && at < out_rows);
    uncert_list[at] = static_cast<int32_t>(row);
  }
  if (!commit) return;   // block-uniform; the caller still runs the done count
  const int nres = nr < k ? nr : k;  // (< k only for uncertified queries, which are completed later)
  for (int j = tid; j < k; j += kSelThreads) {
    out_sim[row * k + j] = rsim[j];
    out_idx[row * k + j] = (j < nres) ? static_cast<int64_t>(key_idx(res[j])) + idx_offset : -1;
    if (tail.out_lab) tail.out_lab[row * k + j] = lab[j];
  }

  // ======================= tail: vote -> peers =======================
  int64_t pred_val = 0;
  const bool do_vote = (tail.pred != nullptr) || (tail.payload == 2);
  if (do_vote && warp == 0) {
    const int best_c = warp_vote(rsim, lab, nres, tail.num_classes, tail.T, lane);
    pred_val = tail.classes ? tail.classes[best_c] : static_cast<int64_t>(best_c);
    if (lane == 0 && tail.pred) tail.pred[row] = pred_val;
  }
  if (tail.world > 0) {
    // this query's results go straight into slot (parity of this step, rank) of EVERY rank's region
    const int64_t st = *tail.step + 1;
    const size_t slot_off = kPeerHdrBytes + (static_cast<size_t>(st & 1) * tail.world + tail.rank) * tail.slot_stride;
    if (tail.payload == 1) {
      const size_t e = static_cast<size_t>(out_rows) * k;
      for (int i = tid; i < k * tail.world; i += kSelThreads) {
        const int g = i / k, j = i - g * k;
        char* base = tail.region[g] + slot_off;
        const size_t at = static_cast<size_t>(row) * k + j;
        const uint64_t key = (j < nres) ? res[j] : 0ull;
        reinterpret_cast<int64_t*>(base)[at] = (j < nres) ? static_cast<int64_t>(key_idx(key)) + idx_offset : -1;
        reinterpret_cast<float*>(base + e * 8)[at] = (j < nres) ? key_sim(key) : -INFINITY;
        if (tail.labels) reinterpret_cast<int32_t*>(base + e * 12)[at] = lab[j];
      }
    } else if (tail.payload == 2) {
      if (warp == 0 && lane < tail.world) reinterpret_cast<int64_t*>(tail.region[lane] + slot_off)[row] = pred_val;
    }
  }
}

// kSelThreads: 128 threads (k <= 32, many queries: twice as many queries in flight hide the barrier / gather
// latencies), 256 in general, 1024 for few queries (streaming regime: the per-query latency IS the kernel
// time, so the whole CTA width goes to one query).
template <int kSelThreads>
__global__ void __launch_bounds__(kSelThreads, (kSelThreads >= 1024) ? 1 : HCIR_K3_THREADS_PER_SM / kSelThreads)
select_rescore_kernel(const __grid_constant__ K3Args a) {
  pdl_wait();
  const int64_t q = blockIdx.x;
  const TailParams& tail = a.tail;
  int32_t* state = a.state;
  const int64_t nq = a.nq;
  const int tid = threadIdx.x;
  const bool live = (tail.active == nullptr) || (q < static_cast<int64_t>(*tail.active));
  if (live) select_rescore_body<kSelThreads>(a, q);
  // ---- last CTA of the launch: publish the uncertified count, reset the counters, signal the peers ----
  __syncthreads();
  if (tid == 0) {
    if (tail.world > 0) __threadfence_system(); else __threadfence();
    const int done = atomicAdd(&state[2], 1);
    if (done == static_cast<int>(nq) - 1) {
      __threadfence();
      const int n_unc = atomicExch(&state[0], 0);
      state[1] = n_unc;
      state[2] = 0;
      if (tail.world > 0 && !tail.no_signal) {
        const int64_t st = *tail.step + 1;
        const size_t par = static_cast<size_t>(st & 1);
        for (int g = 0; g < tail.world; ++g)
          reinterpret_cast<int64_t*>(tail.region[g])[kPeerHdrMeta + par * kPeerMax + tail.rank] = n_unc;
        // every CTA's block stores (ordered before its done-count) and the metas, then the arrival
        __threadfence_system();
        for (int g = 0; g < tail.world; ++g)
          atomicAdd_system(reinterpret_cast<unsigned long long*>(tail.region[g]) + kPeerHdrArrivals + tail.rank, 1ull);
      }
    }
  }
}

// ---- device-driven completion, step 1: the compact batch of uncertified queries --------------------
// grid = capacity CTAs (one per compact row) + grid-stride over the overflow; block 128.
__global__ void __launch_bounds__(128)
retry_setup_kernel(const uint16_t* __restrict__ qbf, const float* __restrict__ q32, const float* __restrict__ qdl,
                   int ld, int64_t nq, int k, const float* __restrict__ out_sim,
                   const int32_t* __restrict__ unc_list, const int32_t* __restrict__ unc_state, float g_delta_max,
                   float eps_acc, int capacity, uint16_t* __restrict__ q2bf, float* __restrict__ q2f,
                   float* __restrict__ q2d, float* __restrict__ thr0, float* __restrict__ thr_hi,
                   int32_t* __restrict__ qmap, int32_t* __restrict__ active, int32_t* __restrict__ final_list,
                   int32_t* __restrict__ final_state) {
  pdl_wait();
  const int count = unc_state[1];  // uncertified queries of the first pass
  const int n = count < capacity ? count : capacity;
  const int r = blockIdx.x;
  if (r == 0 && threadIdx.x == 0) *active = n;
  if (r < n) {
    const int64_t q = unc_list[r];
    const uint4* sb = reinterpret_cast<const uint4*>(qbf + q * ld);
    uint4* db = reinterpret_cast<uint4*>(q2bf + static_cast<int64_t>(r) * ld);
    for (int c = threadIdx.x; c < ld / 8; c += blockDim.x) db[c] = sb[c];
    const float4* sf = reinterpret_cast<const float4*>(q32 + q * ld);
    float4* df = reinterpret_cast<float4*>(q2f + static_cast<int64_t>(r) * ld);
    for (int c = threadIdx.x; c < ld / 4; c += blockDim.x) df[c] = sf[c];
    if (threadIdx.x == 0) {
      const float dq = qdl ? qdl[q] : 0.0f;
      const float eps = g_delta_max * (1.0f + dq) + dq * (1.0f + 1e-6f) + eps_acc;
      // every true top-k row scores >= s_k in fp32, hence > s_k - eps in the bf16 contraction
      const float t = out_sim[q * k + (k - 1)] - eps * 1.001f - 1e-6f;
      q2d[r] = dq;
      thr0[r] = t;
      thr_hi[r] = t;
      qmap[r] = static_cast<int32_t>(q);
    }
  } else if (threadIdx.x == 0) {
    thr0[r] = INFINITY;  // nothing passes; the row's K3 CTA is not live
    thr_hi[r] = INFINITY;
  }
  // more uncertified queries than the batch holds: they go straight to the final list (host completion)
  for (int i = capacity + blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    final_list[atomicAdd(&final_state[0], 1)] = unc_list[i];
}

static int launch_select(const float* q_f32, const float* g_f32, int ld, int64_t nq, int64_t ng, int k,
                         int64_t idx_offset, const hcir_plan_t* plan, const void* workspace, const float* q_delta,
                         float g_delta_max, float eps_acc, float* out_sim, int64_t* out_idx, int32_t* uncert_list,
                         int32_t* state, const hcir_tail_t* tail, hcir_stream_t stream) {
  HCIR_REQUIRE(plan != nullptr, "select_rescore: null plan");
  HCIR_REQUIRE(ld > 0 && ld % 64 == 0, "select_rescore: ld=%d must be a positive multiple of 64", ld);
  HCIR_REQUIRE(nq >= 0 && ng > 0, "select_rescore: bad shape");
  HCIR_REQUIRE(k > 0 && k <= ng && k <= plan->kc, "select_rescore: need 1 <= k=%d <= min(ng=%lld, kc=%d)", k,
               (long long)ng, plan->kc);
  HCIR_REQUIRE(q_f32 && g_f32 && workspace && out_sim && out_idx && uncert_list && state,
               "select_rescore: null pointer");
  TailParams tp{};
  if (tail != nullptr) {
    HCIR_REQUIRE(tail->world >= 0 && tail->world <= kPeerMax && tail->rank >= 0 &&
                     (tail->world == 0 || tail->rank < tail->world),
                 "select_rescore: bad tail rank %d / world %d", tail->rank, tail->world);
    HCIR_REQUIRE(tail->payload >= 0 && tail->payload <= 2, "select_rescore: bad tail payload %d", tail->payload);
    const bool votes = tail->pred != nullptr || (tail->world > 0 && tail->payload == 2);
    HCIR_REQUIRE(!votes || (tail->labels != nullptr && tail->num_classes > 0),
                 "select_rescore: a vote needs labels and num_classes");
    HCIR_REQUIRE(tail->world == 0 || (tail->step != nullptr && tail->payload != 0),
                 "select_rescore: a peer tail needs a step counter and a payload");
    HCIR_REQUIRE(tail->out_lab == nullptr || tail->labels != nullptr, "select_rescore: out_lab without labels");
    tp.labels = tail->labels;
    tp.n_labels = tail->n_labels;
    tp.num_classes = tail->num_classes;
    tp.T = tail->T;
    tp.classes = tail->classes;
    tp.pred = tail->pred;
    tp.out_lab = tail->out_lab;
    tp.world = tail->world;
    tp.rank = tail->rank;
    tp.payload = tail->world > 0 ? tail->payload : 0;
    tp.slot_stride = (tail->slot_bytes + 255) / 256 * 256;
    tp.step = tail->step;
    tp.qmap = tail->qmap;
    tp.active = tail->active;
    tp.commit_certified_only = tail->commit_certified_only;
    tp.no_signal = tail->no_signal;
    tp.out_rows = tail->out_rows;
    HCIR_REQUIRE(tail->out_rows >= 0 && (tail->qmap != nullptr || tail->out_rows == 0 || tail->out_rows == nq),
                 "select_rescore: out_rows without a query map");
    if (tail->world > 0) {
      const int64_t rows = tail->out_rows > 0 ? tail->out_rows : nq;
      const size_t need = tail->payload == 1 ? hcir_packed_block_bytes(rows, k, tail->labels != nullptr ? 1 : 0)
                                             : static_cast<size_t>(rows) * 8;
      HCIR_REQUIRE(need <= tail->slot_bytes, "select_rescore: peer slot of %zu bytes is smaller than the %zu-byte block",
                   static_cast<size_t>(tail->slot_bytes), need);
      for (int g = 0; g < tail->world; ++g) {
        HCIR_REQUIRE(tail->regions[g] != nullptr, "select_rescore: peer region %d is null", g);
        tp.region[g] = static_cast<char*>(tail->regions[g]);
      }
    }
  }
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nq == 0) return HCIR_OK;
  const SelSmem L = sel_smem_layout(plan->nlists, plan->kc, ld);
  HCIR_REQUIRE(L.total <= 220 * 1024, "select_rescore: kc=%d ld=%d needs %zu B of shared memory", plan->kc, ld,
               L.total);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const char* ws = static_cast<const char*>(workspace);
  const int32_t* counts = reinterpret_cast<const int32_t*>(ws + plan->counts_off);
  const uint64_t* cand = reinterpret_cast<const uint64_t*>(ws + plan->keys_off);
  const float* thr_out = reinterpret_cast<const float*>(ws + plan->thr_out_off);
  const float* thr_hi = plan->sample_rows > 0 ? reinterpret_cast<const float*>(ws + plan->thr_hi_off) : nullptr;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
#define HCIR_LAUNCH_SEL(T_)                                                                                        \
  do {                                                                                                             \
    HCIR_CUDA_TRY(cudaFuncSetAttribute(select_rescore_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                       static_cast<int>(L.total)));                                                \
    HCIR_CUDA_TRY(launch_pdl(select_rescore_kernel<T_>, dim3(static_cast<unsigned>(nq)), dim3(T_), L.total, st,    \
                             args));                                                                               \
  } while (0)
  K3Args args{};
  args.q32 = q_f32;
  args.g32 = g_f32;
  args.ld = ld;
  args.nq = nq;
  args.ng = ng;
  args.k = k;
  args.idx_offset = idx_offset;
  args.nlists = plan->nlists;
  args.cap = plan->cap;
  args.kc = plan->kc;
  args.counts = counts;
  args.cand = cand;
  args.thr_out = thr_out;
  args.thr_hi = thr_hi;
  args.q_delta = q_delta;
  args.g_delta_max = g_delta_max;
  args.eps_acc = eps_acc;
  args.out_sim = out_sim;
  args.out_idx = out_idx;
  args.uncert_list = uncert_list;
  args.state = state;
  args.L = L;
  args.tail = tp;
  // CTA width: forced by the plan flags (measurement aid) or chosen from the shape
  const int width = (plan->flags & HCIR_FLAG_K3_WIDTH_MASK) >> HCIR_FLAG_K3_WIDTH_SHIFT;
  if (width == 3 || (width == 0 && nq <= 2 * static_cast<int64_t>(sms))) HCIR_LAUNCH_SEL(1024);
  else if (width == 1 || (width == 0 && k <= 32 && nq >= 512)) HCIR_LAUNCH_SEL(128);  // r3b sweep: 128 wins from ~600 queries
  else HCIR_LAUNCH_SEL(256);
#undef HCIR_LAUNCH_SEL
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

}  // namespace hcir

extern "C" int hcir_select_rescore(const float* q_f32, const float* g_f32, int ld, int64_t nq, int64_t ng, int k,
                                   int64_t idx_offset, const hcir_plan_t* plan, const void* workspace,
                                   const float* q_delta, float g_delta_max, float eps_acc, float* out_sim,
                                   int64_t* out_idx, int32_t* uncert_list, int32_t* uncert_state,
                                   const hcir_tail_t* tail, hcir_stream_t stream) {
  return hcir::launch_select(q_f32, g_f32, ld, nq, ng, k, idx_offset, plan, workspace, q_delta, g_delta_max, eps_acc,
                             out_sim, out_idx, uncert_list, uncert_state, tail, stream);
}

extern "C" int hcir_retry_setup(const uint16_t* q_bf16, const float* q_f32, const float* q_delta, int ld, int64_t nq,
                                int k, const float* out_sim, const int32_t* uncert_list, const int32_t* uncert_state,
                                float g_delta_max, float eps_acc, int capacity, uint16_t* q2_bf16, float* q2_f32,
                                float* q2_delta, float* thr0_2, float* thr_hi_2, int32_t* qmap, int32_t* active,
                                int32_t* final_list, int32_t* final_state, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(ld > 0 && ld % 64 == 0 && nq > 0 && k > 0 && capacity > 0 && capacity <= 4096,
               "retry_setup: bad shape ld=%d nq=%lld k=%d capacity=%d", ld, (long long)nq, k, capacity);
  HCIR_REQUIRE(q_bf16 && q_f32 && out_sim && uncert_list && uncert_state && q2_bf16 && q2_f32 && q2_delta && thr0_2 &&
                   thr_hi_2 && qmap && active && final_list && final_state,
               "retry_setup: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  HCIR_CUDA_TRY(launch_pdl(retry_setup_kernel, dim3(static_cast<unsigned>(capacity)), dim3(128), 0,
                           static_cast<cudaStream_t>(stream), q_bf16, q_f32, q_delta, ld, nq, k, out_sim, uncert_list,
                           uncert_state, g_delta_max, eps_acc, capacity, q2_bf16, q2_f32, q2_delta, thr0_2, thr_hi_2, qmap,
                           active, final_list, final_state));
  return HCIR_OK;
}
