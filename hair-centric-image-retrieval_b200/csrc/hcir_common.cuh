// Shared device/host helpers for the hcir_b200 kernels (sm_100a only).
//
// Internal data layout (see DESIGN.md "Data layout in HBM"):
//   * gallery / query banks are row-major with a row stride `ld` = D rounded up to 64
//     elements, zero padded, so every row is 16-byte aligned, TMA-legal and the fp32 dot
//     product needs no tail handling (zero padding contributes fma(0,0,acc) == acc exactly);
//   * a top-k candidate is a single 64-bit KEY = (order-preserving bits of the fp32
//     similarity) << 32 | (0xFFFFFFFF - gallery_index).  Larger key == better candidate:
//     descending similarity, ties -> ascending gallery index.  This is the build's canonical
//     order and makes every selection / merge / sort an exact
//     operation on unique integers.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/hcir_b200.h"

namespace hcir {

// ---- error plumbing (api.cu) -----------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int check_device();  // HCIR_OK iff current device is sm_100

#define HCIR_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::hcir::cuda_fail(_e, #expr); \
  } while (0)

#define HCIR_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ::hcir::set_error(__VA_ARGS__); \
      return HCIR_EINVAL;            \
    }                                \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;

// ---- device-side bounds checks (debug build) ---------------------------------------------
// compute-sanitizer is closed on the GPU pool this was developed on, so the indices the sanitizer would
// have watched -- list appends of the K2 epilogue, the shared-memory stages / histograms / rank arrays of
// K3, the merge ranks of K5 -- carry explicit checks that a build with -DHCIR_BOUNDS_CHECK
// (HCIR_NVCC_EXTRA="-DHCIR_BOUNDS_CHECK") turns into printf + trap.  The whole -m gpu suite runs clean
// under that build (profiles/README.md); production builds compile the checks away.
#ifdef HCIR_BOUNDS_CHECK
#define HCIR_DEV_CHECK(cond)                                                                              \
  do {                                                                                                    \
    if (!(cond)) {                                                                                        \
      printf("HCIR bounds check failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__,   \
             static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));                                \
      __trap();                                                                                           \
    }                                                                                                     \
  } while (0)
#else
#define HCIR_DEV_CHECK(cond) \
  do {                       \
  } while (0)
#endif

// ---- programmatic dependent launch ---------------------------------------------------------
// The kernels of a search step are launched with the programmatic-stream-serialization attribute and
// call pdl_wait() (griddepcontrol.wait: returns once the preceding kernel of the stream has completed
// and its writes are visible) before they touch anything a predecessor produced.  The dependent grid's
// launch latency and prologue (barrier init, TMEM allocation, tensor-map prefetch) then overlap the
// predecessor's tail instead of following its completion -- the step is a chain of 5-6 short kernels
// in one CUDA graph, so the node-to-node gaps are a measurable part of the streaming-regime step.
// Without the launch attribute the instruction is a no-op.  HCIR_PDL=0 (environment) turns it off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
bool pdl_enabled();  // api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- peer region header (peer.cu; K3's tail and the fused wait+merge+vote kernel write / poll it) ----
//   int64 words: [0..15] arrivals[r]  steps whose block from rank r has landed here (monotone)
//                [16..31] meta[0][r], [32..47] meta[1][r]  one word per rank and parity (uncertified count)
//                [48] step_seen  last completed step   [49] error  != 0: a wait timed out at that step
//                [50] consumer done-CTA counter        [51] producer done-CTA counter (standalone push)
constexpr int kPeerMax = 16;
constexpr size_t kPeerHdrBytes = 512;
constexpr int kPeerHdrArrivals = 0, kPeerHdrMeta = 16, kPeerHdrStep = 48, kPeerHdrError = 49;
constexpr int kPeerHdrConsDone = 50, kPeerHdrProdDone = 51;

__host__ __device__ inline int64_t ceil_div_i64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up_int(int a, int b) { return (a + b - 1) / b * b; }

// ---- candidate keys ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2ord(float f) {
  f += 0.0f;  // canonicalise -0.0 -> +0.0 so that equal similarities have equal bits
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float sim, uint32_t idx) {
  return (static_cast<uint64_t>(f2ord(sim)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float key_sim(uint64_t key) { return ord2f(static_cast<uint32_t>(key >> 32)); }
// RAW key (what the simtopk epilogue appends: fp32 bits << 32 | ~index) <-> ordered key
__device__ __forceinline__ uint64_t raw2key(uint64_t raw) {
  return (static_cast<uint64_t>(f2ord(__uint_as_float(static_cast<uint32_t>(raw >> 32)))) << 32) |
         (raw & 0xFFFFFFFFull);
}
__device__ __forceinline__ uint64_t key2raw(uint64_t key) {
  return (static_cast<uint64_t>(__float_as_uint(key_sim(key))) << 32) | (key & 0xFFFFFFFFull);
}
__device__ __forceinline__ uint32_t key_idx(uint64_t key) { return 0xFFFFFFFFu - static_cast<uint32_t>(key); }

// ---- canonical fp32 dot product ----------------------------------------------------------
// One warp per (query row, gallery row).  Lane l walks float4 chunks l, l+32, ... with one
// sequential fma chain, then a xor-butterfly adds the 32 partials.  EVERY exact similarity
// this library returns comes from this function, so single-GPU, sharded, re-scored and
// fallback results are bit-identical to each other regardless of which kernel produced them.
__device__ __forceinline__ float canonical_dot(const float4* __restrict__ q4,
                                               const float4* __restrict__ g4, int ld4, int lane) {
  float acc = 0.0f;
  for (int c = lane; c < ld4; c += kWarp) {
    const float4 a = q4[c];
    const float4 b = __ldg(g4 + c);
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  return acc;
}

// Two rows at once (same per-row arithmetic order as canonical_dot, so bit-identical results).  The row
// gathers of the re-score are latency-bound: the loads of kDotGroup consecutive chunks of BOTH rows
// (2 * kDotGroup independent 16-byte loads per lane) are issued before any is consumed, so a warp keeps
// 4 KiB in flight instead of whatever the compiler's unrolling happens to stagger (A/B on the GPU box,
// profiles/README.md r2d: groups of 2 / 3 / 4 chunks -> K3 on the C5 shard 9.9 / 9.2 / 8.7 ms).
#ifndef HCIR_DOT_GROUP
#define HCIR_DOT_GROUP 4
#endif
constexpr int kDotGroup = HCIR_DOT_GROUP;
__device__ __forceinline__ void canonical_dot2(const float4* __restrict__ q4, const float4* __restrict__ ga,
                                               const float4* __restrict__ gb, int ld4, int lane, float& sa,
                                               float& sb) {
  float a = 0.0f, b = 0.0f;
  for (int c0 = lane; c0 < ld4; c0 += kWarp * kDotGroup) {
    float4 y[kDotGroup], z[kDotGroup];
#pragma unroll
    for (int u = 0; u < kDotGroup; ++u) {
      const int c = c0 + u * kWarp;
      if (c < ld4) {
        y[u] = __ldg(ga + c);
        z[u] = __ldg(gb + c);
      }
    }
#pragma unroll
    for (int u = 0; u < kDotGroup; ++u) {
      const int c = c0 + u * kWarp;
      if (c < ld4) {
        const float4 x = q4[c];
        a = fmaf(x.x, y[u].x, a);
        a = fmaf(x.y, y[u].y, a);
        a = fmaf(x.z, y[u].z, a);
        a = fmaf(x.w, y[u].w, a);
        b = fmaf(x.x, z[u].x, b);
        b = fmaf(x.y, z[u].y, b);
        b = fmaf(x.z, z[u].z, b);
        b = fmaf(x.w, z[u].w, b);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o);
    b += __shfl_xor_sync(kFull, b, o);
  }
  sa = a;
  sb = b;
}

// ---- kNN vote by one warp ------------------------------------------------------------------
// s[0..n) similarities in rank order, l[0..n) neighbour class indices (shared memory).  T <= 0:
// uniform majority vote == sklearn `_mode`; T > 0: score[c] = sum exp((s_j - s_0) / T) accumulated in
// rank order.  Lane l owns classes l, l+32, ...; arg-max, ties -> smallest class.  Every lane returns
// the winning class index.  One definition for every kernel that votes (K4, K3's tail, K5's tail), so
// their predictions are bit-identical.
__device__ __forceinline__ int warp_vote(const float* s, const int32_t* l, int n, int num_classes, float T,
                                         int lane) {
  const bool weighted = T > 0.0f;
  const float s0 = (weighted && n > 0) ? s[0] : 0.0f;
  const float inv_t = weighted ? 1.0f / T : 0.0f;
  float best = -1.0f;
  int best_c = 0x7FFFFFFF;
  for (int c = lane; c < num_classes; c += kWarp) {
    float acc = 0.0f;
    for (int j = 0; j < n; ++j) {
      if (l[j] == c) acc += weighted ? expf((s[j] - s0) * inv_t) : 1.0f;
    }
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
  return best_c;
}

__device__ __forceinline__ int64_t ld_acquire_sys(const int64_t* p) {
  int64_t v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Poll arrivals[lane] (lane < world) of a peer region header until every rank's block of step `st` has
// landed, bounded by timeout_ns.  Called by one warp; returns true to every lane iff all arrived.
__device__ __forceinline__ bool peer_wait_arrivals(const int64_t* hdr, int world, int64_t st, int64_t timeout_ns,
                                                   int lane) {
  bool ok = true;
  if (lane < world) {
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(hdr + kPeerHdrArrivals + lane) < st) {
      __nanosleep(32);
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (static_cast<int64_t>(t1 - t0) > timeout_ns) {  // a peer never arrived: report, do not hang the GPU
        ok = false;
        break;
      }
    }
  }
  return __all_sync(kFull, ok);
}

// ---- warp-cooperative exact selection ("prune") -----------------------------------------
// Keep the `keep` largest of the `cnt` (> keep) unique keys in buf[0..cnt) -- a buffer in
// GLOBAL memory (L2 resident) -- compacted in place to buf[0..keep) in arbitrary order.
// Returns the keep-th largest key.  MSB-first radix select, 8-bit digits; `hist` is a
// per-warp shared-memory scratch of 256 words.  All 32 lanes must call it convergently.
// kRaw: the buffer holds RAW keys (see raw2key); the returned threshold is an ordered key.
template <bool kRaw = false>
__device__ inline uint64_t warp_prune(uint64_t* buf, int cnt, int keep, uint32_t* hist, int lane) {
  auto load = [&](int i) -> uint64_t {
    const uint64_t x = __ldcg(buf + i);
    return kRaw ? raw2key(x) : x;
  };
  uint64_t prefix = 0, mask = 0;
  uint32_t remaining = static_cast<uint32_t>(keep);
  __syncwarp();
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = lane; b < 256; b += kWarp) hist[b] = 0;
    __syncwarp();
    for (int i = lane; i < cnt; i += kWarp) {
      const uint64_t key = load(i);
      if ((key & mask) == prefix) atomicAdd(&hist[static_cast<uint32_t>(key >> shift) & 0xFFu], 1u);
    }
    __syncwarp();
    // lane l owns bins 255-8l .. 248-8l (descending), so a lane-prefix is a "count above".
    uint32_t c[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      c[j] = hist[255 - (8 * lane + j)];
      s += c[j];
    }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < kWarp; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - s;
    const bool mine = (excl < remaining) && (remaining <= incl);
    const int src = __ffs(__ballot_sync(kFull, mine)) - 1;
    uint32_t digit = 0, newrem = 0, dcount = 0;
    if (mine) {
      uint32_t run = excl;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (dcount == 0 && run + c[j] >= remaining) {
          digit = 255 - (8 * lane + j);
          newrem = remaining - run;
          dcount = c[j];
        }
        run += c[j];
      }
    }
    digit = __shfl_sync(kFull, digit, src);
    remaining = __shfl_sync(kFull, newrem, src);
    dcount = __shfl_sync(kFull, dcount, src);
    prefix |= static_cast<uint64_t>(digit) << shift;
    mask |= 0xFFull << shift;
    __syncwarp();
    if (dcount == 1 && shift > 0) {
      // exactly one key carries this prefix: it IS the keep-th largest; fetch it and stop.
      uint64_t found = 0;
      for (int i = lane; i < cnt; i += kWarp) {
        const uint64_t key = load(i);
        if ((key & mask) == prefix) found = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) found |= __shfl_xor_sync(kFull, found, o);
      prefix = found;
      break;
    }
  }
  const uint64_t thr = prefix;
  int out = 0;
  for (int base = 0; base < cnt; base += kWarp) {
    const int i = base + lane;
    const uint64_t key = (i < cnt) ? load(i) : 0ull;
    const bool keepit = (i < cnt) && (key >= thr);
    const uint32_t b = __ballot_sync(kFull, keepit);
    __syncwarp();
    if (keepit) buf[out + __popc(b & ((1u << lane) - 1u))] = kRaw ? key2raw(key) : key;
    out += __popc(b);
    __syncwarp();
  }
  return thr;
}

// ---- block-cooperative selection on a shared-memory key array ----------------------------
// Find the `keep`-th largest of keys[0..cnt) (unique keys, 1 <= keep <= cnt) and return it to
// every thread.  If `out` is non-null the `keep` keys >= that threshold are also written to
// out[0..keep) (arbitrary order; `out` must not alias `keys`).  MSB-first radix select with
// 8-bit digits.  `hist` = 256 words, `scratch` = 4 words of shared memory.  The whole block
// must call it convergently.
__device__ inline uint64_t block_select(const uint64_t* keys, int cnt, int keep, uint64_t* out,
                                        uint32_t* hist, uint32_t* scratch) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  uint64_t prefix = 0, mask = 0;
  uint32_t remaining = static_cast<uint32_t>(keep);
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = tid; b < 256; b += nthr) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < cnt; i += nthr) {
      const uint64_t key = keys[i];
      if ((key & mask) == prefix) atomicAdd(&hist[static_cast<uint32_t>(key >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (tid < kWarp) {  // warp 0 scans the 256 bins, top bin first
      const int lane = tid;
      uint32_t c[8], s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c[j] = hist[255 - (8 * lane + j)];
        s += c[j];
      }
      uint32_t incl = s;
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t excl = incl - s;
      if ((excl < remaining) && (remaining <= incl)) {
        uint32_t run = excl;
        bool done = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (!done && run + c[j] >= remaining) {
            scratch[0] = 255 - (8 * lane + j);
            scratch[1] = remaining - run;
            done = true;
          }
          run += c[j];
        }
      }
    }
    __syncthreads();
    prefix |= static_cast<uint64_t>(scratch[0]) << shift;
    mask |= 0xFFull << shift;
    remaining = scratch[1];
    __syncthreads();
  }
  const uint64_t thr = prefix;
  if (out != nullptr) {
    if (tid == 0) scratch[2] = 0;
    __syncthreads();
    for (int i = tid; i < cnt; i += nthr) {
      const uint64_t key = keys[i];
      if (key >= thr) out[atomicAdd(&scratch[2], 1u)] = key;
    }
    __syncthreads();
  }
  return thr;
}

// ---- block-cooperative k-th largest of 32-bit ordered values ------------------------------------
// vals[0..n) in shared memory (not modified), 1 <= k <= n.  Returns the k-th largest VALUE to every
// thread (ties are fine: duplicates count).  Linear 256-bin histogram over [lo, hi], narrowed to the
// bin that holds the k-th largest until the bin is one value wide -- two passes for typical score
// distributions, where the MSB-first radix select needs four or eight.  `hist` = 256 words,
// `scratch` = 4 words of shared memory.  The whole block must call it convergently.
__device__ inline uint32_t block_kth_u32(const uint32_t* vals, int n, int k, uint32_t* hist, uint32_t* scratch) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31;
  if (tid == 0) { scratch[0] = 0xFFFFFFFFu; scratch[1] = 0u; }
  __syncthreads();
  {
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = tid; i < n; i += nthr) { mn = min(mn, vals[i]); mx = max(mx, vals[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(kFull, mn, o));
      mx = max(mx, __shfl_xor_sync(kFull, mx, o));
    }
    if (lane == 0) { atomicMin(&scratch[0], mn); atomicMax(&scratch[1], mx); }
  }
  __syncthreads();
  uint32_t lo = scratch[0], hi = scratch[1], need = static_cast<uint32_t>(k);
  __syncthreads();
  while (lo < hi) {
    const uint32_t span = hi - lo;
    int shift = 0;
    while ((span >> shift) >= 256u) ++shift;
    for (int b = tid; b < 256; b += nthr) hist[b] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nthr) {
      const uint32_t v = vals[i];
      if (v >= lo && v <= hi) atomicAdd(&hist[(v - lo) >> shift], 1u);
    }
    __syncthreads();
    if (tid < kWarp) {  // warp 0 scans from the top bin: lane l owns bins 255-8l .. 248-8l
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[255 - (8 * lane + j)]; sum += c[j]; }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t excl = incl - sum;
      if (excl < need && need <= incl) {
        uint32_t run = excl;
        bool done = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (!done && run + c[j] >= need) {
            scratch[2] = 255 - (8 * lane + j);
            scratch[3] = need - run;
            done = true;
          }
          run += c[j];
        }
      }
    }
    __syncthreads();
    const uint32_t b = scratch[2];
    need = scratch[3];
    const uint32_t nlo = lo + (b << shift);
    const uint32_t nhi = (shift == 0) ? nlo : min(hi, nlo + ((1u << shift) - 1u));
    lo = nlo;
    hi = nhi;
    __syncthreads();
  }
  return lo;
}

}  // namespace hcir
