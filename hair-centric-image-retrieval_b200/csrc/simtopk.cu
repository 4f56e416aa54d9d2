// K2: similarity contraction (bf16 x bf16 -> fp32, tcgen05 tensor cores, TMEM accumulators,
// TMA-fed) fused with a running per-query top-kc.  The [nq, ng] similarity matrix never
// leaves the SM: each 128x256 accumulator tile is read back from TMEM by four epilogue warps
// (one thread per query row) and compared against that query's running threshold; only the
// rare survivors are appended (as 64-bit keys) to a per-(query, split) candidate list in L2
// and pruned back to kc with a warp-cooperative radix select when the list fills.
//
// Roofline: tensor pipe for large query batches (2*nq*ng*ld flops), HBM for nq <= ~128
// (ng*ld*2 gallery bytes streamed once).  See DESIGN.md for the numbers.
//
// Warp roles (256 threads, 1 CTA / SM, persistent over work items):
//   warp 0    TMA producer     (one lane)   global -> 4-stage smem ring, SWIZZLE_128B
//   warp 1    MMA issuer       (one lane)   tcgen05.mma 128x256x16, 2 TMEM accumulator stages
//   warp 2    TMEM allocator
//   warps 4-7 epilogue         (128 threads = 128 TMEM lanes = 128 query rows)
// Work item = (query tile of 128 rows) x (gallery split = contiguous range of 256-row tiles).
#include <cuda.h>

#include "hcir_common.cuh"
#include "hcir_ptx.cuh"

namespace hcir {

constexpr int kBlockM = 128;   // query rows per tile  (UMMA M, TMEM lanes)
constexpr int kBlockN = 256;   // gallery rows per tile (UMMA N, TMEM columns)
constexpr int kBlockK = 64;    // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kBBytes = kBlockN * kBlockK * 2;  // 32 KiB
constexpr int kAccStages = 2;
constexpr int kTmemCols = kAccStages * kBlockN;  // 512: all of TMEM
constexpr int kSimThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr int kNumEpiWarps = 4;
constexpr size_t kSimSmemBytes = 1024 /*align slack*/ + static_cast<size_t>(kStages) * (kABytes + kBBytes) +
                                 256 /*barriers + tmem slot*/ + kNumEpiWarps * 256 * sizeof(uint32_t);

struct SimParams {
  int64_t nq, ng;
  int ld, num_kb, num_qt, tiles_total, tiles_per_split, nsplit, cap, kc, num_items;
  int32_t* counts;
  uint64_t* keys;
  float* dump;
};

// Scan one 32-column chunk of accumulator values for one query row.
template <bool kBounded>
__device__ __forceinline__ void scan_chunk(const uint32_t (&v)[32], float thr, uint64_t* buf, int& cnt,
                                           uint32_t gcol0, int lim) {
  bool any = false;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const bool hit = __uint_as_float(v[j]) > thr;
    any |= kBounded ? (hit && j < lim) : hit;
  }
  if (any) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float f = __uint_as_float(v[j]);
      if ((f > thr) && (!kBounded || j < lim)) {
        buf[cnt] = make_key(f, gcol0 + j);
        ++cnt;
      }
    }
  }
}

template <bool kDump>
__global__ void __launch_bounds__(kSimThreads, 1)
simtopk_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_g,
               const SimParams p) {
  extern __shared__ uint8_t smem_dyn[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment; align by hand, do not trust the base.
  uint8_t* smem = smem_dyn + ((1024u - (ptx::smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytes));
  uint64_t* full_bar = bars;                          // [kStages]   TMA -> MMA
  uint64_t* empty_bar = bars + kStages;               // [kStages]   MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;           // [kAccStages] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + kAccStages;  // [kAccStages] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);
  uint32_t* hist_all = reinterpret_cast<uint32_t*>(smem + kStages * (kABytes + kBBytes) + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_g);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], kNumEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, kTmemCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int split = item / p.num_qt, qt = item - split * p.num_qt;
        const int t0 = split * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.tiles_total);
        for (int t = t0; t < t1; ++t) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], kABytes + kBBytes);
            // queries are re-read by every gallery tile: keep them in L2; gallery streams once
            ptx::tma_load_2d(&tmap_q, &full_bar[stage], smem_a + stage * kABytes, kb * kBlockK, qt * kBlockM,
                             ptx::kEvictLast);
            ptx::tma_load_2d(&tmap_g, &full_bar[stage], smem_b + stage * kBBytes, kb * kBlockK, t * kBlockN,
                             ptx::kEvictNormal);
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBlockM, kBlockN);
      uint32_t stage = 0, phase = 0, iter = 0;
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int split = item / p.num_qt;
        const int t0 = split * p.tiles_per_split;
        const int t1 = min(t0 + p.tiles_per_split, p.tiles_total);
        for (int t = t0; t < t1; ++t, ++iter) {
          const uint32_t acc = iter & 1u, aphase = (iter >> 1) & 1u;
          ptx::mbar_wait(&tempty_bar[acc], aphase ^ 1u);
          ptx::tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBlockN;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(smem_a + stage * kABytes);
            const uint32_t b_addr = ptx::smem_u32(smem_b + stage * kBBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advancing K inside the 128-byte swizzle row = +32 bytes on the start address
              const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * kUmmaK * 2);
              const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * kUmmaK * 2);
              ptx::umma_bf16<1>(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::umma_commit(&empty_bar[stage]);              // smem slot free once these MMAs retire
            if (kb == p.num_kb - 1) ptx::umma_commit(&tfull_bar[acc]);  // accumulator ready
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: TMEM -> threshold filter -> candidate lists =====================
    const int ew = warp - kEpiWarp0;  // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t* hist = hist_all + ew * 256;
    const int row = ew * 32 + lane;
    const int prune_at = p.cap - 32;
    uint32_t iter = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int split = item / p.num_qt, qt = item - split * p.num_qt;
      const int t0 = split * p.tiles_per_split;
      const int t1 = min(t0 + p.tiles_per_split, p.tiles_total);
      const int64_t q = static_cast<int64_t>(qt) * kBlockM + row;
      const bool active = q < p.nq;
      float thr = active ? -INFINITY : INFINITY;
      int cnt = 0;
      uint64_t* buf = p.keys + (active ? (q * p.nsplit + split) * static_cast<int64_t>(p.cap) : 0);
      for (int t = t0; t < t1; ++t, ++iter) {
        const uint32_t acc = iter & 1u, aphase = (iter >> 1) & 1u;
        ptx::mbar_wait(&tfull_bar[acc], aphase);
        ptx::tc_fence_after();
        const int64_t gbase = static_cast<int64_t>(t) * kBlockN;
        const bool full_tile = gbase + kBlockN <= p.ng;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * kBlockN;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr + c * 32, v);
          ptx::tmem_ld_wait();
          if (c == kBlockN / 32 - 1) {
            // whole accumulator stage is in registers now: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
          }
          const uint32_t gcol0 = static_cast<uint32_t>(gbase) + c * 32;
          if (full_tile) {
            scan_chunk<false>(v, thr, buf, cnt, gcol0, 32);
          } else {
            const int64_t rem = p.ng - static_cast<int64_t>(gcol0);
            const int lim = rem >= 32 ? 32 : (rem > 0 ? static_cast<int>(rem) : 0);
            scan_chunk<true>(v, thr, buf, cnt, gcol0, lim);
          }
          if (kDump && active) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t g = static_cast<int64_t>(gcol0) + j;
              if (g < p.ng) p.dump[q * p.ng + g] = __uint_as_float(v[j]);
            }
          }
          // lists that could overflow during the next chunk are pruned back to kc now
          uint32_t need = __ballot_sync(kFull, cnt > prune_at);
          while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const uint64_t bptr = __shfl_sync(kFull, reinterpret_cast<uint64_t>(buf), src);
            const int bcnt = __shfl_sync(kFull, cnt, src);
            const uint64_t tk = warp_prune(reinterpret_cast<uint64_t*>(bptr), bcnt, p.kc, hist, lane);
            if (lane == src) {
              cnt = p.kc;
              thr = key_sim(tk);
            }
          }
        }
      }
      // end of work item: leave at most kc candidates per list, publish the count
      uint32_t need = __ballot_sync(kFull, cnt > p.kc);
      while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const uint64_t bptr = __shfl_sync(kFull, reinterpret_cast<uint64_t>(buf), src);
        const int bcnt = __shfl_sync(kFull, cnt, src);
        warp_prune(reinterpret_cast<uint64_t*>(bptr), bcnt, p.kc, hist, lane);
        if (lane == src) cnt = p.kc;
      }
      if (active) p.counts[q * p.nsplit + split] = cnt;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tensorMapEncodeTiled get_encode_fn() {
  static PFN_tensorMapEncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_tensorMapEncodeTiled>(ptr);
    }
  }
  return fn;
}

// [rows, ld] bf16 row-major -> 2D tensor map, box = 64 x box_rows, 128-byte swizzle, zero OOB fill
static int make_bf16_map(CUtensorMap* map, const void* base, int64_t rows, int ld, int box_rows) {
  PFN_tensorMapEncodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return HCIR_ECUDA;
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld ld=%d box_rows=%d base=%p)",
              static_cast<int>(r), (long long)rows, ld, box_rows, base);
    return HCIR_ECUDA;
  }
  return HCIR_OK;
}

static int launch_simtopk(const uint16_t* q_bf16, int64_t nq, const uint16_t* g_bf16, int64_t ng, int ld,
                          const hcir_plan_t* plan, void* workspace, float* dump, cudaStream_t st) {
  HCIR_REQUIRE(plan != nullptr && workspace != nullptr, "simtopk: null plan/workspace");
  HCIR_REQUIRE(q_bf16 && g_bf16, "simtopk: null operand");
  HCIR_REQUIRE(ld > 0 && ld % 64 == 0, "simtopk: ld=%d must be a positive multiple of 64", ld);
  HCIR_REQUIRE(nq > 0 && ng > 0 && ng < (1ll << 31) - 256, "simtopk: bad shape nq=%lld ng=%lld", (long long)nq,
               (long long)ng);
  HCIR_REQUIRE(reinterpret_cast<uintptr_t>(q_bf16) % 16 == 0 && reinterpret_cast<uintptr_t>(g_bf16) % 16 == 0,
               "simtopk: operands must be 16-byte aligned");
  HCIR_REQUIRE(plan->kc > 0 && plan->cap >= plan->kc + 64 && plan->nsplit > 0, "simtopk: inconsistent plan");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;

  SimParams p;
  p.nq = nq;
  p.ng = ng;
  p.ld = ld;
  p.num_kb = ld / kBlockK;
  p.num_qt = static_cast<int>(ceil_div_i64(nq, kBlockM));
  p.tiles_total = static_cast<int>(ceil_div_i64(ng, kBlockN));
  p.tiles_per_split = static_cast<int>(ceil_div_i64(p.tiles_total, plan->nsplit));
  HCIR_REQUIRE(static_cast<int>(ceil_div_i64(p.tiles_total, p.tiles_per_split)) == plan->nsplit,
               "simtopk: plan.nsplit=%d leaves an empty split for %d tiles", plan->nsplit, p.tiles_total);
  p.nsplit = plan->nsplit;
  p.cap = plan->cap;
  p.kc = plan->kc;
  p.num_items = p.num_qt * p.nsplit;
  p.counts = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + plan->counts_off);
  p.keys = reinterpret_cast<uint64_t*>(static_cast<char*>(workspace) + plan->keys_off);
  p.dump = dump;

  CUtensorMap mq, mg;
  rc = make_bf16_map(&mq, q_bf16, nq, ld, kBlockM);
  if (rc != HCIR_OK) return rc;
  rc = make_bf16_map(&mg, g_bf16, ng, ld, kBlockN);
  if (rc != HCIR_OK) return rc;

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = p.num_items < sms ? p.num_items : sms;
  if (dump != nullptr) {
    HCIR_CUDA_TRY(cudaFuncSetAttribute(simtopk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kSimSmemBytes)));
    simtopk_kernel<true><<<grid, kSimThreads, kSimSmemBytes, st>>>(mq, mg, p);
  } else {
    HCIR_CUDA_TRY(cudaFuncSetAttribute(simtopk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kSimSmemBytes)));
    simtopk_kernel<false><<<grid, kSimThreads, kSimSmemBytes, st>>>(mq, mg, p);
  }
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}

}  // namespace hcir

extern "C" int hcir_simtopk_plan(int64_t nq, int64_t ng, int ld, int kc, int sm_count, hcir_plan_t* plan) {
  using namespace hcir;
  HCIR_REQUIRE(plan != nullptr, "simtopk_plan: null plan");
  HCIR_REQUIRE(nq > 0 && ng > 0 && ld > 0 && kc > 0, "simtopk_plan: bad shape");
  HCIR_REQUIRE(kc <= 8192, "simtopk_plan: kc=%d > 8192 unsupported", kc);
  if (sm_count <= 0) sm_count = 148;
  const int64_t num_qt = ceil_div_i64(nq, kBlockM);
  const int64_t tiles = ceil_div_i64(ng, kBlockN);
  int64_t max_split = tiles;
  if (max_split > 16384 / kc) max_split = 16384 / kc;  // select_rescore keeps nsplit*kc keys in smem
  if (max_split > 4 * sm_count) max_split = 4 * sm_count;
  if (max_split < 1) max_split = 1;
  // static round-robin over persistent CTAs: time ~ waves * (tiles per item + list warm-up)
  const double warmup_tiles = 16.0;
  double best_cost = 1e300;
  int best = 1;
  for (int64_t ns = 1; ns <= max_split; ++ns) {
    const int64_t tps = ceil_div_i64(tiles, ns);
    const int64_t ns_eff = ceil_div_i64(tiles, tps);
    if (ns_eff != ns) continue;
    const int64_t items = num_qt * ns_eff;
    const int64_t waves = ceil_div_i64(items, sm_count);
    const double cost = static_cast<double>(waves) * (static_cast<double>(tps) + warmup_tiles);
    if (cost < best_cost * 0.995) {
      best_cost = cost;
      best = static_cast<int>(ns);
    }
  }
  plan->nsplit = best;
  plan->kc = kc;
  plan->cap = round_up_int(2 * kc + 32, 32);
  plan->reserved = 0;
  plan->counts_off = 0;
  uint64_t off = static_cast<uint64_t>(nq) * best * sizeof(int32_t);
  off = (off + 255) / 256 * 256;
  plan->keys_off = off;
  plan->bytes = off + static_cast<uint64_t>(nq) * best * plan->cap * sizeof(uint64_t);
  return HCIR_OK;
}

extern "C" int hcir_simtopk(const uint16_t* q_bf16, int64_t nq, const uint16_t* g_bf16, int64_t ng, int ld,
                            const hcir_plan_t* plan, void* workspace, hcir_stream_t stream) {
  return hcir::launch_simtopk(q_bf16, nq, g_bf16, ng, ld, plan, workspace, nullptr,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int hcir_simtopk_debug(const uint16_t* q_bf16, int64_t nq, const uint16_t* g_bf16, int64_t ng, int ld,
                                  const hcir_plan_t* plan, void* workspace, float* scores, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(scores != nullptr, "simtopk_debug: null scores");
  return launch_simtopk(q_bf16, nq, g_bf16, ng, ld, plan, workspace, scores, static_cast<cudaStream_t>(stream));
}
