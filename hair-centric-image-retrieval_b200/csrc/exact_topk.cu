// Exact fp32 path on CUDA cores: canonical fp32 similarities + exact top-k, canonical order.
// Used (a) as the on-GPU fallback for queries the tensor-core path could not certify and
// (b) as the whole path for problems too small to be worth a tcgen05 launch.
// HBM/L2-bound: every (query, gallery-row) pair costs one canonical_dot.
#include "hcir_common.cuh"

namespace hcir {

constexpr int kExactWarps = 8;  // queries per CTA, one warp each

struct ExactPlan {
  int nsplit;
  int cap;
  size_t counts_off, keys_off, bytes;
};

static ExactPlan make_exact_plan(int64_t nlist, int64_t ng, int k, int sm_count) {
  ExactPlan p;
  const int64_t qblocks = ceil_div_i64(nlist > 0 ? nlist : 1, kExactWarps);
  int64_t nsplit = ceil_div_i64(4 * static_cast<int64_t>(sm_count), qblocks);
  const int64_t by_rows = ng / 512 > 0 ? ng / 512 : 1;     // >= 512 rows per split
  const int64_t by_smem = 16384 / k > 0 ? 16384 / k : 1;   // finalize keeps nsplit*k keys in smem
  if (nsplit > by_rows) nsplit = by_rows;
  if (nsplit > by_smem) nsplit = by_smem;
  if (nsplit < 1) nsplit = 1;
  p.nsplit = static_cast<int>(nsplit);
  p.cap = round_up_int((2 * k > k + 64 ? 2 * k : k + 64), 32);
  p.counts_off = 0;
  p.keys_off = sizeof(int32_t) * static_cast<size_t>(nlist) * p.nsplit;
  p.keys_off = (p.keys_off + 255) / 256 * 256;
  p.bytes = p.keys_off + static_cast<size_t>(nlist) * p.nsplit * p.cap * sizeof(uint64_t);
  return p;
}

// grid (ceil(nlist/8), nsplit); dynamic smem: 8 * ld floats (query rows) + 8 * 256 words (hist)
__global__ void __launch_bounds__(kExactWarps* kWarp)
exact_scan_kernel(const float* __restrict__ q32, const float* __restrict__ g32, int ld, int64_t ng,
                  const int32_t* __restrict__ qlist, int64_t nlist, int k, int nsplit, int cap,
                  int32_t* __restrict__ counts, uint64_t* __restrict__ keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t slot = static_cast<int64_t>(blockIdx.x) * kExactWarps + warp;
  if (slot >= nlist) return;  // warp-uniform; no block-level sync below
  float* qs = reinterpret_cast<float*>(smem_raw) + static_cast<size_t>(warp) * ld;
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + static_cast<size_t>(kExactWarps) * ld * sizeof(float)) +
                   warp * 256;
  const int64_t qi = qlist ? qlist[slot] : slot;
  const int ld4 = ld >> 2;
  {
    const float4* src = reinterpret_cast<const float4*>(q32 + qi * static_cast<int64_t>(ld));
    for (int c = lane; c < ld4; c += kWarp) reinterpret_cast<float4*>(qs)[c] = __ldg(src + c);
  }
  __syncwarp();
  const int split = blockIdx.y;
  const int64_t rows_per = ceil_div_i64(ng, nsplit);
  const int64_t r0 = split * rows_per;
  const int64_t r1 = (r0 + rows_per < ng) ? r0 + rows_per : ng;
  uint64_t* buf = keys + (slot * nsplit + split) * static_cast<int64_t>(cap);
  uint64_t thr = 0;  // every real key is > 0
  int cnt = 0;
  const float4* q4 = reinterpret_cast<const float4*>(qs);
  for (int64_t r = r0; r < r1; ++r) {
    const float s = canonical_dot(q4, reinterpret_cast<const float4*>(g32 + r * static_cast<int64_t>(ld)), ld4, lane);
    const uint64_t key = make_key(s, static_cast<uint32_t>(r));
    if (key > thr) {  // warp-uniform
      if (lane == 0) buf[cnt] = key;
      ++cnt;
      if (cnt == cap) {
        thr = warp_prune(buf, cnt, k, hist, lane);
        cnt = k;
      }
    }
  }
  if (cnt > k) {
    warp_prune(buf, cnt, k, hist, lane);
    cnt = k;
  }
  if (lane == 0) counts[slot * nsplit + split] = cnt;
}

// grid nlist; block 256; dynamic smem: (nsplit*k + k) keys + hist + scratch
__global__ void __launch_bounds__(256)
exact_finalize_kernel(const int32_t* __restrict__ qlist, int k, int nsplit, int cap,
                      const int32_t* __restrict__ counts, const uint64_t* __restrict__ keys,
                      int64_t idx_offset, float* __restrict__ out_sim, int64_t* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tmax = nsplit * k;
  uint64_t* sk = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* tmp = sk + tmax;  // [k]
  uint32_t* hist = reinterpret_cast<uint32_t*>(tmp + k);
  uint32_t* scratch = hist + 256;
  const int64_t slot = blockIdx.x;
  const int64_t qi = qlist ? qlist[slot] : slot;
  const int tid = threadIdx.x;
  if (tid == 0) {
    uint32_t off = 0;
    for (int s = 0; s < nsplit; ++s) off += static_cast<uint32_t>(counts[slot * nsplit + s]);
    scratch[3] = off;
  }
  __syncthreads();
  const int total = static_cast<int>(scratch[3]);
  // gather the split lists contiguously (each split holds <= k keys after its final prune)
  int base = 0;
  for (int s = 0; s < nsplit; ++s) {
    const int c = counts[slot * nsplit + s];
    const uint64_t* src = keys + (slot * nsplit + s) * static_cast<int64_t>(cap);
    for (int i = tid; i < c; i += blockDim.x) sk[base + i] = src[i];
    base += c;
  }
  __syncthreads();
  int m = total;
  const uint64_t* fin = sk;
  if (m > k) {
    block_select(sk, m, k, tmp, hist, scratch);
    m = k;
    fin = tmp;
  }
  // rank sort of the m survivors (unique keys)
  for (int j = tid; j < m; j += blockDim.x) {
    const uint64_t mine = fin[j];
    int rank = 0;
    for (int i = 0; i < m; ++i) rank += (fin[i] > mine) ? 1 : 0;
    out_sim[qi * k + rank] = key_sim(mine);
    out_idx[qi * k + rank] = static_cast<int64_t>(key_idx(mine)) + idx_offset;
  }
}

}  // namespace hcir

extern "C" size_t hcir_exact_workspace_bytes(int64_t nlist, int64_t ng, int k, int sm_count) {
  if (nlist <= 0 || ng <= 0 || k <= 0) return 256;
  return hcir::make_exact_plan(nlist, ng, k, sm_count > 0 ? sm_count : 148).bytes;
}

extern "C" int hcir_exact_topk(const float* q_f32, const float* g_f32, int ld, int64_t ng, int k,
                               int64_t idx_offset, const int32_t* qlist, int64_t nlist, float* out_sim,
                               int64_t* out_idx, void* workspace, size_t workspace_bytes, int sm_count,
                               hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(ld > 0 && ld % 64 == 0, "exact_topk: ld=%d must be a positive multiple of 64", ld);
  HCIR_REQUIRE(ng > 0 && ng < (1ll << 31), "exact_topk: ng=%lld out of range", (long long)ng);
  HCIR_REQUIRE(k > 0 && k <= ng, "exact_topk: k=%d must be in [1, ng=%lld]", k, (long long)ng);
  HCIR_REQUIRE(k <= 4096, "exact_topk: k=%d > 4096 unsupported", k);
  HCIR_REQUIRE(nlist >= 0, "exact_topk: nlist=%lld", (long long)nlist);
  HCIR_REQUIRE(q_f32 && g_f32 && out_sim && out_idx && workspace, "exact_topk: null pointer");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (nlist == 0) return HCIR_OK;
  if (sm_count <= 0) sm_count = 148;
  const ExactPlan p = make_exact_plan(nlist, ng, k, sm_count);
  if (workspace_bytes < p.bytes) {
    set_error("exact_topk: workspace %zu < required %zu", workspace_bytes, p.bytes);
    return HCIR_EWORKSPACE;
  }
  int32_t* counts = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + p.counts_off);
  uint64_t* keys = reinterpret_cast<uint64_t*>(static_cast<char*>(workspace) + p.keys_off);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem_a = static_cast<size_t>(kExactWarps) * ld * sizeof(float) + kExactWarps * 256 * sizeof(uint32_t);
  HCIR_REQUIRE(smem_a <= 200 * 1024, "exact_topk: ld=%d too large for the query staging buffer", ld);
  HCIR_CUDA_TRY(cudaFuncSetAttribute(exact_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem_a)));
  dim3 grid_a(static_cast<unsigned>(ceil_div_i64(nlist, kExactWarps)), static_cast<unsigned>(p.nsplit));
  exact_scan_kernel<<<grid_a, kExactWarps * kWarp, smem_a, st>>>(q_f32, g_f32, ld, ng, qlist, nlist, k, p.nsplit,
                                                                 p.cap, counts, keys);
  HCIR_CUDA_TRY(cudaGetLastError());
  const size_t smem_b = (static_cast<size_t>(p.nsplit) * k + k) * sizeof(uint64_t) + (256 + 8) * sizeof(uint32_t);
  HCIR_CUDA_TRY(cudaFuncSetAttribute(exact_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem_b)));
  exact_finalize_kernel<<<static_cast<unsigned>(nlist), 256, smem_b, st>>>(qlist, k, p.nsplit, p.cap, counts, keys,
                                                                          idx_offset, out_sim, out_idx);
  HCIR_CUDA_TRY(cudaGetLastError());
  return HCIR_OK;
}
