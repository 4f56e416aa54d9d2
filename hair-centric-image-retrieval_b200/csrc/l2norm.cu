// K1: fused row L2-normalise + fp32/bf16 emit + bf16 rounding-error norm.
// HBM-bound: algorithmic bytes per row = 4*d read + (4 + 2)*ld written (+4 for delta).
// One warp per row, 128-bit coalesced loads; the second pass re-reads the row from L1/L2
// (a 2-8 KB row just touched by the same warp), so DRAM sees each input byte once.
#include "hcir_common.cuh"

namespace hcir {

constexpr int kNormWarpsPerBlock = 8;

__global__ void __launch_bounds__(kNormWarpsPerBlock* kWarp)
l2norm_cast_kernel(const float* __restrict__ x, int64_t n, int d, int64_t ldx,
                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int ld,
                   float* __restrict__ out_delta, bool vec_ok) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * kNormWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * kNormWarpsPerBlock;
  const int d4 = d >> 2;
  const int ld4 = ld >> 2;
  for (int64_t row = warp_global; row < n; row += warp_stride) {
    const float* xr = x + row * ldx;
    // ---- pass 1: sum of squares
    float ss = 0.0f;
    if (vec_ok) {
      const float4* x4 = reinterpret_cast<const float4*>(xr);
      for (int c = lane; c < d4; c += kWarp) {
        const float4 v = __ldg(x4 + c);
        ss = fmaf(v.x, v.x, ss);
        ss = fmaf(v.y, v.y, ss);
        ss = fmaf(v.z, v.z, ss);
        ss = fmaf(v.w, v.w, ss);
      }
      for (int c = (d4 << 2) + lane; c < d; c += kWarp) ss = fmaf(xr[c], xr[c], ss);
    } else {
      for (int c = lane; c < d; c += kWarp) ss = fmaf(xr[c], xr[c], ss);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
    // F.normalize: x / max(||x||, eps), eps = 1e-12
    const float den = fmaxf(sqrtf(ss), 1e-12f);
    // ---- pass 2: scale, emit, accumulate the bf16 rounding error
    float err = 0.0f;
    float* of = out_f32 ? out_f32 + row * static_cast<int64_t>(ld) : nullptr;
    __nv_bfloat16* ob = out_bf16 ? out_bf16 + row * static_cast<int64_t>(ld) : nullptr;
    for (int c = lane; c < ld4; c += kWarp) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int e0 = c << 2;
      if (vec_ok && e0 + 3 < d) {
        v = __ldg(reinterpret_cast<const float4*>(xr) + c);
      } else {
        if (e0 + 0 < d) v.x = xr[e0 + 0];
        if (e0 + 1 < d) v.y = xr[e0 + 1];
        if (e0 + 2 < d) v.z = xr[e0 + 2];
        if (e0 + 3 < d) v.w = xr[e0 + 3];
      }
      v.x /= den; v.y /= den; v.z /= den; v.w /= den;
      if (of) reinterpret_cast<float4*>(of)[c] = v;
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
      if (ob) {
        uint2 packed;
        packed.x = *reinterpret_cast<const uint32_t*>(&lo);
        packed.y = *reinterpret_cast<const uint32_t*>(&hi);
        reinterpret_cast<uint2*>(ob)[c] = packed;
      }
      const float ex = v.x - __bfloat162float(lo.x), ey = v.y - __bfloat162float(lo.y);
      const float ez = v.z - __bfloat162float(hi.x), ew = v.w - __bfloat162float(hi.y);
      err = fmaf(ex, ex, err);
      err = fmaf(ey, ey, err);
      err = fmaf(ez, ez, err);
      err = fmaf(ew, ew, err);
    }
    if (out_delta) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) err += __shfl_xor_sync(kFull, err, o);
      // round the bound UP a little: it is used as an upper bound
      if (lane == 0) out_delta[row] = sqrtf(err) * 1.0001f + 1e-12f;
    }
  }
}

}  // namespace hcir

extern "C" int hcir_l2norm_cast(const float* x, int64_t n, int d, int64_t ldx, float* out_f32,
                                uint16_t* out_bf16, int ld, float* out_delta, hcir_stream_t stream) {
  using namespace hcir;
  HCIR_REQUIRE(n >= 0 && d > 0, "l2norm_cast: bad shape n=%lld d=%d", (long long)n, d);
  HCIR_REQUIRE(ldx >= d, "l2norm_cast: ldx=%lld < d=%d", (long long)ldx, d);
  HCIR_REQUIRE(ld == hcir_padded_dim(d), "l2norm_cast: ld=%d must equal hcir_padded_dim(%d)=%d", ld, d,
               hcir_padded_dim(d));
  HCIR_REQUIRE(x != nullptr || n == 0, "l2norm_cast: null input");
  int rc = check_device();
  if (rc != HCIR_OK) return rc;
  if (n == 0) return HCIR_OK;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (ldx % 4 == 0);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = ceil_div_i64(n, kNormWarpsPerBlock);
  const int64_t cap = static_cast<int64_t>(sms) * 16;  // 16 resident CTAs of 8 warps per SM, grid-stride
  const int grid = static_cast<int>(want < cap ? want : cap);
  HCIR_CUDA_TRY(launch_pdl(l2norm_cast_kernel, dim3(grid), dim3(kNormWarpsPerBlock * kWarp), 0,
                           static_cast<cudaStream_t>(stream), x, n, d, ldx, out_f32,
                           reinterpret_cast<__nv_bfloat16*>(out_bf16), ld, out_delta, vec_ok));
  return HCIR_OK;
}
