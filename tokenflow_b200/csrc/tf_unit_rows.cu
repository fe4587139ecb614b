// tf_unit_rows — row L2-normalisation feeding the NN field (reference util.py:66-67).
//
//   out[r,:] = fp16_rne( x[r,:] / ||x[r,:]||_2 )        x fp32 (the autocast dtype of norm1's output)
//                                                        or fp16; no epsilon, like the reference.
// The reference normalises in fp32 and lets autocast round the GEMM operands to fp16 (SURVEY.md
// Appendix A "GPU dtype flow"); this kernel produces exactly those fp16 operands once, so the
// similarity GEMM reads 2 bytes/element and the keyframe side is normalised once per pivotal pass
// instead of once per frame batch.  The squared norm is accumulated in fp64 (the correctly rounded
// value a 1-ulp-accurate fp32 reduction approximates), the quotient is an IEEE fp32 division.
//
// One warp per row, 16-byte loads, HBM-bound: rows*dim*(esz_in + 2) bytes.
#include "tf_common.cuh"
#include "tf_kernels.h"

namespace tf {
namespace {

constexpr int kWarpsPerBlock = 8;

template <bool kInF32>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
unit_rows_kernel(const void* __restrict__ x, long long rows, int dim, long long row_stride,
                 __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kWarpsPerBlock;
  for (long long r = warp0; r < rows; r += nwarps) {
    double ss = 0.0;
    if (kInF32) {
      const float* xr = static_cast<const float*>(x) + r * row_stride;
      for (int c = lane * 4; c < dim; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(xr + c);
        ss += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
      }
    } else {
      const __half* xr = static_cast<const __half*>(x) + r * row_stride;
      for (int c = lane * 4; c < dim; c += 128) {
        uint2 raw = *reinterpret_cast<const uint2*>(xr + c);
        float2 a = __half22float2(*reinterpret_cast<__half2*>(&raw.x));
        float2 b = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
        ss += (double)a.x * a.x + (double)a.y * a.y + (double)b.x * b.x + (double)b.y * b.y;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = (float)sqrt(ss);
    __half* orow = out + r * dim;
    if (kInF32) {
      const float* xr = static_cast<const float*>(x) + r * row_stride;
      for (int c = lane * 4; c < dim; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(xr + c);
        __half2 lo = __floats2half2_rn(__fdiv_rn(v.x, nrm), __fdiv_rn(v.y, nrm));
        __half2 hi = __floats2half2_rn(__fdiv_rn(v.z, nrm), __fdiv_rn(v.w, nrm));
        uint2 o2;
        o2.x = *reinterpret_cast<uint32_t*>(&lo);
        o2.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(orow + c) = o2;
      }
    } else {
      const __half* xr = static_cast<const __half*>(x) + r * row_stride;
      for (int c = lane * 4; c < dim; c += 128) {
        uint2 raw = *reinterpret_cast<const uint2*>(xr + c);
        float2 a = __half22float2(*reinterpret_cast<__half2*>(&raw.x));
        float2 b = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
        __half2 lo = __floats2half2_rn(__fdiv_rn(a.x, nrm), __fdiv_rn(a.y, nrm));
        __half2 hi = __floats2half2_rn(__fdiv_rn(b.x, nrm), __fdiv_rn(b.y, nrm));
        uint2 o2;
        o2.x = *reinterpret_cast<uint32_t*>(&lo);
        o2.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(orow + c) = o2;
      }
    }
  }
}

}  // namespace

int launch_unit_rows(const void* x, int x_is_f32, long long rows, int dim, long long row_stride, void* out_f16,
                     cudaStream_t stream) {
  if (rows == 0) return TF_OK;
  long long blocks = (rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks), block(kWarpsPerBlock * 32);
  if (x_is_f32)
    unit_rows_kernel<true><<<grid, block, 0, stream>>>(x, rows, dim, row_stride, static_cast<__half*>(out_f16));
  else
    unit_rows_kernel<false><<<grid, block, 0, stream>>>(x, rows, dim, row_stride, static_cast<__half*>(out_f16));
  return check_cuda(cudaGetLastError(), "tf_unit_rows launch");
}

}  // namespace tf
