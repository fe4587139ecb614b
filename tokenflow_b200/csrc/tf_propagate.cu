// tf_propagate — NN-indexed feature propagation (reference tokenflow_utils.py:361-397).
//
//   out[s,f,p,:] = w[f] * A[s,kfa[f],idx_a[f,p],:] + (1-w[f]) * A[s,kfb[f],idx_b[f,p],:]  (+ residual[s,f,p,:])
//   (kfb[f] < 0: out = A[s,kfa[f],idx_a[f,p],:] (+ residual))
//
// HBM-bound byte mover.  What the reference does with ~10x the algorithmic traffic (two int64
// [3,B*S,dim] index tensors, an fp32 [3,B,S,dim] weight tensor, 6 elementwise passes; SURVEY.md
// §2.1 k10-k12) is one pass here: each thread owns one 16-byte vector of one token row, reads the
// two int32 NN indices once and reuses them for the three streams, issues all (up to 9) independent
// 16-byte loads before the first use, blends in fp32 and stores the row once.  Consecutive threads
// own consecutive vectors of the contiguous [f,p,dim] output, so stores and residual loads are
// fully coalesced; gathered keyframe rows are dim*2-byte contiguous segments (>= 128 B for every
// SD level) whose re-touches hit L2 (a keyframe slab is <= 8 MB).
//
// Algorithmic bytes: write 3*F*S*dim*esz_out + residual read 3*F*S*dim*2 + unique
// keyframe rows 3*S*dim*2 per referenced keyframe + indices 4*S*F*(1 or 2).
#include "tf_common.cuh"
#include "tf_kernels.h"

namespace tf {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {   // read-once data: keep it out of L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_reuse(const uint4* p) {    // gathered keyframe rows: cacheable
  return __ldg(p);
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

template <bool kOutF32, bool kResidual>
__global__ void __launch_bounds__(kThreads)
propagate_kernel(const __half* __restrict__ A, const int32_t* __restrict__ idx_a,
                 const int32_t* __restrict__ idx_b, FrameTable tab, int F, int S, int dim, int K,
                 const __half* __restrict__ residual, void* __restrict__ out) {
  const int nvec = dim >> 3;                               // 16-byte vectors per row
  const long long rows = (long long)F * S;
  const long long total = rows * nvec;
  const long long stream_out = rows * dim;                 // elements per stream in out / residual
  const long long kf_stride = (long long)S * dim;          // elements per keyframe slab
  const long long stream_A = (long long)K * kf_stride;

  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kThreads) {
    const long long row = i / nvec;
    const int vec = (int)(i - row * nvec);
    const int f = (int)(row / S);
    const int kfa = tab.kf_a[f];
    const int kfb = tab.kf_b[f];
    const float w = tab.w[f];
    const int ia = __ldg(idx_a + row);
    const int ib = (kfb >= 0) ? __ldg(idx_b + row) : 0;

    const __half* a_row = A + (long long)kfa * kf_stride + (long long)ia * dim + vec * 8;
    const __half* b_row = A + (long long)(kfb >= 0 ? kfb : kfa) * kf_stride + (long long)ib * dim + vec * 8;
    const long long o_off = row * dim + vec * 8;

    uint4 va[3], vb[3], vr[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      va[s] = ld_reuse(reinterpret_cast<const uint4*>(a_row + s * stream_A));
      if (kfb >= 0) vb[s] = ld_reuse(reinterpret_cast<const uint4*>(b_row + s * stream_A));
      if (kResidual) vr[s] = ld_stream(reinterpret_cast<const uint4*>(residual + s * stream_out + o_off));
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      float a[8], r[8];
      unpack8(va[s], a);
      if (kfb >= 0) {
        float b[8];
        unpack8(vb[s], b);
        const float w2 = 1.0f - w;
#pragma unroll
        for (int e = 0; e < 8; ++e)   // reference :388: two fp32 products then an fp32 add (no FMA contraction)
          a[e] = __fadd_rn(__fmul_rn(w, a[e]), __fmul_rn(w2, b[e]));
      }
      if (kResidual) {
        unpack8(vr[s], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] += r[e];
      }
      if (kOutF32) {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + s * stream_out + o_off);
        o[0] = make_float4(a[0], a[1], a[2], a[3]);
        o[1] = make_float4(a[4], a[5], a[6], a[7]);
      } else {
        st_stream(reinterpret_cast<uint4*>(reinterpret_cast<__half*>(out) + s * stream_out + o_off), pack8(a));
      }
    }
  }
}

}  // namespace

int launch_propagate(const void* A, const int32_t* idx_a, const int32_t* idx_b, const FrameTable& tab, int F,
                     int S, int dim, int K, const void* residual, void* out, int out_is_f32,
                     cudaStream_t stream) {
  const long long total = (long long)F * S * (dim >> 3);
  if (total == 0) return TF_OK;
  // one wave of resident CTAs (8 x 256 threads per SM), grid-stride over the rest
  long long blocks = (total + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks), block(kThreads);
  const __half* Ah = static_cast<const __half*>(A);
  const __half* Rh = static_cast<const __half*>(residual);
#define TF_LAUNCH(OUTF32, RES)                                                                         \
  propagate_kernel<OUTF32, RES><<<grid, block, 0, stream>>>(Ah, idx_a, idx_b, tab, F, S, dim, K, Rh, out)
  if (out_is_f32) {
    if (residual) TF_LAUNCH(true, true); else TF_LAUNCH(true, false);
  } else {
    if (residual) TF_LAUNCH(false, true); else TF_LAUNCH(false, false);
  }
#undef TF_LAUNCH
  return check_cuda(cudaGetLastError(), "tf_propagate launch");
}

}  // namespace tf
