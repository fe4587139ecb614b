// tf_propagate — NN-indexed feature propagation (reference tokenflow_utils.py:361-397).
//
//   out[s,f,p,:] = w[f] * A[s,kfa[f],idx_a[f,p],:] + (1-w[f]) * A[s,kfb[f],idx_b[f,p],:]  (+ residual[s,f,p,:])
//   (kfb[f] < 0: out = A[s,kfa[f],idx_a[f,p],:] (+ residual))
//
// HBM-bound byte mover.  What the reference does with ~10x the algorithmic traffic (two int64
// [3,B*S,dim] index tensors, an fp32 [3,B,S,dim] weight tensor, 6 elementwise passes; SURVEY.md
// §2.1 k10-k12) is one pass here: each thread owns one 16-byte vector column of a token row, reads the
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

constexpr int kThreads = 320;   // 8 x 40, 4 x 80, 2 x 160 vectors: whole rows for every SD channel width

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

// Thread layout: blockIdx.y = frame; inside a block, threadIdx.x -> (row-in-group, 16-byte vector) with
// the vector index fixed per thread, so the row loop has no integer division and consecutive
// threads cover consecutive 16-byte vectors of consecutive output rows (fully coalesced).
template <bool kOutF32, bool kResidual>
__global__ void __launch_bounds__(kThreads)
propagate_kernel(const __half* __restrict__ A, const int32_t* __restrict__ idx_a,
                 const int32_t* __restrict__ idx_b, FrameTable tab, int F, int S, int dim, int K,
                 const __half* __restrict__ residual, void* __restrict__ out, int nvec, int rows_per_block,
                 long long stream_out) {      // stream_out: elements per stream in out / residual (= F_total*S*dim)
  const int f = blockIdx.y;
  const int r_in = threadIdx.x / nvec;
  const int vec = threadIdx.x - r_in * nvec;
  if (r_in >= rows_per_block) return;
  const int kfa = tab.kf_a[f];
  const int kfb = tab.kf_b[f];
  const float w = tab.w[f];
  const float w2 = 1.0f - w;
  const long long kf_stride = (long long)S * dim;           // elements per keyframe slab
  const long long stream_A = (long long)K * kf_stride;
  const __half* A_a = A + (long long)kfa * kf_stride + vec * 8;
  const __half* A_b = A + (long long)(kfb >= 0 ? kfb : kfa) * kf_stride + vec * 8;
  const int32_t* ia_f = idx_a + (long long)f * S;
  const int32_t* ib_f = idx_b + (long long)f * S;
  const long long frame_off = (long long)f * S * dim + vec * 8;

  for (int p = blockIdx.x * rows_per_block + r_in; p < S; p += gridDim.x * rows_per_block) {
    const int ia = __ldg(ia_f + p);
    const int ib = (kfb >= 0) ? __ldg(ib_f + p) : 0;
    const __half* a_row = A_a + (long long)ia * dim;
    const __half* b_row = A_b + (long long)ib * dim;
    const long long o_off = frame_off + (long long)p * dim;

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

// `F` frames starting at the pointers given (a chunk of at most kMaxFrames frames of a call with `F_total`
// frames: the three streams of out / residual are F_total*S*dim elements apart).
int launch_propagate(const void* A, const int32_t* idx_a, const int32_t* idx_b, const FrameTable& tab, int F,
                     int S, int dim, int K, const void* residual, void* out, int out_is_f32, long long F_total,
                     cudaStream_t stream) {
  if ((long long)F * S == 0) return TF_OK;
  const int nvec = dim >> 3;
  if (nvec > kThreads) {
    set_last_error("tf_propagate: dim=%d > %d is not supported", dim, kThreads * 8);
    return TF_ERR_UNSUPPORTED;
  }
  const int rows_per_block = kThreads / nvec;
  // about one wave of resident CTAs in total (6 CTAs of 320 threads per SM), split evenly over frames
  int bx = (S + rows_per_block - 1) / rows_per_block;
  const int cap = (sm_count() * 6 + F - 1) / F;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)F), block(kThreads);
  const __half* Ah = static_cast<const __half*>(A);
  const __half* Rh = static_cast<const __half*>(residual);
#define TF_LAUNCH(OUTF32, RES)                                                                         \
  propagate_kernel<OUTF32, RES><<<grid, block, 0, stream>>>(Ah, idx_a, idx_b, tab, F, S, dim, K, Rh, out, nvec, \
                                                            rows_per_block, F_total * (long long)S * dim)
  if (out_is_f32) {
    if (residual) TF_LAUNCH(true, true); else TF_LAUNCH(true, false);
  } else {
    if (residual) TF_LAUNCH(false, true); else TF_LAUNCH(false, false);
  }
#undef TF_LAUNCH
  return check_cuda(cudaGetLastError(), "tf_propagate launch");
}

}  // namespace tf
