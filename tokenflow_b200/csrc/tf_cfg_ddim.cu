// tf_cfg_ddim — classifier-free guidance + DDIM update of one denoising step in ONE pass
// (reference run_tokenflow_pnp.py:213-217: `noise_pred_uncond + g * (noise_pred_cond - noise_pred_uncond)`
// followed by `scheduler.step(noise_pred, t, x)['prev_sample']`, eta = 0).
//
// The reference runs these as ~8 fp16 elementwise launches; each rounds its result to fp16.  This kernel
// keeps that rounding sequence (every intermediate is rounded to fp16 exactly where the eager expression
// rounds), so the fused step is bit-identical to the eager one:
//     d  = h(c - u)            m  = h(g * d)            e  = h(u + m)                 (CFG)
//     a  = h(s1 * e)           b  = h(x - a)            p  = h(b * inv_s2)            (pred_x0; ATen divides by a
//     c1 = h(s3 * p)           c2 = h(s4 * e)           out = h(c1 + c2)               host scalar as x * (1/s))
// with s1 = sqrt(1 - alpha_t), inv_s2 = 1 / sqrt(alpha_t), s3 = sqrt(alpha_prev), s4 = sqrt(1 - alpha_prev)
// as fp32 values.  The four step coefficients are read from DEVICE memory so that a CUDA graph of the step
// can be replayed for every timestep (the caller copies the step's row of its coefficient table into the
// 4-float buffer before the replay).
//
// HBM-bound and tiny (3 reads + 1 write of the latents, 0.65 MB each at C2): 8 halves per thread.
#include "tf_common.cuh"
#include "tf_kernels.h"

namespace tf {
namespace {

__device__ __forceinline__ float rh(float x) { return __half2float(__float2half_rn(x)); }

__global__ void __launch_bounds__(256)
cfg_ddim_kernel(const __half* __restrict__ eu, const __half* __restrict__ ec, const __half* __restrict__ x,
                const float* __restrict__ coef, float g, long long n_vec, long long n, __half* __restrict__ out) {
  const float s1 = coef[0], inv_s2 = coef[1], s3 = coef[2], s4 = coef[3];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto one = [&](float u, float c, float xv) -> float {
    const float e = rh(u + rh(g * rh(c - u)));
    const float p = rh(rh(xv - rh(s1 * e)) * inv_s2);
    return rh(rh(s3 * p) + rh(s4 * e));
  };
  if (i < n_vec) {
    const uint4 ru = reinterpret_cast<const uint4*>(eu)[i];
    const uint4 rc = reinterpret_cast<const uint4*>(ec)[i];
    const uint4 rx = reinterpret_cast<const uint4*>(x)[i];
    const __half2* hu = reinterpret_cast<const __half2*>(&ru);
    const __half2* hc = reinterpret_cast<const __half2*>(&rc);
    const __half2* hx = reinterpret_cast<const __half2*>(&rx);
    uint4 w;
    __half2* ho = reinterpret_cast<__half2*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 u = __half22float2(hu[e]), c = __half22float2(hc[e]), xv = __half22float2(hx[e]);
      ho[e] = __floats2half2_rn(one(u.x, c.x, xv.x), one(u.y, c.y, xv.y));
    }
    reinterpret_cast<uint4*>(out)[i] = w;
  }
  if (i == 0) {                                   // tail (n not a multiple of 8)
    for (long long j = n_vec * 8; j < n; ++j)
      out[j] = __float2half_rn(one(__half2float(eu[j]), __half2float(ec[j]), __half2float(x[j])));
  }
}

}  // namespace

int launch_cfg_ddim(const void* eps_uncond, const void* eps_cond, const void* x, const float* coef_dev, float guidance,
                    long long n, void* out, cudaStream_t stream) {
  if (n == 0) return TF_OK;
  const long long n_vec = n / 8;
  const long long threads = n_vec > 0 ? n_vec : 1;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  cfg_ddim_kernel<<<blocks, 256, 0, stream>>>(static_cast<const __half*>(eps_uncond), static_cast<const __half*>(eps_cond),
                                             static_cast<const __half*>(x), coef_dev, guidance, n_vec, n,
                                             static_cast<__half*>(out));
  return check_cuda(cudaGetLastError(), "tf_cfg_ddim launch");
}

}  // namespace tf
