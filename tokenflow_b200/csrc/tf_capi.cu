// C-ABI layer (include/tokenflow_b200.h): argument validation, per-frame tables, error plumbing.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/tokenflow_b200.h"
#include "tf_common.cuh"
#include "tf_kernels.h"

namespace tf {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return TF_OK;
  set_last_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return TF_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swizzle) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p)
      return CUDA_ERROR_NOT_FOUND;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstrides[i] = strides_bytes[i];
  }
  return fn(map, dtype, rank, const_cast<void*>(base), gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int fill_table(FrameTable& tab, const int32_t* kf_a, const int32_t* kf_b, const float* w, int F, int K,
                      const char* who) {
  if (F < 0 || F > kMaxFrames) { set_last_error("%s: F=%d outside [0,%d]", who, F, kMaxFrames); return TF_ERR_INVALID_ARGUMENT; }
  if (F > 0 && !kf_a) { set_last_error("%s: kf_a is NULL", who); return TF_ERR_INVALID_ARGUMENT; }
  memset(&tab, 0, sizeof(tab));
  for (int f = 0; f < F; ++f) {
    const int a = kf_a[f];
    const int b = kf_b ? kf_b[f] : -1;
    if (a < 0 || a >= K || b >= K) {
      set_last_error("%s: frame %d has keyframe ids (%d,%d) outside [0,%d)", who, f, a, b, K);
      return TF_ERR_INVALID_ARGUMENT;
    }
    tab.kf_a[f] = a;
    tab.kf_b[f] = b < 0 ? -1 : b;
    tab.w[f] = w ? w[f] : 1.0f;
  }
  return TF_OK;
}

}  // namespace tf

using namespace tf;

extern "C" {

int tf_version(void) { return 1000; }

const char* tf_last_error(void) { return g_err; }

int64_t tf_launch_count(void) { return (int64_t)g_launches.load(); }

int tf_unit_rows(const void* x, int x_is_f32, int64_t rows, int dim, int64_t x_row_stride, void* out_f16,
                 tf_stream_t stream) {
  if (rows < 0 || dim <= 0 || (dim & 7) || (x_row_stride & 3) || x_row_stride < dim) {
    set_last_error("tf_unit_rows: bad shape rows=%lld dim=%d stride=%lld (dim %% 8 == 0 required)", (long long)rows,
                   dim, (long long)x_row_stride);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (rows == 0) return TF_OK;
  if (!x || !out_f16 || !aligned16(x) || !aligned16(out_f16)) {
    set_last_error("tf_unit_rows: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  int e = launch_unit_rows(x, x_is_f32, rows, dim, x_row_stride, out_f16, static_cast<cudaStream_t>(stream));
  if (!e) g_launches += 1;
  return e;
}

int tf_layernorm_unit_rows(const void* x_f16, int64_t rows, int dim, int64_t x_row_stride, const float* gamma,
                           const float* beta, float eps, void* out_f16, tf_stream_t stream) {
  if (rows < 0 || dim <= 0 || (dim & 7) || (x_row_stride & 7) || x_row_stride < dim) {
    set_last_error("tf_layernorm_unit_rows: bad shape rows=%lld dim=%d stride=%lld (dim %% 8 == 0 required)",
                   (long long)rows, dim, (long long)x_row_stride);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (rows == 0) return TF_OK;
  if (!x_f16 || !gamma || !beta || !out_f16 || !aligned16(x_f16) || !aligned16(out_f16) || !aligned16(gamma) ||
      !aligned16(beta)) {
    set_last_error("tf_layernorm_unit_rows: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  int e = launch_layernorm_unit_rows(x_f16, rows, dim, x_row_stride, gamma, beta, eps, out_f16,
                                     static_cast<cudaStream_t>(stream));
  if (!e) g_launches += 1;
  return e;
}

int tf_layernorm_rows(const void* x_f16, int64_t rows, int dim, int64_t x_row_stride, const float* gamma,
                      const float* beta, float eps, void* y_out_f16, int64_t y_row_stride, void* unit_out_f16,
                      int64_t unit_row_stride, int64_t unit_rows, tf_stream_t stream) {
  if (rows < 0 || dim <= 0 || (dim & 7) || (x_row_stride & 7) || x_row_stride < dim ||
      (y_out_f16 && ((y_row_stride & 7) || y_row_stride < dim)) ||
      (unit_out_f16 && ((unit_row_stride & 7) || unit_row_stride < dim)) || unit_rows < 0) {
    set_last_error("tf_layernorm_rows: bad shape rows=%lld dim=%d strides=(%lld,%lld,%lld) (multiples of 8 required)",
                   (long long)rows, dim, (long long)x_row_stride, (long long)y_row_stride, (long long)unit_row_stride);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (rows == 0) return TF_OK;
  if (!x_f16 || !gamma || !beta || !aligned16(x_f16) || !aligned16(gamma) || !aligned16(beta) ||
      (y_out_f16 && !aligned16(y_out_f16)) || (unit_out_f16 && !aligned16(unit_out_f16))) {
    set_last_error("tf_layernorm_rows: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (!y_out_f16 && (!unit_out_f16 || unit_rows == 0)) return TF_OK;
  int e = launch_layernorm_rows(x_f16, rows, dim, x_row_stride, gamma, beta, eps, y_out_f16, y_row_stride,
                                unit_out_f16, unit_row_stride, unit_rows, static_cast<cudaStream_t>(stream));
  if (!e) g_launches += 1;
  return e;
}

int tf_cfg_ddim(const void* eps_uncond, const void* eps_cond, const void* x, const float* coef, float guidance,
                int64_t n, void* out, tf_stream_t stream) {
  if (n < 0) { set_last_error("tf_cfg_ddim: n=%lld", (long long)n); return TF_ERR_INVALID_ARGUMENT; }
  if (n == 0) return TF_OK;
  if (!eps_uncond || !eps_cond || !x || !coef || !out || !aligned16(eps_uncond) || !aligned16(eps_cond) ||
      !aligned16(x) || !aligned16(out)) {
    set_last_error("tf_cfg_ddim: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  int e = launch_cfg_ddim(eps_uncond, eps_cond, x, coef, guidance, n, out, static_cast<cudaStream_t>(stream));
  if (!e) g_launches += 1;
  return e;
}

int tf_nn_field(const void* x_unit, const void* piv_unit, const int32_t* kf_a, const int32_t* kf_b, int F, int S,
                int dim, int K, int32_t* idx_a, int32_t* idx_b, tf_stream_t stream) {
  if (F < 0 || S < 0 || dim <= 0 || (dim & 7) || K <= 0) {
    set_last_error("tf_nn_field: bad shape F=%d S=%d dim=%d K=%d", F, S, dim, K);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (F > 0 && !kf_a) { set_last_error("tf_nn_field: kf_a is NULL"); return TF_ERR_INVALID_ARGUMENT; }
  // validate every chunk before the first launch
  for (int f0 = 0; f0 < F; f0 += kMaxFrames) {
    FrameTable tab;
    const int fc = F - f0 < kMaxFrames ? F - f0 : kMaxFrames;
    if (int e = fill_table(tab, kf_a + f0, kf_b ? kf_b + f0 : nullptr, nullptr, fc, K, "tf_nn_field")) return e;
  }
  if (F == 0 || S == 0) return TF_OK;
  bool need_b = false;
  if (kf_b) for (int f = 0; f < F; ++f) need_b |= kf_b[f] >= 0;
  if (!x_unit || !piv_unit || !idx_a || (need_b && !idx_b) || !aligned16(x_unit) || !aligned16(piv_unit)) {
    set_last_error("tf_nn_field: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  const __half* xu = static_cast<const __half*>(x_unit);
  for (int f0 = 0; f0 < F; f0 += kMaxFrames) {         // any number of frames: kMaxFrames per launch
    FrameTable tab;
    const int fc = F - f0 < kMaxFrames ? F - f0 : kMaxFrames;
    fill_table(tab, kf_a + f0, kf_b ? kf_b + f0 : nullptr, nullptr, fc, K, "tf_nn_field");
    const long long off = (long long)f0 * S;
    int e = launch_nn_field(xu + off * dim, piv_unit, tab, fc, S, dim, K, idx_a + off, idx_b ? idx_b + off : nullptr,
                            static_cast<cudaStream_t>(stream));
    if (e) return e;
    g_launches += 1;
  }
  return TF_OK;
}

int tf_propagate(const void* A, const int32_t* idx_a, const int32_t* idx_b, const int32_t* kf_a,
                 const int32_t* kf_b, const float* w, int F, int S, int dim, int K, const void* residual,
                 void* out, int out_is_f32, tf_stream_t stream) {
  if (F < 0 || S < 0 || dim <= 0 || (dim & 7) || K <= 0) {
    set_last_error("tf_propagate: bad shape F=%d S=%d dim=%d K=%d", F, S, dim, K);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (F > 0 && !kf_a) { set_last_error("tf_propagate: kf_a is NULL"); return TF_ERR_INVALID_ARGUMENT; }
  bool need_b = false;
  for (int f0 = 0; f0 < F; f0 += kMaxFrames) {
    FrameTable tab;
    const int fc = F - f0 < kMaxFrames ? F - f0 : kMaxFrames;
    if (int e = fill_table(tab, kf_a + f0, kf_b ? kf_b + f0 : nullptr, w ? w + f0 : nullptr, fc, K, "tf_propagate"))
      return e;
    for (int f = 0; f < fc; ++f) need_b |= tab.kf_b[f] >= 0;
  }
  if (F == 0 || S == 0) return TF_OK;
  if (!A || !idx_a || !out || (need_b && (!idx_b || !w)) || !aligned16(A) || !aligned16(out) ||
      (residual && !aligned16(residual))) {
    set_last_error("tf_propagate: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  const size_t out_esz = out_is_f32 ? 4 : 2;
  for (int f0 = 0; f0 < F; f0 += kMaxFrames) {         // any number of frames: kMaxFrames per launch, no copies
    FrameTable tab;
    const int fc = F - f0 < kMaxFrames ? F - f0 : kMaxFrames;
    fill_table(tab, kf_a + f0, kf_b ? kf_b + f0 : nullptr, w ? w + f0 : nullptr, fc, K, "tf_propagate");
    const long long off = (long long)f0 * S;
    int e = launch_propagate(A, idx_a + off, idx_b ? idx_b + off : nullptr, tab, fc, S, dim, K,
                             residual ? static_cast<const __half*>(residual) + off * dim : nullptr,
                             static_cast<char*>(out) + (size_t)off * dim * out_esz, out_is_f32, F,
                             static_cast<cudaStream_t>(stream));
    if (e) return e;
    g_launches += 1;
  }
  return TF_OK;
}

int tf_ext_attn_fwd_rows(const void* q, int q_slabs, int64_t q_tok_stride, const void* k, const void* v,
                         int kv_slabs, int64_t kv_tok_stride, int n_out, const int32_t* out_slab,
                         const int32_t* q_slab, const int32_t* k_slab0, const int32_t* v_slab0,
                         const int32_t* n_kv, int S, int heads, int d, float scale, int q_row0, int q_nrows,
                         void* out, tf_stream_t stream) {
  if (q_row0 < 0 || q_nrows < 0 || (q_row0 & 127)) {
    set_last_error("tf_ext_attn: bad query row range [%d, +%d) (start must be a multiple of 128)", q_row0, q_nrows);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (n_out < 0 || S < 0 || heads <= 0 || d <= 0 || (d & 7) ||
      q_tok_stride < (int64_t)heads * d || kv_tok_stride < (int64_t)heads * d || (q_tok_stride & 7) ||
      (kv_tok_stride & 7)) {
    set_last_error("tf_ext_attn: bad shape n_out=%d S=%d heads=%d d=%d strides=(%lld,%lld)", n_out, S, heads, d,
                   (long long)q_tok_stride, (long long)kv_tok_stride);
    return TF_ERR_INVALID_ARGUMENT;
  }
  if (n_out == 0 || S == 0 || q_nrows == 0 || q_row0 >= S) return TF_OK;
  if (!q || !k || !v || !out || !out_slab || !q_slab || !k_slab0 || !v_slab0 || !n_kv || !aligned16(q) ||
      !aligned16(k) || !aligned16(v) || !aligned16(out)) {
    set_last_error("tf_ext_attn: NULL or misaligned pointer");
    return TF_ERR_INVALID_ARGUMENT;
  }
  for (int i = 0; i < n_out; ++i) {
    if (q_slab[i] < 0 || q_slab[i] >= q_slabs || n_kv[i] <= 0 || k_slab0[i] < 0 || v_slab0[i] < 0 ||
        k_slab0[i] + n_kv[i] > kv_slabs || v_slab0[i] + n_kv[i] > kv_slabs || out_slab[i] < 0) {
      set_last_error("tf_ext_attn: sample %d has an out-of-range slab (q=%d k0=%d v0=%d n_kv=%d)", i, q_slab[i],
                     k_slab0[i], v_slab0[i], n_kv[i]);
      return TF_ERR_INVALID_ARGUMENT;
    }
  }
  // Samples that share q and k (PnP q/k injection: the uncond and cond sample of a keyframe, reference :124-130) have
  // identical scores and probabilities: pair them and let one kernel compute S / P once and P [V_a | V_b] together.
  std::vector<int> partner(n_out, -1);
  int n_pairs = 0;
  const int rows_here = (q_row0 + q_nrows < S ? q_row0 + q_nrows : S) - q_row0;
  if (ext_attn_pairs_supported(rows_here, d)) {
    for (int i = 0; i < n_out; ++i) {
      if (partner[i] >= 0) continue;
      for (int j = i + 1; j < n_out; ++j) {
        if (partner[j] < 0 && q_slab[j] == q_slab[i] && k_slab0[j] == k_slab0[i] && n_kv[j] == n_kv[i] &&
            v_slab0[j] != v_slab0[i] && out_slab[j] != out_slab[i]) {
          partner[i] = j;
          partner[j] = i;
          ++n_pairs;
          break;
        }
      }
    }
  }
  if (n_pairs > 0) {
    std::vector<int> lead;
    for (int i = 0; i < n_out; ++i)
      if (partner[i] > i) lead.push_back(i);
    std::stable_sort(lead.begin(), lead.end(), [&](int a, int b) { return n_kv[a] > n_kv[b]; });
    for (int p0 = 0; p0 < n_pairs; p0 += kMaxAttnPairs) {
      const int nc = n_pairs - p0 < kMaxAttnPairs ? n_pairs - p0 : kMaxAttnPairs;
      AttnPairTable ptab;
      memset(&ptab, 0, sizeof(ptab));
      for (int slot = 0; slot < nc; ++slot) {
        const int i = lead[p0 + slot], j = partner[i];
        ptab.p[slot].out_u = out_slab[i];
        ptab.p[slot].out_c = out_slab[j];
        ptab.p[slot].q_sample = q_slab[i];
        ptab.p[slot].k_sample0 = k_slab0[i];
        ptab.p[slot].v_u0 = v_slab0[i];
        ptab.p[slot].v_c0 = v_slab0[j];
        ptab.p[slot].n_kv = n_kv[i];
      }
      int e = launch_ext_attn_pairs(q, k, v, q_tok_stride, kv_tok_stride, q_slabs, kv_slabs, ptab, nc, S, heads, d, scale,
                                    out, q_row0, q_nrows, static_cast<cudaStream_t>(stream));
      if (e) return e;
      g_launches += 1;
    }
  }
  // heavy samples (most key slabs) first: the hardware block scheduler then fills the tail of the
  // grid with the cheap own-frame (source stream) samples
  std::vector<int> order;
  for (int i = 0; i < n_out; ++i)
    if (partner[i] < 0) order.push_back(i);
  const int n_single = (int)order.size();
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n_kv[a] > n_kv[b]; });
  for (int s0 = 0; s0 < n_single; s0 += kMaxAttnSamples) {    // any number of samples: kMaxAttnSamples per launch
    const int nc = n_single - s0 < kMaxAttnSamples ? n_single - s0 : kMaxAttnSamples;
    AttnTable tab;
    memset(&tab, 0, sizeof(tab));
    for (int slot = 0; slot < nc; ++slot) {
      const int i = order[s0 + slot];
      tab.s[slot].out_sample = out_slab[i];
      tab.s[slot].q_sample = q_slab[i];
      tab.s[slot].k_sample0 = k_slab0[i];
      tab.s[slot].v_sample0 = v_slab0[i];
      tab.s[slot].n_kv = n_kv[i];
    }
    int e = launch_ext_attn(q, k, v, q_tok_stride, kv_tok_stride, q_slabs, kv_slabs, tab, nc, S, heads, d, scale, out,
                            q_row0, q_nrows, static_cast<cudaStream_t>(stream));
    if (e) return e;
    g_launches += 1;
  }
  return TF_OK;
}

int tf_ext_attn_fwd_table(const void* q, int q_slabs, int64_t q_tok_stride, const void* k, const void* v,
                          int kv_slabs, int64_t kv_tok_stride, int n_out, const int32_t* out_slab,
                          const int32_t* q_slab, const int32_t* k_slab0, const int32_t* v_slab0,
                          const int32_t* n_kv, int S, int heads, int d, float scale, void* out,
                          tf_stream_t stream) {
  return tf_ext_attn_fwd_rows(q, q_slabs, q_tok_stride, k, v, kv_slabs, kv_tok_stride, n_out, out_slab, q_slab, k_slab0,
                              v_slab0, n_kv, S, heads, d, scale, 0, S, out, stream);
}

int tf_ext_attn_fwd(const void* q, const void* k, const void* v, int64_t tok_stride, int n_frames, int S,
                    int heads, int d, float scale, int inject, void* out, tf_stream_t stream) {
  const int n = n_frames;
  if (n < 0) {
    set_last_error("tf_ext_attn_fwd: n_frames=%d", n);
    return TF_ERR_INVALID_ARGUMENT;
  }
  std::vector<int32_t> out_slab(3 * n), q_slab(3 * n), k0(3 * n), v0(3 * n), nkv(3 * n);
  for (int s = 0; s < 3; ++s) {
    for (int f = 0; f < n; ++f) {
      const int i = s * n + f;
      out_slab[i] = i;
      if (s == 0) {                       // source stream: own frame only (reference :173,:177)
        q_slab[i] = f; k0[i] = f; v0[i] = f; nkv[i] = 1;
      } else {                            // uncond / cond: all n frames of the stream (:133-138)
        q_slab[i] = inject ? f : i;       // injection (:126,:129): q of the source stream
        k0[i] = inject ? 0 : s * n;       // injection (:127,:130): k of the source stream
        v0[i] = s * n;                    // v is never injected
        nkv[i] = n;
      }
    }
  }
  return tf_ext_attn_fwd_table(q, 3 * n, tok_stride, k, v, 3 * n, tok_stride, 3 * n, out_slab.data(), q_slab.data(),
                               k0.data(), v0.data(), nkv.data(), S, heads, d, scale, out, stream);
}

}  // extern "C"
