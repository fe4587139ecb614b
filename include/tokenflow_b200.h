/* tokenflow_b200 — C ABI of the B200-native TokenFlow hot path.
 *
 * The reference (omerbt/TokenFlow @ 5dd6a69) is pure Python and has no FFI layer; its boundary for
 * this path is the hook surface of tokenflow_utils.py.  This library is what a replacement for those
 * hooks binds (ctypes stub shown in INTEGRATION.md; `tokenflow_b200/ops.py` is that stub in
 * product form).  Every entry point cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, explicit shapes/strides, an opaque CUDA stream handle
 *     (`cudaStream_t`; pass `torch.cuda.current_stream().cuda_stream`).  No torch types.
 *   - every function returns an int status: 0 = ok, nonzero = error (tf_last_error() explains).
 *     Nothing throws, nothing allocates device memory, nothing synchronises the device: work is
 *     enqueued on `stream` and the caller owns every buffer.
 *   - "host" pointers are read synchronously during the call (small per-frame tables passed to the
 *     kernels by value); "device" pointers must be valid on the current device.
 *   - fp16 activations, int32 indices.  dim and head_dim must be multiples of 8; device pointers
 *     16-byte aligned; tensors contiguous unless a stride argument says otherwise.
 *   - compiled for sm_100a only; the kernels use tcgen05 / TMEM / TMA.
 */
#ifndef TOKENFLOW_B200_H_
#define TOKENFLOW_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tf_stream_t; /* cudaStream_t */

#define TF_MAX_FRAMES 64        /* frames per kernel launch; tf_nn_field / tf_propagate take any F and chunk */
#define TF_MAX_ATTN_SAMPLES 160 /* output samples (stream, keyframe) per kernel launch; calls take any number */

/* Library ABI version (major*1000 + minor). */
int tf_version(void);
/* Thread-local description of the last error returned on this thread ("" if none). */
const char* tf_last_error(void);
/* Number of kernel launches enqueued by this library since load (all threads); bench.py reports
 * the per-step difference as `gpu_launches`. */
int64_t tf_launch_count(void);

/* Row L2-normalisation feeding the NN field:  out[r,:] = fp16(x[r,:] / ||x[r,:]||_2).
 * Replaces util.py:66-67 (`x / x.norm(dim=-1, keepdim=True)`, fp32 under autocast) plus the fp16
 * operand cast autocast applies to the following matmul (util.py:68).
 *   x            device, [rows, dim] fp32 (x_is_f32 != 0) or fp16, row pitch `x_row_stride` elements
 *   out_f16      device, [rows, dim] fp16, contiguous */
int tf_unit_rows(const void* x, int x_is_f32, int64_t rows, int dim, int64_t x_row_stride, void* out_f16,
                 tf_stream_t stream);

/* norm1 fused with the row normalisation, for the frame pass: out = fp16(LN(x) / ||LN(x)||_2) with the
 * LayerNorm evaluated in fp32 like autocast does (tokenflow_utils.py:323 norm1 + util.py:66-67).  Only
 * the source stream's rows are needed there (reference :335), so callers pass that third only.
 *   x_f16        device [rows, dim] fp16, row pitch `x_row_stride` elements (multiple of 8)
 *   gamma, beta  device [dim] fp32 (norm1.weight / norm1.bias), eps = norm1.eps
 *   out_f16      device [rows, dim] fp16 contiguous;  dim <= 1280 */
int tf_layernorm_unit_rows(const void* x_f16, int64_t rows, int dim, int64_t x_row_stride, const float* gamma,
                           const float* beta, float eps, void* out_f16, tf_stream_t stream);

/* norm1 of the PIVOTAL pass fused with both of its consumers (tokenflow_utils.py:323 -> :120-122 and
 * :326-327 + util.py:66-67): one read of hidden_states produces
 *   y_out     fp16(LN(x)) for every row — the operand autocast would cast for the to_q/to_k/to_v GEMMs
 *   unit_out  fp16(LN(x) / ||LN(x)||_2) for the first `unit_rows` rows (the source-stream samples: the
 *             pivot features the NN field correlates), from the unrounded fp32 LN like the reference
 * Either output may be NULL.  Outputs may be strided (packed all-gather buffers): row pitch in elements,
 * multiple of 8.  dim <= 1280. */
int tf_layernorm_rows(const void* x_f16, int64_t rows, int dim, int64_t x_row_stride, const float* gamma,
                      const float* beta, float eps, void* y_out_f16, int64_t y_row_stride, void* unit_out_f16,
                      int64_t unit_row_stride, int64_t unit_rows, tf_stream_t stream);

/* Token nearest-neighbour field.  Replaces tokenflow_utils.py:329-348 (+ util.py:68 `x @ y.T`,
 * fp16 output under autocast, and the two argmax reductions :340-343):
 *   idx_a[f,p] = argmax_c fp16( x_unit[f,p,:] . piv_unit[kf_a[f],c,:] )     first index on ties
 *   idx_b[f,p] = same against keyframe kf_b[f]            (skipped for frames with kf_b[f] < 0)
 *   x_unit    device [F, S, dim] fp16 unit rows (tf_unit_rows of the source-stream norm1 output)
 *   piv_unit  device [K, S, dim] fp16 unit rows of the cached source-stream pivot features
 *   kf_a,kf_b host   [F] keyframe ids in [0,K); reference batch i: kf_a = i, kf_b = i-1 (or -1)
 *   idx_a,idx_b device [F, S] int32 (idx_b may be NULL when no frame has a second keyframe) */
int tf_nn_field(const void* x_unit, const void* piv_unit, const int32_t* kf_a, const int32_t* kf_b, int F, int S,
                int dim, int K, int32_t* idx_a, int32_t* idx_b, tf_stream_t stream);

/* NN-indexed feature propagation.  Replaces tokenflow_utils.py:361-397 (keyframe slice, two
 * gathers, blend weights, blend, residual add):
 *   out[s,f,p,:] = w[f]*A[s,kf_a[f],idx_a[f,p],:] + (1-w[f])*A[s,kf_b[f],idx_b[f,p],:] (+ residual[s,f,p,:])
 *   (kf_b[f] < 0:  out = A[s,kf_a[f],idx_a[f,p],:] (+ residual))
 *   A         device [3, K, S, dim] fp16   (cached attn1 output of the pivotal pass, `kf_attn_output`)
 *   w         host   [F] fp32 blend weights (reference: sigmoid(d2/(d1+d2)), :375-383)
 *   residual  device [3, F, S, dim] fp16 or NULL  (the block's `hidden_states`, :396-397)
 *   out       device [3, F, S, dim] fp16 (out_is_f32 == 0) or fp32 (the reference's promoted dtype) */
int tf_propagate(const void* A, const int32_t* idx_a, const int32_t* idx_b, const int32_t* kf_a,
                 const int32_t* kf_b, const float* w, int F, int S, int dim, int K, const void* residual,
                 void* out, int out_is_f32, tf_stream_t stream);

/* Extended attention of one pivotal pass.  Replaces the body of the attn1 closure between the
 * q/k/v projections and `to_out` (tokenflow_utils.py:124-197 PnP flavour, :234-279 SDEdit flavour).
 *   q,k,v   device [3n, S, heads, d] fp16; consecutive tokens `tok_stride` elements apart (heads*d
 *           when contiguous, 3*heads*d for a fused qkv buffer); samples S*tok_stride apart
 *   out     device [3n, S, heads*d] fp16 contiguous
 *   inject  != 0: PnP q/k injection — uncond and cond samples read the source stream's q and k
 *           (reference :124-130) by aliasing, nothing is copied
 * Batch axis is [source | uncond | cond] thirds like the reference (:117).  Source samples attend
 * to their own frame, uncond/cond samples to all n frames of their stream. */
int tf_ext_attn_fwd(const void* q, const void* k, const void* v, int64_t tok_stride, int n_frames, int S,
                    int heads, int d, float scale, int inject, void* out, tf_stream_t stream);

/* General form used when the pivotal pass is sharded across GPUs: `n_out` output samples, each
 * described by host arrays (all length n_out): which slab of `out` it writes, which slab of q it
 * reads, the first k / v slab it attends to and how many consecutive slabs (1 = own frame, n = all
 * keyframes).  q has q_slabs slabs of [S, heads, d]; k and v have kv_slabs. */
int tf_ext_attn_fwd_table(const void* q, int q_slabs, int64_t q_tok_stride, const void* k, const void* v,
                          int kv_slabs, int64_t kv_tok_stride, int n_out, const int32_t* out_slab,
                          const int32_t* q_slab, const int32_t* k_slab0, const int32_t* v_slab0,
                          const int32_t* n_kv, int S, int heads, int d, float scale, void* out,
                          tf_stream_t stream);

/* The same for a RANGE of query tokens: only queries [q_row0, q_row0 + q_nrows) of every sample are computed
 * (against all keys), and `out` is [slabs, q_nrows, heads*d] with row = token - q_row0.  q_row0 must be a
 * multiple of 128.  The multi-GPU pivotal pass splits the query rows of ALL samples evenly over the ranks this
 * way (every rank holds all K/V after the all-gather), which balances the attention work exactly and keeps
 * paired (q/k-injected) samples together. */
int tf_ext_attn_fwd_rows(const void* q, int q_slabs, int64_t q_tok_stride, const void* k, const void* v,
                         int kv_slabs, int64_t kv_tok_stride, int n_out, const int32_t* out_slab,
                         const int32_t* q_slab, const int32_t* k_slab0, const int32_t* v_slab0,
                         const int32_t* n_kv, int S, int heads, int d, float scale, int q_row0, int q_nrows,
                         void* out, tf_stream_t stream);

/* Classifier-free guidance + DDIM update (eta = 0) of one denoising step, one pass over the latents.
 * Replaces run_tokenflow_pnp.py:213-217 (`u + g*(c - u)`, `scheduler.step(...)['prev_sample']`), with the
 * fp16 rounding sequence of the eager expression (bit-identical results).
 *   eps_uncond, eps_cond, x, out   device [n] fp16 contiguous (x = the latents being denoised)
 *   coef   device [4] fp32: sqrt(1-alpha_t), 1/sqrt(alpha_t), sqrt(alpha_prev), sqrt(1-alpha_prev) — in
 *          device memory so a captured CUDA graph of the step replays for every timestep */
int tf_cfg_ddim(const void* eps_uncond, const void* eps_cond, const void* x, const float* coef, float guidance,
                int64_t n, void* out, tf_stream_t stream);

/* ---- multi-GPU: all-gather of keyframe tensors along the pivotal-sample axis ----
 * NCCL (all-gather over NVLink 5 / NVSwitch) bound at run time; one communicator per process/GPU.
 * Rendezvous: rank 0 calls tf_comm_unique_id and ships the TF_COMM_ID_BYTES to the other ranks by any
 * channel (the Python host uses a torch.distributed broadcast); every rank then calls tf_comm_init.
 * tf_allgather enqueues on `stream` (CUDA-graph capturable): recv = [nranks][bytes_per_rank]. */
#define TF_COMM_ID_BYTES 128
typedef void* tf_comm_t;
int tf_comm_nccl_version(void);                 /* NCCL version code, 0 if NCCL cannot be loaded */
int tf_comm_unique_id(void* id_out);
int tf_comm_init(const void* id, int nranks, int rank, tf_comm_t* comm_out);
int tf_allgather(tf_comm_t comm, const void* send, void* recv, int64_t bytes_per_rank, tf_stream_t stream);
int tf_comm_destroy(tf_comm_t comm);

#ifdef __cplusplus
}
#endif
#endif /* TOKENFLOW_B200_H_ */
