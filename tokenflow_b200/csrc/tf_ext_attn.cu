// tf_ext_attn — extended cross-frame self-attention over the keyframes
// (reference tokenflow_utils.py:114-199 PnP flavour, :224-281 SDEdit flavour).
//
// For every output sample (stream s, keyframe f), head j and query token p:
//     O[p,:] = softmax_c( Q[p,:] . K[c,:] * scale ) V[c,:]
// where c runs over the S tokens of the sample's own frame (source stream) or over the n*S tokens
// of ALL keyframes of its stream, frame-major (uncond / cond streams).  The reference materialises
// per head an [n, S, n*S] fp16 score tensor and an fp32 probability tensor (0.84 + 1.68 GB at the
// 40-frame SD1.5 top level), replicates K and V n times and shuffles heads through ~10 copies
// (SURVEY.md §2.1 k1-k6).  Here none of that exists: one CTA owns a 128-query tile of one
// (sample, head), streams the key/value tiles of every attended keyframe through shared memory
// with TMA, keeps scores, probabilities and the output accumulator in tensor memory, and addresses
// heads by stride inside the [sample, S, heads, d] tensors.  PnP q/k injection (:124-130) is pure
// aliasing: the per-sample table names which q / k slab to read.
//
// CTA = 6 warps:   warp 0  TMA producer (Q tile once, then a ring of {K tile, V tile} stages)
//                  warp 1  tcgen05.mma issuer:  S[b] = Q K_t^T (SS),  O += P_t V_t (TS, P read from TMEM)
//                  warps 2-5  softmax, one thread per query row: tcgen05.ld S -> running max / exp2 /
//                             row sum -> fp16 P written back over S with tcgen05.st; lazy O rescale
//                             (only when the running max grows by > 2^8, FA4-style); final O / l.
// TMEM (512 cols): S/P buffers 2 x kBlockN fp32 columns, O accumulator d_pad columns.
// QK(t+1) is issued before PV(t) waits for P(t), so the tensor pipe computes the next score tile
// while the softmax warps work on the current one.
//
// Head dims that are not a multiple of 64 (SD1.5: 40, 80, 160) are zero-padded by TMA out-of-bound
// fill: the tensor maps describe [d, heads, S, samples] with the true inner extent d and a 64-wide
// box, so shared-memory rows are always one full 128-byte swizzle row.
//
// Roofline: tensor-bound, 4*S_q*S_kv*d flops per (sample, head); HBM traffic is q,k,v,out once
// (K/V tiles re-read by the other query tiles hit L2).
#include <cstdlib>

#include "tf_common.cuh"
#include "tf_kernels.h"

namespace tf {
namespace {

constexpr int kBlockM = 128;
constexpr float kRescaleThreshold = 8.0f;      // log2 units: P stays <= 2^8 without touching O

struct AttnCtl {
  uint64_t q_full;
  uint64_t kv_full[8];
  uint64_t kv_empty[8];
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t pv_done;
  uint32_t tmem_base;
};

struct AttnParams {
  int S, heads, d, n_out;
  int tiles_m;            // query tiles per (sample, head)
  int stages;
  float scale_log2;       // scale * log2(e)
  long long out_tok_stride;   // elements between consecutive tokens of `out` (= heads*d)
};

template <int kDChunks, int kBlockN>
__global__ void __launch_bounds__(192, 1)
ext_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const AttnTable tab, const AttnParams prm,
                __half* __restrict__ out) {
  constexpr int kQChunkBytes = kBlockM * 128;
  constexpr int kKVChunkBytes = kBlockN * 128;
  constexpr int kQBytes = kDChunks * kQChunkBytes;
  constexpr int kTileBytes = kDChunks * kKVChunkBytes;         // one K tile or one V tile
  constexpr int kStageBytes = 2 * kTileBytes;
  constexpr int kSCols = kBlockN;                               // fp32 score columns per buffer
  constexpr int kOCol = 2 * kSCols;                             // O accumulator starts after the S buffers
  constexpr int kDPad = 64 * kDChunks;
  static_assert(kOCol + kDPad <= 512, "tensor memory overflow");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + kQBytes;
  AttnCtl* ctl = reinterpret_cast<AttnCtl*>(ring + prm.stages * kStageBytes);

  // ---- work item: heavy (extended) samples first so the tail of the grid is made of light items ----
  const int S = prm.S, d = prm.d, stages = prm.stages;
  const int per_sample = prm.heads * prm.tiles_m;
  const int sample_slot = blockIdx.x / per_sample;
  const int rem = blockIdx.x - sample_slot * per_sample;
  const int head = rem / prm.tiles_m;
  const int m0 = (rem - head * prm.tiles_m) * kBlockM;
  const AttnSample smp = tab.s[sample_slot];
  const int out_sample = smp.out_sample;
  const int q_slab = smp.q_sample;
  const int tiles_per_slab = (S + kBlockN - 1) / kBlockN;
  const int T = smp.n_kv * tiles_per_slab;
  const int ksteps = (d + 15) / 16;                             // QK^T k-steps (zero padded to 16)
  const int n_pv = ((d + 15) / 16) * 16;                        // PV MMA N (multiple of 16)

  const int warp = threadIdx.x >> 5;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(&ctl->q_full, 1);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&ctl->kv_full[i], 1);
      mbar_init(&ctl->kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->s_full[i], 1);
      mbar_init(&ctl->p_full[i], 4);
    }
    mbar_init(&ctl->pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&ctl->q_full, (uint32_t)kQBytes);
#pragma unroll
      for (int c = 0; c < kDChunks; ++c)
        tma_load_4d(q_smem + c * kQChunkBytes, &map_q, &ctl->q_full, c * 64, head, m0, q_slab);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < T; ++t) {
        const int slab = t / tiles_per_slab;
        const int n0 = (t - slab * tiles_per_slab) * kBlockN;
        mbar_wait(&ctl->kv_empty[stage], phase ^ 1);
        uint8_t* st = ring + stage * kStageBytes;
        mbar_arrive_expect_tx(&ctl->kv_full[stage], (uint32_t)kStageBytes);
#pragma unroll
        for (int c = 0; c < kDChunks; ++c) {
          tma_load_4d(st + c * kKVChunkBytes, &map_k, &ctl->kv_full[stage], c * 64, head, n0, smp.k_sample0 + slab);
          tma_load_4d(st + kTileBytes + c * kKVChunkBytes, &map_v, &ctl->kv_full[stage], c * 64, head, n0,
                      smp.v_sample0 + slab);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_qk = umma_idesc_f16(128, kBlockN, 0);
    const uint32_t idesc_pv = umma_idesc_f16(128, (uint32_t)n_pv, 1);     // B = V tile, MN-major
    const uint32_t q_addr = smem_u32(q_smem);
    auto issue_qk = [&](int t, int stage) {        // S[t&1] = Q K_t^T
      const uint32_t k_addr = smem_u32(ring + stage * kStageBytes);
      const uint32_t s_tmem = tmem_base + (uint32_t)((t & 1) * kSCols);
      for (int ks = 0; ks < ksteps; ++ks) {
        const int c = ks >> 2, k4 = ks & 3;
        const uint64_t da = umma_smem_desc(q_addr + c * kQChunkBytes + k4 * 32, 16, 1024);
        const uint64_t db = umma_smem_desc(k_addr + c * kKVChunkBytes + k4 * 32, 16, 1024);
        tc_mma_ss(s_tmem, da, db, idesc_qk, ks > 0 ? 1u : 0u);
      }
      tc_commit(&ctl->s_full[t & 1]);
    };
    mbar_wait(&ctl->q_full, 0);
    mbar_wait(&ctl->kv_full[0], 0);
    tc_fence_after_sync();
    if (elect_one()) issue_qk(0, 0);
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < T; ++t) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == stages) { nstage = 0; nphase ^= 1; }
      if (t + 1 < T) {                              // next score tile first: overlaps softmax(t)
        mbar_wait(&ctl->kv_full[nstage], nphase);
        tc_fence_after_sync();
        if (elect_one()) issue_qk(t + 1, nstage);
        __syncwarp();
      }
      mbar_wait(&ctl->p_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after_sync();
      if (elect_one()) {                            // O (+)= P_t V_t,  P_t = fp16 [128 x kBlockN] in TMEM
        const uint32_t v_addr = smem_u32(ring + stage * kStageBytes + kTileBytes);
        const uint32_t p_tmem = tmem_base + (uint32_t)((t & 1) * kSCols);
#pragma unroll
        for (int k = 0; k < kBlockN / 16; ++k) {
          const uint64_t db = umma_smem_desc(v_addr + k * (16 * 128), (uint32_t)kKVChunkBytes, 1024);
          tc_mma_ts(tmem_base + kOCol, p_tmem + k * 8, db, idesc_pv, (t > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&ctl->kv_empty[stage]);
        tc_commit(&ctl->pv_done);
      }
      __syncwarp();
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ===================== softmax / correction / epilogue: one thread per query row ==============
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane_id();
    const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
    const float sl2 = prm.scale_log2;
    float m_run = 0.f;       // running max in scaled log2 units
    float l_run = 0.f;
    for (int t = 0; t < T; ++t) {
      const int slab_tile = t % tiles_per_slab;
      const int valid = min(kBlockN, S - slab_tile * kBlockN);   // key columns of this tile inside the slab
      const uint32_t s_addr = tmem_base + t_lane + (uint32_t)((t & 1) * kSCols);
      mbar_wait(&ctl->s_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after_sync();
      // ---- pass 1: tile max ----
      float mt = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < kBlockN; c0 += 32) {
        if (c0 >= valid) break;
        uint32_t v[32];
        tmem_ld32(s_addr + c0, v);
        tmem_wait_ld();
        if (valid - c0 >= 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mt = fmaxf(mt, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (c0 + i < valid) mt = fmaxf(mt, __uint_as_float(v[i]));
        }
      }
      const float mt_s = mt * sl2;
      if (t == 0) {
        m_run = mt_s;
      } else {
        const bool need = mt_s > m_run + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {       // warp-uniform: tcgen05.ld/st are warp-collective
          mbar_wait(&ctl->pv_done, (uint32_t)((t - 1) & 1));     // PV(t-1) retired: O is quiescent
          tc_fence_after_sync();
          const float m_new = fmaxf(m_run, mt_s);
          const float alpha = fast_exp2(m_run - m_new);
          l_run *= alpha;
          m_run = m_new;
          for (int c0 = 0; c0 < n_pv; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(tmem_base + t_lane + kOCol + c0, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tmem_base + t_lane + kOCol + c0, o);
          }
          tmem_wait_st();
        }
      }
      // ---- pass 2: p = 2^(s*scale*log2e - m), row sum, fp16 P written over the score columns ----
      float lsum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < kBlockN; c0 += 32) {
        uint32_t pk[16];
        if (c0 < valid) {
          uint32_t v[32];
          tmem_ld32(s_addr + c0, v);
          tmem_wait_ld();
          float p[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = fast_exp2(fmaf(__uint_as_float(v[i]), sl2, -m_run));
            if (valid - c0 < 32 && c0 + i >= valid) p[i] = 0.f;
            lsum += p[i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_f16x2_rn(p[2 * i], p[2 * i + 1]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = 0u;
        }
        tmem_st16(s_addr + (c0 >> 1), pk);
      }
      l_run += lsum;
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&ctl->p_full[t & 1]);
    }
    // ---- final: O / l -> fp16 -> out[sample, p, head, :] ----
    mbar_wait(&ctl->pv_done, (uint32_t)((T - 1) & 1));
    tc_fence_after_sync();
    const float inv_l = 1.0f / l_run;
    const int p_tok = m0 + row;
    __half* orow = out + ((long long)out_sample * S + p_tok) * prm.out_tok_stride + (long long)head * d;
    for (int c0 = 0; c0 < n_pv; c0 += 16) {
      uint32_t o[16];
      tmem_ld16(tmem_base + t_lane + kOCol + c0, o);
      tmem_wait_ld();
      if (p_tok < S) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c0 + g * 8 < d) {                     // d is a multiple of 8: whole 16-byte groups
            uint4 w;
            w.x = pack_f16x2_rn(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
            w.y = pack_f16x2_rn(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
            w.z = pack_f16x2_rn(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
            w.w = pack_f16x2_rn(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c0 + g * 8) = w;
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}


// ================================================================================================
// v2: two 128-query tiles per CTA ("ping-pong"), head dim <= 64.
//
// The tensor pipe and the softmax warps alternate between the two query tiles, so S_B = Q_B K^T and
// O_A += P_A V run while the other tile's warps are in their exp2 loop; K/V tiles are fetched once
// for 256 queries.  Each softmax thread keeps its whole 128-column score row in registers (one TMEM
// read per tile), and the row sum is not computed by the softmax warps at all: an extra N=16
// tcgen05.mma against a constant tile of ones accumulates L = sum_k P[.,k] in TMEM from exactly the
// fp16 probabilities that feed P V, so numerator and denominator stay consistent.
//   warp 0: TMA   warp 1: MMA issue   warps 2-5: softmax tile A   warps 6-9: softmax tile B
// TMEM: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) L_A [384,400) L_B [400,416)
// ================================================================================================
struct AttnCtl2 {
  uint64_t q_full;
  uint64_t kv_full[8];
  uint64_t kv_empty[8];
  uint64_t s_full[2];
  uint64_t p_full[2];
  uint64_t pv_done[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(320, 1)
ext_attn_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const AttnTable tab, const AttnParams prm,
                   __half* __restrict__ out) {
  constexpr int kBlockN = 128;
  constexpr int kQTileBytes = kBlockM * 128;          // one 128-query tile, 64-wide d chunk
  constexpr int kQBytes = 2 * kQTileBytes;
  constexpr int kOnesBytes = 16 * 128;                // 16 key rows of ones (B operand of the row-sum MMA)
  constexpr int kTileBytes = kBlockN * 128;
  constexpr int kStageBytes = 2 * kTileBytes;
  constexpr int kOCol = 256, kLCol = 384;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* ones_smem = smem + kQBytes;
  uint8_t* ring = ones_smem + kOnesBytes;
  AttnCtl2* ctl = reinterpret_cast<AttnCtl2*>(ring + prm.stages * kStageBytes);

  const int S = prm.S, d = prm.d, stages = prm.stages;
  const int per_sample = prm.heads * prm.tiles_m;              // tiles_m = 256-query tile pairs here
  const int sample_slot = blockIdx.x / per_sample;
  const int rem = blockIdx.x - sample_slot * per_sample;
  const int head = rem / prm.tiles_m;
  const int m0 = (rem - head * prm.tiles_m) * (2 * kBlockM);
  const AttnSample smp = tab.s[sample_slot];
  const int tiles_per_slab = (S + kBlockN - 1) / kBlockN;
  const int T = smp.n_kv * tiles_per_slab;
  const int ksteps = (d + 15) / 16;
  const int n_pv = ((d + 15) / 16) * 16;

  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ones_smem)[i] = 0x3C003C00u;   // fp16 1.0 pairs (layout/swizzle agnostic)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(&ctl->q_full, 1);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&ctl->kv_full[i], 1);
      mbar_init(&ctl->kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->s_full[i], 1);
      mbar_init(&ctl->p_full[i], 4);
      mbar_init(&ctl->pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(&ctl->q_full, (uint32_t)kQBytes);
      tma_load_4d(q_smem, &map_q, &ctl->q_full, 0, head, m0, smp.q_sample);
      tma_load_4d(q_smem + kQTileBytes, &map_q, &ctl->q_full, 0, head, m0 + kBlockM, smp.q_sample);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < T; ++t) {
        const int slab = t / tiles_per_slab;
        const int n0 = (t - slab * tiles_per_slab) * kBlockN;
        mbar_wait(&ctl->kv_empty[stage], phase ^ 1);
        uint8_t* st = ring + stage * kStageBytes;
        mbar_arrive_expect_tx(&ctl->kv_full[stage], (uint32_t)kStageBytes);
        tma_load_4d(st, &map_k, &ctl->kv_full[stage], 0, head, n0, smp.k_sample0 + slab);
        tma_load_4d(st + kTileBytes, &map_v, &ctl->kv_full[stage], 0, head, n0, smp.v_sample0 + slab);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_qk = umma_idesc_f16(128, kBlockN, 0);
    const uint32_t idesc_pv = umma_idesc_f16(128, (uint32_t)n_pv, 1);
    const uint32_t idesc_l = umma_idesc_f16(128, 16, 1);
    const uint32_t q_addr = smem_u32(q_smem);
    const uint64_t ones_desc = umma_smem_desc(smem_u32(ones_smem), 16, 1024);
    auto issue_qk = [&](int X, int stage) {       // S_X = Q_X K^T
      const uint32_t k_addr = smem_u32(ring + stage * kStageBytes);
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = umma_smem_desc(q_addr + X * kQTileBytes + ks * 32, 16, 1024);
        const uint64_t db = umma_smem_desc(k_addr + ks * 32, 16, 1024);
        tc_mma_ss(tmem_base + (uint32_t)(X * 128), da, db, idesc_qk, ks > 0 ? 1u : 0u);
      }
      tc_commit(&ctl->s_full[X]);
    };
    mbar_wait(&ctl->q_full, 0);
    mbar_wait(&ctl->kv_full[0], 0);
    tc_fence_after_sync();
    if (elect_one()) { issue_qk(0, 0); issue_qk(1, 0); }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < T; ++t) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == stages) { nstage = 0; nphase ^= 1; }
#pragma unroll
      for (int X = 0; X < 2; ++X) {
        mbar_wait(&ctl->p_full[X], (uint32_t)(t & 1));
        tc_fence_after_sync();
        if (elect_one()) {
          const uint32_t v_addr = smem_u32(ring + stage * kStageBytes + kTileBytes);
          const uint32_t p_tmem = tmem_base + (uint32_t)(X * 128);
#pragma unroll
          for (int k = 0; k < kBlockN / 16; ++k) {
            const uint64_t db = umma_smem_desc(v_addr + k * (16 * 128), (uint32_t)kTileBytes, 1024);
            const uint32_t acc = (t > 0 || k > 0) ? 1u : 0u;
            tc_mma_ts(tmem_base + kOCol + X * 64, p_tmem + k * 8, db, idesc_pv, acc);
            tc_mma_ts(tmem_base + kLCol + X * 16, p_tmem + k * 8, ones_desc, idesc_l, acc);
          }
          tc_commit(&ctl->pv_done[X]);
          if (X == 1) tc_commit(&ctl->kv_empty[stage]);
        }
        __syncwarp();
        if (t + 1 < T) {
          if (X == 0) {
            mbar_wait(&ctl->kv_full[nstage], nphase);
            tc_fence_after_sync();
          }
          if (elect_one()) issue_qk(X, nstage);
          __syncwarp();
        }
      }
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ===================== softmax warps: tile X = (warp - 2) / 4 =====================
    const int X = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane_id();
    const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + t_lane + (uint32_t)(X * 128);
    const uint32_t o_addr = tmem_base + t_lane + kOCol + X * 64;
    const uint32_t l_addr = tmem_base + t_lane + kLCol + X * 16;
    const float sl2 = prm.scale_log2;
    float m_run = 0.f;
    for (int t = 0; t < T; ++t) {
      const int slab_tile = t % tiles_per_slab;
      const int valid = min(kBlockN, S - slab_tile * kBlockN);
      mbar_wait(&ctl->s_full[X], (uint32_t)(t & 1));
      tc_fence_after_sync();
      uint32_t v0[32], v1[32], v2[32], v3[32];
      tmem_ld32(s_addr, v0);
      tmem_ld32(s_addr + 32, v1);
      tmem_ld32(s_addr + 64, v2);
      tmem_ld32(s_addr + 96, v3);
      tmem_wait_ld();
      if (valid < kBlockN) {                       // ragged last tile of a keyframe: mask the padding keys
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) v0[i] = 0xFF800000u;
          if (32 + i >= valid) v1[i] = 0xFF800000u;
          if (64 + i >= valid) v2[i] = 0xFF800000u;
          if (96 + i >= valid) v3[i] = 0xFF800000u;
        }
      }
      float mt = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mt = fmaxf(mt, fmaxf(fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])),
                             fmaxf(__uint_as_float(v2[i]), __uint_as_float(v3[i]))));
      }
      const float mt_s = mt * sl2;
      if (t == 0) {
        m_run = mt_s;
      } else {
        const bool need = mt_s > m_run + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&ctl->pv_done[X], (uint32_t)((t - 1) & 1));
          tc_fence_after_sync();
          const float m_new = fmaxf(m_run, mt_s);
          const float alpha = fast_exp2(m_run - m_new);
          m_run = m_new;
          for (int c0 = 0; c0 < n_pv; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(o_addr + c0, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(o_addr + c0, o);
          }
          {
            uint32_t o[16];
            tmem_ld16(l_addr, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(l_addr, o);
          }
          tmem_wait_st();
        }
      }
      const float neg_m = -m_run;
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_f16x2_rn(fast_exp2(fmaf(__uint_as_float(v0[2 * i]), sl2, neg_m)),
                              fast_exp2(fmaf(__uint_as_float(v0[2 * i + 1]), sl2, neg_m)));
      tmem_st16(s_addr, pk);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_f16x2_rn(fast_exp2(fmaf(__uint_as_float(v1[2 * i]), sl2, neg_m)),
                              fast_exp2(fmaf(__uint_as_float(v1[2 * i + 1]), sl2, neg_m)));
      tmem_st16(s_addr + 16, pk);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_f16x2_rn(fast_exp2(fmaf(__uint_as_float(v2[2 * i]), sl2, neg_m)),
                              fast_exp2(fmaf(__uint_as_float(v2[2 * i + 1]), sl2, neg_m)));
      tmem_st16(s_addr + 32, pk);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_f16x2_rn(fast_exp2(fmaf(__uint_as_float(v3[2 * i]), sl2, neg_m)),
                              fast_exp2(fmaf(__uint_as_float(v3[2 * i + 1]), sl2, neg_m)));
      tmem_st16(s_addr + 48, pk);
      tmem_wait_st();
      tc_fence_before_sync();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&ctl->p_full[X]);
    }
    // ---- final: O / L -> fp16 ----
    mbar_wait(&ctl->pv_done[X], (uint32_t)((T - 1) & 1));
    tc_fence_after_sync();
    uint32_t lreg[16];
    tmem_ld16(l_addr, lreg);
    tmem_wait_ld();
    const float inv_l = 1.0f / __uint_as_float(lreg[0]);
    const int p_tok = m0 + X * kBlockM + row;
    __half* orow = out + ((long long)smp.out_sample * S + p_tok) * prm.out_tok_stride + (long long)head * d;
    for (int c0 = 0; c0 < n_pv; c0 += 16) {
      uint32_t o[16];
      tmem_ld16(o_addr + c0, o);
      tmem_wait_ld();
      if (p_tok < S) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (c0 + g * 8 < d) {
            uint4 w;
            w.x = pack_f16x2_rn(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
            w.y = pack_f16x2_rn(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
            w.z = pack_f16x2_rn(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
            w.w = pack_f16x2_rn(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c0 + g * 8) = w;
          }
        }
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_pp(const void* q, const void* k, const void* v, long long q_tok_stride, long long kv_tok_stride,
              int q_samples_total, int kv_samples_total, const AttnTable& tab, int n_out, int S, int heads, int d,
              float scale, void* out, cudaStream_t stream) {
  constexpr int kQBytes = 2 * kBlockM * 128, kOnesBytes = 16 * 128, kStageBytes = 2 * 128 * 128;
  int stages = (227 * 1024 - 2048 - kQBytes - kOnesBytes) / kStageBytes;
  if (stages > 8) stages = 8;
  const size_t smem_bytes = 1024 + kQBytes + kOnesBytes + (size_t)stages * kStageBytes + sizeof(AttnCtl2);
  CUtensorMap map_q, map_k, map_v;
  auto make = [&](CUtensorMap* m, const void* base, long long tok_stride, int samples, int box_rows) -> int {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)S, (uint64_t)samples};
    const uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)tok_stride * 2, (uint64_t)S * tok_stride * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    CUresult r = encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (r != CUDA_SUCCESS) { set_last_error("tf_ext_attn: cuTensorMapEncodeTiled failed: %d", (int)r); return TF_ERR_DRIVER; }
    return TF_OK;
  };
  if (int e = make(&map_q, q, q_tok_stride, q_samples_total, kBlockM)) return e;
  if (int e = make(&map_k, k, kv_tok_stride, kv_samples_total, 128)) return e;
  if (int e = make(&map_v, v, kv_tok_stride, kv_samples_total, 128)) return e;
  AttnParams prm;
  prm.S = S; prm.heads = heads; prm.d = d; prm.n_out = n_out;
  prm.tiles_m = (S + 2 * kBlockM - 1) / (2 * kBlockM);
  prm.stages = stages;
  prm.scale_log2 = scale * 1.4426950408889634f;
  prm.out_tok_stride = (long long)heads * d;
  if (check_cuda(cudaFuncSetAttribute(ext_attn_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes),
                 "tf_ext_attn smem attribute"))
    return TF_ERR_CUDA;
  const long long grid = (long long)n_out * heads * prm.tiles_m;
  ext_attn_pp_kernel<<<(unsigned)grid, 320, smem_bytes, stream>>>(map_q, map_k, map_v, tab, prm,
                                                                 static_cast<__half*>(out));
  return check_cuda(cudaGetLastError(), "tf_ext_attn launch");
}

template <int kDChunks, int kBlockN>
int launch_cfg(const void* q, const void* k, const void* v, long long q_tok_stride, long long kv_tok_stride,
               int q_samples_total, int kv_samples_total, const AttnTable& tab, int n_out, int S, int heads, int d,
               float scale, void* out, cudaStream_t stream) {
  constexpr int kQBytes = kDChunks * kBlockM * 128;
  constexpr int kStageBytes = 2 * kDChunks * kBlockN * 128;
  int stages = (227 * 1024 - 2048 - kQBytes) / kStageBytes;
  if (stages > 8) stages = 8;
  if (stages < 2) { set_last_error("tf_ext_attn: configuration does not fit shared memory"); return TF_ERR_UNSUPPORTED; }
  const size_t smem_bytes = 1024 + kQBytes + (size_t)stages * kStageBytes + sizeof(AttnCtl);

  CUtensorMap map_q, map_k, map_v;
  auto make = [&](CUtensorMap* m, const void* base, long long tok_stride, int samples, int box_rows) -> int {
    const uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)S, (uint64_t)samples};
    const uint64_t strides[3] = {(uint64_t)d * 2, (uint64_t)tok_stride * 2, (uint64_t)S * tok_stride * 2};
    const uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    CUresult r = encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box,
                              CU_TENSOR_MAP_SWIZZLE_128B);
    if (r != CUDA_SUCCESS) { set_last_error("tf_ext_attn: cuTensorMapEncodeTiled failed: %d", (int)r); return TF_ERR_DRIVER; }
    return TF_OK;
  };
  if (int e = make(&map_q, q, q_tok_stride, q_samples_total, kBlockM)) return e;
  if (int e = make(&map_k, k, kv_tok_stride, kv_samples_total, kBlockN)) return e;
  if (int e = make(&map_v, v, kv_tok_stride, kv_samples_total, kBlockN)) return e;

  AttnParams prm;
  prm.S = S; prm.heads = heads; prm.d = d; prm.n_out = n_out;
  prm.tiles_m = (S + kBlockM - 1) / kBlockM;
  prm.stages = stages;
  prm.scale_log2 = scale * 1.4426950408889634f;
  prm.out_tok_stride = (long long)heads * d;

  auto kern = ext_attn_kernel<kDChunks, kBlockN>;
  if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes),
                 "tf_ext_attn smem attribute"))
    return TF_ERR_CUDA;
  const long long grid = (long long)n_out * heads * prm.tiles_m;
  kern<<<(unsigned)grid, 192, smem_bytes, stream>>>(map_q, map_k, map_v, tab, prm, static_cast<__half*>(out));
  return check_cuda(cudaGetLastError(), "tf_ext_attn launch");
}

}  // namespace

int launch_ext_attn(const void* q, const void* k, const void* v, long long q_tok_stride, long long kv_tok_stride,
                    int q_samples_total, int kv_samples_total, const AttnTable& tab, int n_out, int S, int heads,
                    int d, float scale, void* out, cudaStream_t stream) {
  if (n_out == 0 || S == 0) return TF_OK;
  static const bool force_v1 = getenv("TF_EXT_ATTN_V1") != nullptr;     // A/B switch for profiling
  if (d <= 64 && S > 128 && !force_v1)
    return launch_pp(q, k, v, q_tok_stride, kv_tok_stride, q_samples_total, kv_samples_total, tab, n_out, S, heads,
                     d, scale, out, stream);
  if (d <= 64)
    return launch_cfg<1, 128>(q, k, v, q_tok_stride, kv_tok_stride, q_samples_total, kv_samples_total, tab, n_out,
                              S, heads, d, scale, out, stream);
  if (d <= 128)
    return launch_cfg<2, 128>(q, k, v, q_tok_stride, kv_tok_stride, q_samples_total, kv_samples_total, tab, n_out,
                              S, heads, d, scale, out, stream);
  if (d <= 192)
    return launch_cfg<3, 64>(q, k, v, q_tok_stride, kv_tok_stride, q_samples_total, kv_samples_total, tab, n_out,
                             S, heads, d, scale, out, stream);
  set_last_error("tf_ext_attn: head dim %d > 192 is not supported", d);
  return TF_ERR_UNSUPPORTED;
}

}  // namespace tf
