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
constexpr float kRescaleThreshold = 8.0f;

// Optional event trace (build with -DTF_TRACE): clock64 stamps of CTA 0's roles for the first tiles.
// layout: g_attn_trace[role][tile][event], role 0/1 = softmax tile A/B (warp quadrant 0, lane 0), 2 = MMA issuer
#ifdef TF_TRACE
constexpr int kTraceTiles = 40, kTraceEvents = 8;
__device__ long long g_attn_trace[3 * kTraceTiles * kTraceEvents];
#define TF_TRACE_EV(role, tile, ev)                                                              \
  do {                                                                                           \
    if (blockIdx.x == 0 && (tile) < kTraceTiles)                                                 \
      g_attn_trace[((role) * kTraceTiles + (tile)) * kTraceEvents + (ev)] = clock64();           \
  } while (0)
#else
#define TF_TRACE_EV(role, tile, ev) do {} while (0)
#endif      // log2 units: P stays <= 2^8 without touching O

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
  int handoff;            // ping-pong kernel: chunk index after which the MUFU token is handed over
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
// Ping-pong kernel (head dim <= 64): two 128-query tiles per CTA.
//
// The softmax inner loop is MUFU-bound on this chip (ex2: 8.1 cycles per warp instruction per SM
// sub-partition, profiles/r01_pipe_throughput_ubench.txt; at d = 40 the two MMAs of a 128x128 tile
// need only ~450 tensor cycles against ~1040 MUFU cycles), and a tcgen05.mma costs its issuing thread
// 45-100 cycles whatever its shape (profiles/r01_ext_attn_trace.md).  The structure follows from that:
//   * two query tiles A and B share every K/V tile (fetched once per 256 queries);
//   * one MMA-issuer warp per query tile, so the two issue streams run in parallel; operands stay
//     warp-uniform so ptxas emits back-to-back UTCHMMA;
//   * each softmax thread holds its whole score row in registers (one TMEM read per tile), takes the
//     row max with FMNMX3, rescales O lazily, writes fp16 P over the score columns and keeps the row
//     sum in a register;
//   * the two softmax warps that share an SM sub-partition (and its MUFU) take turns in their exp2
//     phase through an mbarrier token, handed over at 3/4 of the loop.
// kBlockN = 128 (default): one score buffer per query tile, S_X[t+1] issued right after P_X[t] V_t.
// kBlockN = 64: two score buffers per query tile, S_X[t+2] issued after P_X[t] V_t (more, smaller MMAs:
// measured slower at d = 40; kept selectable with TF_EXT_ATTN_MODE=pp64).
//   warp 0: TMA   warps 1,2: MMA issue for tile A / B   warp 3: spare   warps 4-7: softmax A   warps 8-11: softmax B
// TMEM: score buffers [0,256), O_A [256,320), O_B [320,384)
// ================================================================================================
constexpr int kPPStagesMax = 12;
struct AttnCtl2 {
  uint64_t q_full;
  uint64_t kv_full[kPPStagesMax];
  uint64_t kv_empty[kPPStagesMax];
  uint64_t s_full[2][2];
  uint64_t p_full[2][2];
  uint64_t xu_turn[2][4];    // [next tile X][SM sub-partition]: exp2-phase token passed between the two softmax
                             // warps that share a sub-partition (and therefore its MUFU)
  uint64_t pv_done[2][2];    // [tile X][t & 1]: two alternating barriers, so a softmax warp that runs two
                             // tiles ahead of the tensor pipe can still name "P V of tile t" unambiguously
  uint32_t tmem_base;
};

template <int kBlockN>
__global__ void __launch_bounds__(384, 1)
ext_attn_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const AttnTable tab, const AttnParams prm,
                   __half* __restrict__ out) {
  constexpr int kNBuf = 128 / kBlockN;                // score buffers per query tile (TMEM columns [0,256) in total)
  constexpr int kChunks = kBlockN / 32;
  constexpr int kQTileBytes = kBlockM * 128;          // one 128-query tile, 64-wide d chunk
  constexpr int kQBytes = 2 * kQTileBytes;
  constexpr int kOnesBytes = 0;
  constexpr int kTileBytes = kBlockN * 128;
  constexpr int kStageBytes = 2 * kTileBytes;
  constexpr int kOCol = 256;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;
  uint8_t* ring = smem + kQBytes + kOnesBytes;
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

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(&ctl->q_full, 1);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&ctl->kv_full[i], 1);
      mbar_init(&ctl->kv_empty[i], 2);
    }
    for (int x = 0; x < 2; ++x) {
      for (int b = 0; b < 2; ++b) {
        mbar_init(&ctl->s_full[x][b], 1);
        mbar_init(&ctl->p_full[x][b], 4);
      }
      mbar_init(&ctl->pv_done[x][0], 1);
      mbar_init(&ctl->pv_done[x][1], 1);
      for (int qd = 0; qd < 4; ++qd) mbar_init(&ctl->xu_turn[x][qd], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);     // warp-uniform for the compiler

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
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: one warp per query tile (warp 1 -> A, warp 2 -> B) ==============
    // Measured (profiles/r01_ext_attn_trace.md): with one issuer serving both query tiles the issue path
    // (8 small P V MMAs + 3 Q K^T MMAs + commits + barrier polls per tile and query tile) was the
    // bottleneck and the softmax warps spent half their time waiting for the next score tile.  Two
    // issuers run the two streams in parallel (all hazards — P_X[t] V before S_X[t+kNBuf] — are inside
    // one stream).  Control flow and operands stay warp-uniform (the whole warp runs the loop, values
    // derive from shfl-broadcast / kernel parameters) so that ptxas keeps descriptors in uniform
    // registers and emits back-to-back UTCHMMA instead of an ELECT/R2UR loop around every MMA.
    const int X = warp - 1;
    const uint32_t idesc_qk = umma_idesc_f16(128, kBlockN, 0);
    const uint32_t idesc_pv = umma_idesc_f16(128, (uint32_t)n_pv, 1);
    constexpr uint32_t hi_kmaj = umma_desc_hi(1024);
    const uint32_t q_lo = umma_desc_lo(smem_u32(q_smem + X * kQTileBytes), 16);
    const uint32_t ring_k_lo = umma_desc_lo(smem_u32(ring), 16);                       // K tile of stage 0
    const uint32_t ring_v_lo = umma_desc_lo(smem_u32(ring + kTileBytes), kTileBytes);  // V tile of stage 0
    constexpr uint32_t kStageStep = kStageBytes >> 4;
    const uint32_t o_tmem = tmem_base + kOCol + X * 64;
    // S_X[buf] = Q_X K^T against the K tile in ring stage `st` (called by the elected lane only)
    auto issue_qk = [&](int st, int buf) {
      const uint32_t k_lo = ring_k_lo + (uint32_t)st * kStageStep;
      const uint32_t s_tmem = tmem_base + (uint32_t)((X * kNBuf + buf) * kBlockN);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)               // head dim <= 64: at most 4 k-steps of 16
        if (ks < ksteps) tc_mma_ss_lh(s_tmem, q_lo + ks * 2, hi_kmaj, k_lo + ks * 2, hi_kmaj, idesc_qk, ks > 0 ? 1u : 0u);
      tc_commit(&ctl->s_full[X][buf]);
    };
    mbar_wait(&ctl->q_full, 0);
    // Tile B starts once tile A's first probabilities have arrived: that staggers the two softmax warps
    // of every SM sub-partition by about half a period instead of letting them convoy.
    if (X == 1) mbar_wait(&ctl->p_full[0][0], 0);
    // ring position of the next K tile to be used by a Q K^T (tile t + kNBuf in the main loop)
    int qk_stage = 0;
    uint32_t qk_phase = 0;
    for (int t0 = 0; t0 < kNBuf && t0 < T; ++t0) {
      mbar_wait(&ctl->kv_full[qk_stage], qk_phase);
      tc_fence_after_sync();
      if (elect_one()) issue_qk(qk_stage, t0);
      __syncwarp();
      if (++qk_stage == stages) { qk_stage = 0; qk_phase ^= 1; }
    }
    int stage = 0;                                 // ring position of tile t (its V tile feeds P V)
    for (int t = 0; t < T; ++t) {
      const int buf = t % kNBuf;                   // kNBuf is 1 or 2
      const bool refill = t + kNBuf < T;
      if (refill) mbar_wait(&ctl->kv_full[qk_stage], qk_phase);   // K tile of the refill, polled while idle anyway
      if (X == 0 && lane_id() == 0) TF_TRACE_EV(2, t, 0);
      mbar_wait(&ctl->p_full[X][buf], (uint32_t)((t / kNBuf) & 1));
      tc_fence_after_sync();
      if (X == 0 && lane_id() == 0) TF_TRACE_EV(2, t, 1);
      if (elect_one()) {
        const uint32_t v_lo = ring_v_lo + (uint32_t)stage * kStageStep;
        const uint32_t p_tmem = tmem_base + (uint32_t)((X * kNBuf + buf) * kBlockN);
#pragma unroll
        for (int k = 0; k < kBlockN / 16; ++k)      // O_X (+)= P_X[:, 16k:16k+16] V[16k:16k+16, :]
          tc_mma_ts_lh(o_tmem, p_tmem + k * 8, v_lo + k * 128, hi_kmaj, idesc_pv, (t > 0 || k > 0) ? 1u : 0u);
        tc_commit(&ctl->pv_done[X][t & 1]);
        tc_commit(&ctl->kv_empty[stage]);          // count 2: the stage is free once both tiles' MMAs retired
        if (refill) issue_qk(qk_stage, buf);        // refill the score buffer P_X[t] vacates
      }
      __syncwarp();
      if (X == 0 && lane_id() == 0) TF_TRACE_EV(2, t, 3);
      if (++stage == stages) stage = 0;
      if (refill && ++qk_stage == stages) { qk_stage = 0; qk_phase ^= 1; }
    }
  } else if (warp >= 4) {              // (warp 3 is a spare: it keeps the softmax warps aligned to TMEM lane quadrants)
    // ===================== softmax warps: tile X = (warp - 4) / 4 =====================
    const int X = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + (int)lane_id();
    const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
    const uint32_t o_addr = tmem_base + t_lane + kOCol + X * 64;
    const float sl2 = prm.scale_log2;
    float m_run = 0.f;
    float l_run = 0.f;       // row sum of the (unrounded) probabilities, same scale as O
    int slab_tile = 0;
    for (int t = 0; t < T; ++t) {
      const int valid = min(kBlockN, S - slab_tile * kBlockN);
      if (++slab_tile == tiles_per_slab) slab_tile = 0;
      const uint32_t s_addr = tmem_base + t_lane + (uint32_t)((X * kNBuf + (t % kNBuf)) * kBlockN);
      const bool tracer = (quad == 0 && lane_id() == 0);
      if (tracer) TF_TRACE_EV(X, t, 0);
      mbar_wait(&ctl->s_full[X][t % kNBuf], (uint32_t)((t / kNBuf) & 1));
      tc_fence_after_sync();
      if (tracer) TF_TRACE_EV(X, t, 1);
      uint32_t v[kChunks][32];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) tmem_ld32(s_addr + 32 * c, v[c]);
      tmem_wait_ld();
      if (tracer) TF_TRACE_EV(X, t, 2);
      if (valid < kBlockN) {                       // ragged last tile of a keyframe: mask the padding keys
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (32 * c + i >= valid) v[c][i] = 0xFF800000u;
      }
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};     // 4 independent FMNMX3 chains
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          mx[0] = fmax3(mx[0], __uint_as_float(v[c][i + 0]), __uint_as_float(v[c][i + 1]));
          mx[1] = fmax3(mx[1], __uint_as_float(v[c][i + 2]), __uint_as_float(v[c][i + 3]));
          mx[2] = fmax3(mx[2], __uint_as_float(v[c][i + 4]), __uint_as_float(v[c][i + 5]));
          mx[3] = fmax3(mx[3], __uint_as_float(v[c][i + 6]), __uint_as_float(v[c][i + 7]));
        }
      const float mt = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const float mt_s = mt * sl2;
      if (t == 0) {
        m_run = mt_s;
      } else {
        const bool need = mt_s > m_run + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(&ctl->pv_done[X][(t - 1) & 1], (uint32_t)(((t - 1) >> 1) & 1));   // P V of tile t-1 retired
          tc_fence_after_sync();
          const float m_new = fmaxf(m_run, mt_s);
          const float alpha = fast_exp2(m_run - m_new);
          m_run = m_new;
          l_run *= alpha;
          for (int c0 = 0; c0 < n_pv; c0 += 16) {
            uint32_t o[16];
            tmem_ld16(o_addr + c0, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(o_addr + c0, o);
          }
          tmem_wait_st();
        }
      }
      if (tracer) TF_TRACE_EV(X, t, 3);
      // exp2 phase, MUFU-bound: the two warps of a sub-partition take turns (A t, B t, A t+1, ...) instead
      // of running their exp2 loops concurrently at half speed each and then idling together while the
      // tensor pipe produces their next score tiles (profiles/r01_ext_attn_trace.md).
      if (X == 0) {
        if (t > 0) mbar_wait(&ctl->xu_turn[0][quad], (uint32_t)((t - 1) & 1));
      } else {
        mbar_wait(&ctl->xu_turn[1][quad], (uint32_t)(t & 1));
      }
      const float neg_m = -m_run;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      // Software-pipelined by hand: the exp2 of chunk c are issued interleaved with the row-sum / fp16 pack /
      // TMEM store of chunk c-1, so every consumer sits >= 32 instructions behind its MUFU.EX2 and one warp
      // alone keeps the MUFU pipe streaming (ptxas otherwise schedules "MUFU, MUFU, FADD of those two",
      // which stalls on the MUFU latency after every pair: profiles/r01_ext_attn_trace.md).
#pragma unroll
      for (int c = 0; c <= kChunks; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (c < kChunks) {
            v[c][2 * i] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(v[c][2 * i]), sl2, neg_m)));
            v[c][2 * i + 1] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(v[c][2 * i + 1]), sl2, neg_m)));
          }
          if (c > 0) {
            const float p0 = __uint_as_float(v[c - 1][2 * i]), p1 = __uint_as_float(v[c - 1][2 * i + 1]);
            ls[i & 3] += p0 + p1;
            pk[i] = pack_f16x2_rn(p0, p1);
          }
        }
        if (c > 0) tmem_st16(s_addr + 16 * (c - 1), pk);
        if (c == prm.handoff) {      // hand the MUFU over while the later chunks are still in flight: their tail
          __syncwarp();              // (dependent FADD / F2FP / TMEM store latencies) overlaps the other warp's start
          if (lane_id() == 0) mbar_arrive(&ctl->xu_turn[1 - X][quad]);
        }
      }
      l_run += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      if (tracer) TF_TRACE_EV(X, t, 4);
      tmem_wait_st();
      if (tracer) TF_TRACE_EV(X, t, 5);
      tc_fence_before_sync();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&ctl->p_full[X][t % kNBuf]);
      if (tracer) TF_TRACE_EV(X, t, 6);
    }
    // ---- final: O / L -> fp16 ----
    mbar_wait(&ctl->pv_done[X][(T - 1) & 1], (uint32_t)(((T - 1) >> 1) & 1));
    tc_fence_after_sync();
    const float inv_l = 1.0f / l_run;
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

template <int kBlockN>
int launch_pp(const void* q, const void* k, const void* v, long long q_tok_stride, long long kv_tok_stride,
              int q_samples_total, int kv_samples_total, const AttnTable& tab, int n_out, int S, int heads, int d,
              float scale, void* out, cudaStream_t stream) {
  constexpr int kQBytes = 2 * kBlockM * 128, kOnesBytes = 0, kStageBytes = 2 * kBlockN * 128;
  int stages = (227 * 1024 - 2048 - kQBytes - kOnesBytes) / kStageBytes;
  if (stages > kPPStagesMax) stages = kPPStagesMax;
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
  if (int e = make(&map_k, k, kv_tok_stride, kv_samples_total, kBlockN)) return e;
  if (int e = make(&map_v, v, kv_tok_stride, kv_samples_total, kBlockN)) return e;
  AttnParams prm;
  prm.S = S; prm.heads = heads; prm.d = d; prm.n_out = n_out;
  prm.tiles_m = (S + 2 * kBlockM - 1) / (2 * kBlockM);
  static const char* env_handoff = getenv("TF_EXT_ATTN_HANDOFF");     // tuning knob (profiling)
  prm.handoff = env_handoff ? atoi(env_handoff) : (kBlockN / 32 - 2);
  prm.stages = stages;
  prm.scale_log2 = scale * 1.4426950408889634f;
  prm.out_tok_stride = (long long)heads * d;
  auto kern = ext_attn_pp_kernel<kBlockN>;
  if (check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes),
                 "tf_ext_attn smem attribute"))
    return TF_ERR_CUDA;
  const long long grid = (long long)n_out * heads * prm.tiles_m;
  kern<<<(unsigned)grid, 384, smem_bytes, stream>>>(map_q, map_k, map_v, tab, prm, static_cast<__half*>(out));
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
  // A/B switch for profiling: TF_EXT_ATTN_MODE = v1 (one query tile per CTA) | pp128 | pp64
  static const char* mode = getenv("TF_EXT_ATTN_MODE");
  const bool force_v1 = mode && mode[0] == 'v';
  if (d <= 64 && S > 128 && !force_v1) {
    if (mode && mode[2] == '6')
      return launch_pp<64>(q, k, v, q_tok_stride, kv_tok_stride, q_samples_total, kv_samples_total, tab, n_out, S,
                           heads, d, scale, out, stream);
    return launch_pp<128>(q, k, v, q_tok_stride, kv_tok_stride, q_samples_total, kv_samples_total, tab, n_out, S,
                          heads, d, scale, out, stream);
  }
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

#ifdef TF_TRACE
int read_attn_trace(long long* host, int n) {
  const int total = 3 * kTraceTiles * kTraceEvents;
  if (n > total) n = total;
  return check_cuda(cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(long long) * n), "trace read");
}
#endif

}  // namespace tf

#ifdef TF_TRACE
extern "C" int tf_debug_read_attn_trace(long long* host, int n) { return tf::read_attn_trace(host, n); }
#endif
