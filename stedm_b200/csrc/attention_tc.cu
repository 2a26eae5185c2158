// K5: fused flash-style self-attention on tcgen05 for the U-Net AttentionBlock (QKVAttentionLegacy,
// ldm/modules/diffusionmodules/openaimodel.py:378-394): 8 heads x 128 (or 64) channels, T = (L/4)^2 tokens.
// The B*heads*T^2 score matrix the reference materialises in fp32 never leaves the SM.
//
// One CTA = one (sample, head, 128-query tile).  Per 128-key tile j:
//   S = Q K_j^T        tcgen05.mma, A = Q (smem, K-major SW128), B = K_j (smem, K-major SW128), D = S in TMEM (fp32)
//   softmax            4 warps, one query row per thread (TMEM lane == row, so row max / row sum need no shuffles):
//                      online max / sum in fp32, P = exp2((S - m) * scale * log2e) rounded to bf16 into shared
//                      memory in the SW128 K-major layout; the running O in TMEM is rescaled by exp2(m_old - m_new)
//   O += P V_j         tcgen05.mma, A = P (smem), B = V_j (smem as loaded by TMA = MN-major SW128), D = O in TMEM
// Epilogue: O / l -> bf16 -> out[b, t, head*d + c].
// Warp roles: warp 0 = TMA producer (Q once, then a 2-stage K/V ring), warp 1 = TMEM alloc + MMA issuer,
// warps 2-5 = softmax / correction / epilogue.  Keys >= T are masked to -inf; query rows >= T are not stored
// (TMA zero-fills them), so T need not be a multiple of 128.
#include "../../include/stedm_b200.h"
#include <stdlib.h>

#include "common.cuh"

using namespace stedm;

namespace {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per tile
constexpr int AT_THREADS = 192;
constexpr int AT_STAGES = 2;

struct AttnParams {
  __nv_bfloat16* out;
  int tokens, heads, head_dim;
  float scale_log2e;  // scale * log2(e)
  long long out_stride_b;  // elements between samples of `out`
  int mask_diag;      // 1: a token does not attend to itself (sViT's LSA, networks/vit_set.py:52-54)
  int tokens_kv;      // keys / values per sample (== tokens for self-attention; the context length for cross-attention)
};

// smem descriptor for an MN-major SW128 operand whose K rows are 128 B apart (as TMA writes a [rows][64] bf16 box):
// LBO = byte distance between consecutive 64-element blocks along MN, SBO = byte distance between 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor with B MN-major (bit 16).
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) {
  return umma_idesc_bf16(M, N) | (1u << 16);
}

template <int HD>
struct AttnCfg {
  static constexpr int DB = HD / 64;                       // 64-channel boxes per token row
  static constexpr int Q_BYTES = AT_BM * HD * 2;
  static constexpr int K_BYTES = AT_BN * HD * 2;
  static constexpr int V_BYTES = AT_BN * HD * 2;
  static constexpr int P_BYTES = AT_BM * AT_BN * 2;
  static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + AT_STAGES * STAGE_BYTES + P_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 256;                    // S: 128 columns, O: HD columns
};

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  using Cfg = AttnCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = s_q + Cfg::Q_BYTES;
  uint8_t* s_p = s_kv + AT_STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + Cfg::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;               // [AT_STAGES]
  uint64_t* kv_empty = kv_full + AT_STAGES;   // [AT_STAGES]
  uint64_t* s_full = kv_empty + AT_STAGES;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM, head = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.tokens_kv + AT_BN - 1) / AT_BN;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < AT_STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;          // columns [0, 128)
  const uint32_t tmem_o = tmem_base + 128;    // columns [128, 128 + HD)

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int d = 0; d < Cfg::DB; ++d) tma_load_4d(s_q + d * (AT_BM * 128), &map_q, q_full, d * 64, q0, head, b);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % AT_STAGES;
        const uint32_t ph = (j / AT_STAGES) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], Cfg::STAGE_BYTES);
        uint8_t* sk = s_kv + s * Cfg::STAGE_BYTES;
        uint8_t* sv = sk + Cfg::K_BYTES;
#pragma unroll
        for (int d = 0; d < Cfg::DB; ++d) {
          tma_load_4d(sk + d * (AT_BN * 128), &map_k, &kv_full[s], d * 64, j * AT_BN, head, b);
          tma_load_4d(sv + d * (AT_BN * 128), &map_v, &kv_full[s], d * 64, j * AT_BN, head, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(AT_BM, HD);
      mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % AT_STAGES;
        const uint32_t ph = (j / AT_STAGES) & 1;
        mbar_wait(&kv_full[s], ph);
        tc_fence_after();
        const uint32_t sk = smem_u32(s_kv + s * Cfg::STAGE_BYTES);
        const uint32_t sv = sk + Cfg::K_BYTES;
        // S = Q K^T : K (reduction) = head_dim, 16 per MMA; box d covers channels [64d, 64d+64)
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t off = (k / 4) * (AT_BM * 128) + (k % 4) * 32;
          umma_bf16(tmem_s, umma_desc_sw128(smem_u32(s_q) + off), umma_desc_sw128(sk + off), idesc_s, k != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        // O += P V : reduction = 128 keys, 16 per MMA; P slab kk/4 holds keys [64*(kk/4), +64)
        mbar_wait(p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < AT_BN / 16; ++kk) {
          const uint32_t poff = (kk / 4) * (AT_BM * 128) + (kk % 4) * 32;
          const uint64_t vdesc = umma_desc_mn_sw128(sv + kk * 16 * 128, AT_BN * 128, 1024);
          umma_bf16(tmem_o, umma_desc_sw128(smem_u32(s_p) + poff), vdesc, idesc_o, (j | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(o_full);
      }
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                      // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int valid_keys = min(AT_BN, p.tokens_kv - j * AT_BN);
      // LSA diagonal mask: column of this tile that holds the query's own key (out of range when not in the tile)
      const int self_col = p.mask_diag ? (q0 + row - j * AT_BN) : -1;
      // pass 1: row max
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid_keys && c + i != self_col) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, m_tile);
      // exp2(-inf) = 0 on the first tile; a tile whose only live key is the masked diagonal leaves m at -inf
      const float alpha = m_new == -INFINITY ? 1.0f : exp2f((m_run - m_new) * p.scale_log2e);
      // previous P V must be complete before P is overwritten and O is rescaled
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_o + lane_addr + c, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st32(tmem_o + lane_addr + c, r);
        }
        tmem_st_wait();
      }
      // pass 2: P = exp2((S - m) * scale*log2e) -> bf16 -> shared memory (K-major SW128: 16 B chunk ^ (row % 8))
      float l_tile = 0.f;
      const float mb = m_new == -INFINITY ? 0.f : m_new * p.scale_log2e;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c + i < valid_keys && c + i != self_col)
                         ? exp2f(__uint_as_float(r[i]) * p.scale_log2e - mb) : 0.f;
          float p1 = (c + i + 1 < valid_keys && c + i + 1 != self_col)
                         ? exp2f(__uint_as_float(r[i + 1]) * p.scale_log2e - mb) : 0.f;
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          // the row sum uses the bf16-rounded probabilities that the P V product actually consumes
          l_tile += __low2float(h2) + __high2float(h2);
          pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        uint8_t* slab = s_p + (c / 64) * (AT_BM * 128) + (row / 8) * 1024 + (row % 8) * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = ((c % 64) / 8 + g) ^ (row % 8);
          *reinterpret_cast<uint4*>(slab + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      fence_proxy_async_smem();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();          // orders this thread's TMEM loads/stores before the arrive
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 -> out[b, q0 + row, head*HD + c]
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const int t = q0 + row;
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* dst = p.out + static_cast<size_t>(b) * p.out_stride_b + static_cast<size_t>(t) * (p.heads * HD) + head * HD;
#pragma unroll 1
    for (int c = 0; c < HD; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_addr + c, r);
      tmem_ld_wait();
      if (t < p.tokens) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c + i) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// Wide single-head attention (the VAE decoder's AttnBlock, model.py:178-202: one head, d = C = 512, T = L^2 tokens):
// S + O for d = 512 need 128 + 512 TMEM columns and Q + one K tile 256 KB of shared memory, so the work is cut
// differently.  One CTA = (sample, 128-query tile, HALF of the output channels):
//   * Q (128 x 512 bf16 = 128 KB) stays resident as eight K-major 64-channel slabs;
//   * K and V stream through a 96 KB ring managed in 8 KB units: per key tile eight K slabs [128 keys x 64 channels]
//     (S accumulates over them), then the V slabs of this CTA's 256 output channels;
//   * S is double-buffered in TMEM (2 x 128 columns) next to the 256 O columns: Q K^T of tile j + 1 runs while the
//     softmax warps work on tile j; P (bf16) is written back over the S columns it came from and feeds the P V product
//     as a TMEM A operand; the running O is rescaled in TMEM only when some row's maximum moved (warp-uniform test).
// The two CTAs of a query tile's channel halves both compute S (1.5x the FLOPs of an ideal kernel, no T x T tensor
// anywhere, no traffic between them).
// PAIR (cta_group::2, the default whenever there are two query tiles): the two CTAs of a cluster take NEIGHBOURING query
// tiles of the same sample and channel half and form one 256-row MMA: every K / V slab is shared — each CTA stages only
// its half (64 of the 128 keys of a K slab; 64 of the 128 channels of a V slab), so the ring holds a whole key tile
// (twice the prefetch distance in time) and half the bytes cross L2 -> SM.  The single-CTA version was latency-bound:
// 96 KB in flight against a 1.5-2.5 us TMA round trip under load (577 TFLOP/s useful).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AW_D = 512;              // head dim (q / k channels)
constexpr int AW_OD = 256;             // output channels per CTA
constexpr int AW_UNIT = 8192;          // ring unit
constexpr int AW_UNITS = 12;           // 96 KB
constexpr int AW_SLAB = AT_BN * 128;   // 16 KB: 128 rows x 64 bf16
constexpr int AW_Q_BYTES = AT_BM * AW_D * 2;
constexpr int AW_THREADS = 320;         // producer, MMA, and EIGHT softmax warps: two threads per query row (64 keys each)
constexpr int AW_SMEM_BYTES = AW_Q_BYTES + AW_UNITS * AW_UNIT + 1024 + 256 + 1024;   // + barriers + row-max exchange

template <bool PAIR>
__global__ void __launch_bounds__(AW_THREADS, 1)
attention_wide_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  constexpr int KU = PAIR ? 1 : 2;                 // ring units per K slab staged in this CTA
  constexpr int VN = PAIR ? 128 : 64;              // output channels per V slab (MMA N of the P V product)
  constexpr int VS = AW_OD / VN;                   // V slabs per key tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_ring = s_q + AW_Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + AW_UNITS * AW_UNIT);
  uint64_t* q_full = bars;
  uint64_t* r_full = bars + 1;               // [AW_UNITS]: an entry signals the barrier of its FIRST unit
  uint64_t* r_empty = r_full + AW_UNITS;     // [AW_UNITS]: every unit of an entry is released
  uint64_t* s_full = r_empty + AW_UNITS;     // [2]
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* s_red = reinterpret_cast<float*>(s_ring + AW_UNITS * AW_UNIT + 256);   // [2 column halves][128 rows]
  static_assert((1 + 2 * AW_UNITS + 2 + 1 + 1) * 8 + 4 <= 256, "barrier area");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int q0 = blockIdx.x * AT_BM, half = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.tokens_kv + AT_BN - 1) / AT_BN;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, PAIR ? 2 : 1);           // PAIR: the leader's "landed" barriers collect both CTAs' producers
    for (int i = 0; i < AW_UNITS; ++i) {
      mbar_init(&r_full[i], PAIR ? 2 : 1);
      mbar_init(&r_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, PAIR ? 512 : 256);       // every softmax thread (of both CTAs) arrives
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_slot, 512);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;    // S buffers: columns [0,128) and [128,256); O: [256, 512)

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      auto arrive = [&](uint64_t* bar, uint32_t bytes) {
        if constexpr (PAIR) mbar_arrive_expect_tx_cluster(bar, bytes, 0);
        else mbar_arrive_expect_tx(bar, bytes);
      };
      auto load = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int ch, int tok) {
        if constexpr (PAIR) tma_load_4d_2sm(dst, m, bar, ch, tok, 0, b);
        else tma_load_4d(dst, m, bar, ch, tok, 0, b);
      };
      arrive(q_full, AW_Q_BYTES);
#pragma unroll
      for (int d = 0; d < AW_D / 64; ++d) load(s_q + d * AW_SLAB, &map_q, q_full, d * 64, q0);
      uint32_t cur = 0, fills = 0;             // ring cursor (units); bit u of `fills` = parity of the fills of unit u
      auto put = [&](const CUtensorMap* m, int ch, int tok, int units) {
        for (int u = 0; u < units; ++u) {      // every unit of the entry must have been released
          mbar_wait(&r_empty[cur + u], ((fills >> (cur + u)) & 1u) ^ 1u);
          fills ^= 1u << (cur + u);
        }
        arrive(&r_full[cur], units * AW_UNIT);
        load(s_ring + cur * AW_UNIT, m, &r_full[cur], ch, tok);
        cur += units;
        if (cur == AW_UNITS) cur = 0;
      };
      // PAIR: this CTA stages keys [64 r, 64 r + 64) of a K slab (map_k's box is 64 keys tall) and channels
      // [.. + 64 r, .. + 64 r + 64) of a 128-channel V slab
      const int ktok = PAIR ? static_cast<int>(cta_rank) * 64 : 0;
      const int vch = half * AW_OD + (PAIR ? static_cast<int>(cta_rank) * 64 : 0);
      for (int d = 0; d < AW_D / 64; ++d) put(&map_k, d * 64, ktok, KU);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv)
          for (int d = 0; d < AW_D / 64; ++d) put(&map_k, d * 64, (j + 1) * AT_BN + ktok, KU);
        for (int d = 0; d < VS; ++d) put(&map_v, vch + d * VN, j * AT_BN, 2);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if ((!PAIR || cta_rank == 0) && elect_one()) {   // PAIR: only the leader CTA issues (for both)
      constexpr uint32_t idesc_s = umma_idesc_bf16(PAIR ? 2 * AT_BM : AT_BM, AT_BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(PAIR ? 2 * AT_BM : AT_BM, VN);
      uint32_t cur = 0, starts = 0;            // bit u of `starts` = parity of the entries that began at unit u
      auto take = [&]() {                      // wait for the entry at the cursor; returns its shared address
        mbar_wait(&r_full[cur], (starts >> cur) & 1u);
        starts ^= 1u << cur;
        tc_fence_after();
        return smem_u32(s_ring + cur * AW_UNIT);
      };
      auto release = [&](int units) {          // hand the entry's units back once the MMAs issued so far have read them
        for (int u = 0; u < units; ++u) {
          if constexpr (PAIR) umma_commit_2sm_mcast(&r_empty[cur + u], 3);
          else umma_commit(&r_empty[cur + u]);
        }
        cur += units;
        if (cur == AW_UNITS) cur = 0;
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_2sm_mcast(bar, 3);
        else umma_commit(bar);
      };
      mbar_wait(q_full, 0);
      auto qk = [&](int j) {      // S[j & 1] = Q K_j^T over the eight channel slabs
        const uint32_t tmem_s = tmem_base + (j & 1) * 128;
        for (int d = 0; d < AW_D / 64; ++d) {
          const uint32_t sk = take(), sq = smem_u32(s_q + d * AW_SLAB);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (d | k) != 0 ? 1u : 0u;
            if constexpr (PAIR) umma_bf16_2sm(tmem_s, umma_desc_sw128(sq + k * 32), umma_desc_sw128(sk + k * 32), idesc_s, acc);
            else umma_bf16(tmem_s, umma_desc_sw128(sq + k * 32), umma_desc_sw128(sk + k * 32), idesc_s, acc);
          }
          release(KU);
        }
        commit(&s_full[j & 1]);
      };
      qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) qk(j + 1);          // runs while the softmax warps are busy with tile j
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        for (int d = 0; d < VS; ++d) {        // O[:, VN d ..] += P V_j[:, VN d ..]: reduction over the 128 keys
          const uint32_t sv = take();
#pragma unroll
          for (int kk = 0; kk < AT_BN / 16; ++kk) {
            const uint64_t vdesc = umma_desc_mn_sw128(sv + kk * 16 * 128, AW_SLAB, 1024);
            // A = P from TMEM: 16 keys = 8 packed columns of the S buffer P was written over
            const uint32_t pa = tmem_base + (j & 1) * 128 + kk * 8, acc = (j | kk) != 0 ? 1u : 0u;
            if constexpr (PAIR) umma_bf16_ts_2sm(tmem_o + d * VN, pa, vdesc, idesc_o, acc);
            else umma_bf16_ts(tmem_o + d * VN, pa, vdesc, idesc_o, acc);
          }
          release(2);
        }
        commit(o_full);
      }
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    // Eight warps: warps 2-5 take keys [0, 64) of the tile, warps 6-9 keys [64, 128) of the SAME rows (a warp reaches the
    // TMEM lanes of quadrant warp % 4), so two warps share each scheduler and cover each other's TMEM-load and MUFU
    // latency — with four warps the softmax (~3000 clocks per key tile) was as long as the tile's MMAs and the kernel
    // ran at half the tensor rate.  A thread keeps its 64 scores in registers between the max and the exp pass.
    const int quad = warp & 3, h = (warp - 2) >> 2;
    const int row = quad * 32 + lane;                      // query row within the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;                  // l_run: this thread's half of the row sum
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t tmem_s = tmem_base + (j & 1) * 128;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int valid_keys = min(AT_BN, p.tokens_kv - j * AT_BN) - h * 64;   // valid keys among this thread's 64
      uint32_t sv[64];
      {
        uint32_t (&lo)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
        uint32_t (&hi)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);
        tmem_ld32(tmem_s + lane_addr + h * 64, lo);
        tmem_ld32(tmem_s + lane_addr + h * 64 + 32, hi);
        tmem_ld_wait();
      }
      float m_part = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i < valid_keys) m_part = fmaxf(m_part, __uint_as_float(sv[i]));
      s_red[h * AT_BM + row] = m_part;
      named_bar_sync(1, 256);      // also: every thread's S loads are complete -> P may be written over the S columns
      const float m_new = fmaxf(m_run, fmaxf(m_part, s_red[(1 - h) * AT_BM + row]));
      named_bar_sync(2, 256);      // the exchange buffer may be rewritten
      const float alpha = m_new == -INFINITY ? 1.0f : exp2f((m_run - m_new) * p.scale_log2e);
      // the previous P V must be complete before O is rescaled
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {      // the running maximum of some row of this warp moved
#pragma unroll 1
          for (int c = h * (AW_OD / 2); c < (h + 1) * (AW_OD / 2); c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_o + lane_addr + c, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(tmem_o + lane_addr + c, r);
          }
          tmem_st_wait();
        }
      }
      float l_tile = 0.f;
      const float mb = m_new == -INFINITY ? 0.f : m_new * p.scale_log2e;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        float p0 = (i < valid_keys) ? exp2f(__uint_as_float(sv[i]) * p.scale_log2e - mb) : 0.f;
        float p1 = (i + 1 < valid_keys) ? exp2f(__uint_as_float(sv[i + 1]) * p.scale_log2e - mb) : 0.f;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
        l_tile += __low2float(h2) + __high2float(h2);     // the sum of what the P V product actually consumes
        pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      // P (bf16 pairs) over columns [32 h, 32 h + 32) of this S buffer: every thread's scores are in registers by now
      tmem_st32(tmem_s + lane_addr + h * 32, pk);
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      tmem_st_wait();
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(p_full, 0);   // the leader's MMA thread waits for both CTAs' P
      else mbar_arrive(p_full);
    }
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    s_red[h * AT_BM + row] = l_run;
    named_bar_sync(1, 256);
    const float inv_l = 1.0f / (l_run + s_red[(1 - h) * AT_BM + row]);
    const int t = q0 + row;
    __nv_bfloat16* dst = p.out + static_cast<size_t>(b) * p.out_stride_b + static_cast<size_t>(t) * AW_D + half * AW_OD;
#pragma unroll 1
    for (int c = h * (AW_OD / 2); c < (h + 1) * (AW_OD / 2); c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_addr + c, r);
      tmem_ld_wait();
      if (t < p.tokens) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(r[i + 4]) * inv_l, __uint_as_float(r[i + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(r[i + 6]) * inv_l, __uint_as_float(r[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c + i) = u;
        }
      }
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // no CTA may exit while its peer can still signal it
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

int make_qkv_map(CUtensorMap* m, const void* base, int batch, int heads, int tokens, int head_dim, long long sb,
                 long long sh, long long st, int box_rows = AT_BM) {
  // dims innermost-first: channel, token, head, batch
  const uint64_t dims[4] = {static_cast<uint64_t>(head_dim), static_cast<uint64_t>(tokens),
                            static_cast<uint64_t>(heads), static_cast<uint64_t>(batch)};
  const uint64_t head_stride = heads > 1 ? static_cast<uint64_t>(sh) * 2 : static_cast<uint64_t>(st) * 2 * tokens;
  const uint64_t batch_stride = batch > 1 ? static_cast<uint64_t>(sb) * 2 : static_cast<uint64_t>(st) * 2 * tokens;
  const uint64_t str[3] = {static_cast<uint64_t>(st) * 2, head_stride, batch_stride};
  const uint32_t box[4] = {64, static_cast<uint32_t>(box_rows), 1, 1};
  return make_tmap_bf16(m, base, 4, dims, str, box);
}

template <int HD>
int launch_attn(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const AttnParams& p, int batch,
                cudaStream_t stream) {
  using Cfg = AttnCfg<HD>;
  static DeviceOnce configured;
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("attention_tc: cudaFuncSetAttribute(%d B smem): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return ERR_CUDA;
    }
    configured.done();
  }
  dim3 grid((p.tokens + AT_BM - 1) / AT_BM, p.heads, batch);
  attention_tc_kernel<HD><<<grid, AT_THREADS, Cfg::SMEM_BYTES, stream>>>(mq, mk, mv, p);
  return check_launch("attention_tc");
}

}  // namespace

extern "C" int stedm_attention_tc(const void* q, const void* k, const void* v, void* out, int batch, int heads,
                                  int tokens, int head_dim, long long stride_b, long long stride_h,
                                  long long stride_t, float scale, long long out_stride_b, int mask_diag,
                                  int tokens_kv, long long kv_stride_b, long long kv_stride_h, long long kv_stride_t,
                                  void* stream) {
  STEDM_REQUIRE(q && k && v && out, "attention_tc: null pointer");
  STEDM_REQUIRE(head_dim == 64 || head_dim == 128 || (head_dim == AW_D && heads == 1 && !mask_diag),
                "attention_tc: head_dim %d unsupported (64 or 128 per head, or one 512-wide head)", head_dim);
  STEDM_REQUIRE(batch > 0 && heads > 0 && tokens > 0 && batch <= 65535 && heads <= 65535, "attention_tc: bad shape");
  STEDM_REQUIRE(stride_t % 8 == 0 && stride_h % 8 == 0 && stride_b % 8 == 0,
                "attention_tc: strides must be multiples of 8 elements (16 bytes)");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_qkv_map(&mq, q, batch, heads, tokens, head_dim, stride_b, stride_h, stride_t))) return rc;
  // cross-attention: keys / values come from a context of tokens_kv tokens with its own strides (0 => self-attention)
  if (tokens_kv <= 0) {
    tokens_kv = tokens; kv_stride_b = stride_b; kv_stride_h = stride_h; kv_stride_t = stride_t;
  }
  STEDM_REQUIRE(kv_stride_t % 8 == 0 && kv_stride_h % 8 == 0 && kv_stride_b % 8 == 0 && !(mask_diag && tokens_kv != tokens),
                "attention_tc: bad key/value strides or a diagonal mask on cross-attention");
  if ((rc = make_qkv_map(&mk, k, batch, heads, tokens_kv, head_dim, kv_stride_b, kv_stride_h, kv_stride_t))) return rc;
  if ((rc = make_qkv_map(&mv, v, batch, heads, tokens_kv, head_dim, kv_stride_b, kv_stride_h, kv_stride_t))) return rc;
  AttnParams p;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.tokens = tokens; p.heads = heads; p.head_dim = head_dim;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out_stride_b = out_stride_b > 0 ? out_stride_b : static_cast<long long>(tokens) * heads * head_dim;
  p.mask_diag = mask_diag ? 1 : 0;
  p.tokens_kv = tokens_kv;
  STEDM_REQUIRE(p.out_stride_b % 8 == 0 && !(mask_diag && tokens < 2), "attention_tc: bad out stride / diagonal mask");
  auto s = static_cast<cudaStream_t>(stream);
  if (head_dim == AW_D) {
    // PAIR (cta_group::2) whenever there are at least two query tiles; an odd last tile pairs with an all-out-of-range
    // one (zero-filled loads, rows not stored)
    const int q_tiles = (tokens + AT_BM - 1) / AT_BM;
    static const bool pair_enabled = [] { const char* e = getenv("STEDM_AW_PAIR"); return !(e && e[0] == '0'); }();
    const bool pair = pair_enabled && q_tiles >= 2;
    if (pair && (rc = make_qkv_map(&mk, k, batch, heads, tokens_kv, head_dim, kv_stride_b, kv_stride_h, kv_stride_t, 64)))
      return rc;                                           // K half-slabs: 64 keys per CTA
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(attention_wide_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM_BYTES);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_wide_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM_BYTES);
      if (e != cudaSuccess) {
        set_error("attention_tc: cudaFuncSetAttribute(%d B smem): %s", AW_SMEM_BYTES, cudaGetErrorString(e));
        return ERR_CUDA;
      }
      configured.done();
    }
    if (!pair) {
      dim3 grid(q_tiles, AW_D / AW_OD, batch);
      attention_wide_kernel<false><<<grid, AW_THREADS, AW_SMEM_BYTES, s>>>(mq, mk, mv, p);
      return check_launch("attention_tc (wide)");
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((q_tiles + 1) / 2 * 2, AW_D / AW_OD, batch);
    cfg.blockDim = dim3(AW_THREADS);
    cfg.dynamicSmemBytes = AW_SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, attention_wide_kernel<true>, mq, mk, mv, p);
    if (e != cudaSuccess) {
      set_error("attention_tc: launch failed: %s", cudaGetErrorString(e));
      return ERR_CUDA;
    }
    return check_launch("attention_tc (wide, CTA pairs)");
  }
  return head_dim == 128 ? launch_attn<128>(mq, mk, mv, p, batch, s) : launch_attn<64>(mq, mk, mv, p, batch, s);
}
