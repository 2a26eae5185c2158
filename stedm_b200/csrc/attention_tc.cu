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
// Wide single-head attention (the VAE decoder's AttnBlock, model.py:178-202: one head, d = C = 512, T = (4L/4)^2 ... L^2
// tokens): S + O for d = 512 need 128 + 512 TMEM columns and Q + one K tile 256 KB of shared memory, so the work is cut
// differently.  One CTA = (sample, 128-query tile, HALF of the output channels):
//   * Q (128 x 512 bf16 = 128 KB) stays resident as eight K-major 64-channel slabs;
//   * K and V stream through a ring of four 16 KB slabs [128 keys x 64 channels]: eight K slabs per key tile (S accumulates
//     over them, four N = 128 MMAs each), then the four V slabs of this CTA's 256 output channels (eight N = 64 MMAs each
//     into their own 64 TMEM columns) — slab groups are multiples of the ring size, so the V slabs always sit in slots 0-3;
//   * S is double-buffered in TMEM (2 x 128 columns) next to the 256 O columns: Q K^T of tile j + 1 runs while the softmax
//     warps work on tile j; the running O is rescaled in TMEM only when some row's maximum moved (warp-uniform test).
// The two CTAs of a query tile both compute S (1.5x the FLOPs of an ideal kernel, no T x T tensor anywhere, no
// inter-CTA traffic); at 64 B of K / V per tensor-pipe clock the kernel sits at the SM's operand-ingest limit.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AW_D = 512;              // head dim (q / k channels)
constexpr int AW_OD = 256;             // output channels per CTA
#ifndef STEDM_AW_P_TMEM
#define STEDM_AW_P_TMEM 1   // P (bf16 probabilities) stays in TMEM, aliased over the S buffer it came from, and feeds the
#endif                      // P V product as a TMEM A operand: no 32 KB staging buffer -> a six-slab K / V ring
constexpr bool AW_P_TMEM = STEDM_AW_P_TMEM != 0;
constexpr int AW_RING = AW_P_TMEM ? 6 : 4;
constexpr int AW_SLAB = AT_BN * 128;   // 16 KB: 128 rows x 64 bf16
constexpr int AW_Q_BYTES = AT_BM * AW_D * 2;
constexpr int AW_P_BYTES = AW_P_TMEM ? 0 : AT_BM * AT_BN * 2;
constexpr int AW_SMEM_BYTES = AW_Q_BYTES + AW_RING * AW_SLAB + AW_P_BYTES + 1024 + 256;

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_wide_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                      const __grid_constant__ CUtensorMap map_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_ring = s_q + AW_Q_BYTES;
  uint8_t* s_p = s_ring + AW_RING * AW_SLAB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + AW_P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* r_full = bars + 1;               // [AW_RING]
  uint64_t* r_empty = r_full + AW_RING;      // [AW_RING]
  uint64_t* s_full = r_empty + AW_RING;      // [2]
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AT_BM, half = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.tokens_kv + AT_BN - 1) / AT_BN;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < AW_RING; ++i) {
      mbar_init(&r_full[i], 1);
      mbar_init(&r_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;    // S buffers: columns [0,128) and [128,256); O: [256, 512)

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, AW_Q_BYTES);
#pragma unroll
      for (int d = 0; d < AW_D / 64; ++d) tma_load_4d(s_q + d * AW_SLAB, &map_q, q_full, d * 64, q0, 0, b);
      uint32_t slot = 0, ph = 0;
      auto put = [&](const CUtensorMap* m, int ch, int tok) {
        mbar_wait(&r_empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&r_full[slot], AW_SLAB);
        tma_load_4d(s_ring + slot * AW_SLAB, m, &r_full[slot], ch, tok, 0, b);
        if (++slot == AW_RING) { slot = 0; ph ^= 1; }
      };
      for (int d = 0; d < AW_D / 64; ++d) put(&map_k, d * 64, 0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv)
          for (int d = 0; d < AW_D / 64; ++d) put(&map_k, d * 64, (j + 1) * AT_BN);
        for (int d = 0; d < AW_OD / 64; ++d) put(&map_v, half * AW_OD + d * 64, j * AT_BN);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN);
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(AT_BM, 64);
      uint32_t slot = 0, ph = 0;
      mbar_wait(q_full, 0);
      auto qk = [&](int j) {      // S[j & 1] = Q K_j^T over the eight channel slabs
        const uint32_t tmem_s = tmem_base + (j & 1) * 128;
        for (int d = 0; d < AW_D / 64; ++d) {
          mbar_wait(&r_full[slot], ph);
          tc_fence_after();
          const uint32_t sk = smem_u32(s_ring + slot * AW_SLAB), sq = smem_u32(s_q + d * AW_SLAB);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_s, umma_desc_sw128(sq + k * 32), umma_desc_sw128(sk + k * 32), idesc_s, (d | k) != 0 ? 1u : 0u);
          umma_commit(&r_empty[slot]);
          if (++slot == AW_RING) { slot = 0; ph ^= 1; }
        }
        umma_commit(&s_full[j & 1]);
      };
      qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) qk(j + 1);          // runs while the softmax warps are busy with tile j
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        for (int d = 0; d < AW_OD / 64; ++d) {  // O[:, 64 d ..] += P V_j[:, 64 d ..]: reduction over the 128 keys
          mbar_wait(&r_full[slot], ph);
          tc_fence_after();
          const uint32_t sv = smem_u32(s_ring + slot * AW_SLAB);
#pragma unroll
          for (int kk = 0; kk < AT_BN / 16; ++kk) {
            const uint64_t vdesc = umma_desc_mn_sw128(sv + kk * 16 * 128, AW_SLAB, 1024);
            if constexpr (AW_P_TMEM) {   // A = P from TMEM: 16 keys = 8 packed columns of the S buffer P was written over
              umma_bf16_ts(tmem_o + d * 64, tmem_base + (j & 1) * 128 + kk * 8, vdesc, idesc_o, (j | kk) != 0 ? 1u : 0u);
            } else {
              const uint32_t poff = (kk / 4) * AW_SLAB + (kk % 4) * 32;
              umma_bf16(tmem_o + d * 64, umma_desc_sw128(smem_u32(s_p) + poff), vdesc, idesc_o, (j | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&r_empty[slot]);
          if (++slot == AW_RING) { slot = 0; ph ^= 1; }
        }
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
      const uint32_t tmem_s = tmem_base + (j & 1) * 128;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const int valid_keys = min(AT_BN, p.tokens_kv - j * AT_BN);
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < valid_keys) m_tile = fmaxf(m_tile, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, m_tile);
      const float alpha = m_new == -INFINITY ? 1.0f : exp2f((m_run - m_new) * p.scale_log2e);
      // previous P V must be complete before P is overwritten and O is rescaled
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {      // the running maximum of some row of this warp moved
#pragma unroll 1
          for (int c = 0; c < AW_OD; c += 32) {
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
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_s + lane_addr + c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c + i < valid_keys) ? exp2f(__uint_as_float(r[i]) * p.scale_log2e - mb) : 0.f;
          float p1 = (c + i + 1 < valid_keys) ? exp2f(__uint_as_float(r[i + 1]) * p.scale_log2e - mb) : 0.f;
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
          l_tile += __low2float(h2) + __high2float(h2);   // the sum of what the P V product actually consumes
          pk[i / 2] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        if constexpr (AW_P_TMEM) {
          // columns [c / 2, c / 2 + 16) of this S buffer: all of them were read by this or an earlier chunk
          tmem_st16(tmem_s + lane_addr + c / 2, pk);
        } else {
          uint8_t* slab = s_p + (c / 64) * AW_SLAB + (row / 8) * 1024 + (row % 8) * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = ((c % 64) / 8 + g) ^ (row % 8);
            *reinterpret_cast<uint4*>(slab + chunk * 16) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          }
        }
      }
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      if constexpr (AW_P_TMEM) tmem_st_wait();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const int t = q0 + row;
    const float inv_l = 1.0f / l_run;
    __nv_bfloat16* dst = p.out + static_cast<size_t>(b) * p.out_stride_b + static_cast<size_t>(t) * AW_D + half * AW_OD;
#pragma unroll 1
    for (int c = 0; c < AW_OD; c += 32) {
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
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_qkv_map(CUtensorMap* m, const void* base, int batch, int heads, int tokens, int head_dim, long long sb,
                 long long sh, long long st) {
  // dims innermost-first: channel, token, head, batch
  const uint64_t dims[4] = {static_cast<uint64_t>(head_dim), static_cast<uint64_t>(tokens),
                            static_cast<uint64_t>(heads), static_cast<uint64_t>(batch)};
  const uint64_t head_stride = heads > 1 ? static_cast<uint64_t>(sh) * 2 : static_cast<uint64_t>(st) * 2 * tokens;
  const uint64_t batch_stride = batch > 1 ? static_cast<uint64_t>(sb) * 2 : static_cast<uint64_t>(st) * 2 * tokens;
  const uint64_t str[3] = {static_cast<uint64_t>(st) * 2, head_stride, batch_stride};
  const uint32_t box[4] = {64, AT_BM, 1, 1};
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
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaError_t e = cudaFuncSetAttribute(attention_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AW_SMEM_BYTES);
      if (e != cudaSuccess) {
        set_error("attention_tc: cudaFuncSetAttribute(%d B smem): %s", AW_SMEM_BYTES, cudaGetErrorString(e));
        return ERR_CUDA;
      }
      configured.done();
    }
    dim3 grid((tokens + AT_BM - 1) / AT_BM, AW_D / AW_OD, batch);
    attention_wide_kernel<<<grid, AT_THREADS, AW_SMEM_BYTES, s>>>(mq, mk, mv, p);
    return check_launch("attention_tc (wide)");
  }
  return head_dim == 128 ? launch_attn<128>(mq, mk, mv, p, batch, s) : launch_attn<64>(mq, mk, mv, p, batch, s);
}
