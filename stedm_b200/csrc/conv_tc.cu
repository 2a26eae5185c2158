// K1/K3/K4/K10: convolution (3x3 pad 1 or 1x1, stride 1) and plain GEMM as an implicit GEMM on Blackwell's
// 5th-generation tensor cores: tcgen05.mma (bf16 x bf16 -> fp32) issued by one thread, operands staged in
// shared memory by TMA with the 128-byte swizzle, accumulators in TMEM, epilogue through tcgen05.ld.
//
//   M = B*H*W output pixels (tile 128 = UMMA_M), N = Cout (tile BN = UMMA_N), K = taps * Cin (slab 64).
//
// A operand (im2col of the NHWC activation) is never materialised: for filter tap (r, s) the producer issues a
// 4-D TMA load of the box {64 channels, tile_w, tile_h, tile_b} at pixel offset (s-1, r-1); out-of-bounds rows
// and columns are zero-filled by the TMA unit, which is exactly the convolution's zero padding.  The box lands
// in shared memory as 128 rows x 128 B, i.e. the canonical K-major SWIZZLE_128B UMMA layout.  A channel concat
// [x0 | x1] (openaimodel.py:800) is a second tensor map selected per K slab — no concatenated copy exists.
// B operand = weights repacked once at load time to bf16 [Cout][tap][Cin] (K-major), 2-D TMA box {64, BN/CL}.
//
// PERSISTENT kernel: one CTA (or CTA pair) per SM loops over output tiles (output-channel tile fastest, so the
// n-tiles of a pixel tile run concurrently and share activation slabs through L2).  Warp roles (192 threads):
// warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2-5 =
// epilogue (TMEM lane quadrant = warp % 4).  Three pipelines: the smem ring of STAGES {A,B} slabs (full/empty
// mbarriers, released by tcgen05.commit) runs continuously across tiles; the accumulator is DOUBLE-BUFFERED in
// TMEM (tmem_full/tmem_empty), so the epilogue of tile i overlaps the main loop of tile i+1 — measured necessary:
// with a single accumulator the tensor pipe idled ~30 % of the time behind the epilogue's global loads/stores.
// CL = 2: the two CTAs of a cluster take neighbouring pixel tiles of the same channel tile and TMA-multicast one
// half of the weight slab each into both CTAs (per-SM L2->SM traffic A+B/2 instead of A+B per slab).
// HALO mode (k x k taps, tiles of >= 2 whole image rows of one sample): the activation box of the tile's rows plus the
// n_t - 1 halo rows is loaded ONCE per (64-channel block, horizontal tap) and serves the n_t vertical taps — the A
// descriptor's start address advances by one image row (W * 128 B = whole swizzle atoms) per vertical tap — so an
// activation byte crosses L2->SM 3 (th + 2) / th times instead of 9.  Boxes and weight slabs then live in two rings
// with their own barriers (a box outlives n_t weight slabs); otherwise both rings run in lock step on one barrier pair.
// The producer is ONE thread and its instruction count per slab was the measured limiter: no divisions in its slab loop.
// K loop extras: the 1x1 skip_connection of a channel-changing ResBlock (openaimodel.py:246-256, 288) is folded in as
// extra K slabs read from the block input through two more tensor maps, so conv3x3(a) + skip1x1(x) is one accumulator.
// XF mode (GroupNorm + SiLU in the operand path, util.py:199-216 / openaimodel.py:268-288): the convolution reads the RAW
// producer output; four extra "transform" warps turn each raw halo box (ONE TMA load per 64-channel block) into the three
// horizontally shifted operand boxes y = silu(x * a[b][c] + s[b][c]) (the per-(sample, channel) coefficients come from
// stedm_gn_fold_tiles), re-zeroing the padding rows / columns, between TMA arrival and the MMA — the normalised tensor
// is never written to HBM and an activation byte crosses L2->SM once per channel tile instead of three times.
// Epilogue: + bias[n] + emb[b][n] (timestep / style embedding, openaimodel.py:278-287), optional exact-erf GELU,
// + residual[m][n] (identity skip, openaimodel.py:288); the bias / embedding rows are fetched while tcgen05.ld is in
// flight, and NHWC outputs pass through a per-warp XOR-swizzled shared-memory transpose so that 4 lanes write one pixel
// row's 64 contiguous bytes (whole 32-byte sectors; a thread owns a ROW of the accumulator, so direct stores were 32
// half-written sectors per instruction); channel-major output for the eps / image heads; per-tile GroupNorm statistics.
#include "../../include/stedm_b200.h"
#include <stdlib.h>

#include "common.cuh"

using namespace stedm;

namespace {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr int TC_THREADS_XF = 352;  // + 4 transform warps (warps 6-9) + the activation producer (warp 10)
// XF mode shared-memory operand pool: 2 sets x 3 boxes ([dx=-1 | raw -> dx=0 | dx=+1]) + a weight ring in what is left
constexpr int TC_XF_POOL = 209 * 1024;
constexpr int TC_XF_SETS = 2;
constexpr int TC_XF_SLOTS = 3 * TC_XF_SETS;
constexpr int TC_XF_MAX_NB = 8;
constexpr int TC_XF_BAR_BYTES = 512;

// STEDM_TC_CLUSTER=0 disables the 2-CTA weight multicast (A/B measurements, debugging)
static bool g_tc_cluster_enabled = [] {
  const char* e = getenv("STEDM_TC_CLUSTER");
  return !(e && e[0] == '0');
}();

// STEDM_TC_SPLITK=0 disables split-K for launches with few output tiles
static bool g_tc_splitk_enabled = [] {
  const char* e = getenv("STEDM_TC_SPLITK");
  return !(e && e[0] == '0');
}();

// STEDM_TC_PAIR=0 falls back from the cta_group::2 (CTA-pair MMA) variant to cta_group::1 + weight multicast
static bool g_tc_pair_enabled = [] {
  const char* e = getenv("STEDM_TC_PAIR");
  return !(e && e[0] == '0');
}();

// STEDM_TC_HALO=0 loads one activation slab per filter tap instead of one halo box per (channel block, horizontal tap)
static bool g_tc_halo_enabled = [] {
  const char* e = getenv("STEDM_TC_HALO");
  return !(e && e[0] == '0');
}();

struct TcParams {
  const float* bias;
  const float* emb;
  const void* residual;
  void* out;
  int M, H, W, HW, cout;
  int taps, ksize;
  int c0_blks, c_blks;  // 64-channel slabs in source 0 / in the concat
  int x1_batch;
  int n_tiles;          // output-channel tiles
  int num_work;         // cluster work items = ceil(m_tiles / CL) * n_tiles * ksplit
  int ksplit;           // split-K factor (1 = off): work item = (tile, K range); partials go to `ws`
  int kb_per_split;     // K slabs per split
  float* ws;            // split-K workspace: fp32 [ksplit][m_pad][cout]
  int m_pad;            // rows of one workspace slice (all tiles, including the out-of-bounds one)
  int emb_stride, res_dtype, out_dtype;
  int out_nchw, cout_store;
  // 2-D pixel tiles (3x3 convolutions on maps at least 32 wide): a tile is 16 columns x 8 rows of one sample instead of
  // 128 consecutive pixels, so that the halo box (10 x 16 pixels, 20 KB) carries 25 % extra rows whatever the map width —
  // full-row tiles need 2 x (W = 64) or cannot use halo boxes at all (W >= 128)
  int tile2d, txs, tps;  // flag, tiles per tile row (W / 16), tiles per sample (H * W / 128)
  int cs;                // input pixels per output pixel (2 for the stride-2 Downsample convolution: strided TMA boxes)
  int tap_mode, py, px;  // tap_mode 1: 2x2 sub-pixel phase (py, px) of nearest-x2-upsample + 3x3 conv
  int act;               // STEDM_ACT_*: applied to acc + bias + emb, before the residual
  // fused 1x1 skip convolution (ResBlock skip_connection / nin_shortcut): extra K slabs after the k x k taps, read at
  // the output pixel itself from a second input [skip0 | skip1]; its weights are appended along K
  int skip_c0_blks, skip_blks, skip_x1_batch;
  int res_rows;          // > 0: the residual has fewer samples than the output and is broadcast: row = m % res_rows
  // halo mode (k x k taps over tiles made of whole image rows of one sample): ONE TMA load of the tile's rows plus the
  // n_t - 1 halo rows per (64-channel block, horizontal tap) serves the n_t vertical taps of that column — the MMA's A
  // descriptor start advances by one image row (W * 128 B, a whole number of swizzle atoms) per vertical tap
  int halo;              // 0: one A slab per tap
  int n_t;               // taps per dimension in halo mode: 3, or 2 for the sub-pixel phases
  int na;                // A ring buffers (a_buf_bytes each, carved out of the STAGES * 16 KB pool)
  int a_buf_bytes;       // 16 KB, or the halo box (th + n_t - 1) * W * 128 B
  int a_row_bytes;       // W * 128
  // XF mode: GroupNorm (+ SiLU) applied to the raw input inside the kernel
  const float* gn_coef;  // fp32 [samples][gn_cstride][2] = (scale, shift) per (sample, concat channel); nullptr = off
  int gn_cstride, gn_c_off, gn_silu;
  int nb;                // weight ring depth (XF: what fits beside the 6 boxes; otherwise Cfg::STAGES)
  int a_pool_bytes;      // XF: 6 * a_buf_bytes
  int log2w;             // XF: W is a power of two
  float* stats_out;      // optional [tile entries][cout][2]: per-(pixel tile, channel) sum / sum of squares of the output
  int stats_tile_base;   // first tile entry of this launch (phase * m_tiles for the sub-pixel phases)
};

// Origin (sample, row, column) of pixel tile `t` and the linear NHWC pixel index of its row `r`.
__device__ __forceinline__ void tc_tile_origin(const TcParams& p, int t, int& b0, int& y0, int& x0) {
  if (p.tile2d) {
    b0 = t / p.tps;
    const int r = t - b0 * p.tps, ty = r / p.txs;
    y0 = ty * 8;
    x0 = (r - ty * p.txs) * 16;
  } else {
    const int m0 = t * TC_BM;
    x0 = m0 % p.W;
    y0 = (m0 / p.W) % p.H;
    b0 = m0 / p.HW;
  }
}
// row r of tile t -> linear pixel index (b * H + y) * W + x, or -1 past the end of the tensor
__device__ __forceinline__ int tc_tile_pixel(const TcParams& p, int t, int r) {
  if (p.tile2d) {
    if (t * TC_BM >= p.M) return -1;     // the all-out-of-bounds tile of an odd tile count
    int b0, y0, x0;
    tc_tile_origin(p, t, b0, y0, x0);
    return (b0 * p.H + y0 + (r >> 4)) * p.W + x0 + (r & 15);
  }
  const int m = t * TC_BM + r;
  return m < p.M ? m : -1;
}

// PAIR: cta_group::2 — the two CTAs of a cluster form one 256 x BN MMA tile (128 pixel rows each); each CTA stages
// only its HALF of the weight slab and the tensor core reads both halves, so the bytes that must enter an SM per
// MMA drop from A + B to A + B/2 (the measured limiter of cta_group::1 at BN = 256: ~75 B/clk/SM of SM ingest
// against 96 B/clk/SM needed -> tensor pipe 79 % active).
constexpr int tc_stages(int bn, bool pair) { return pair ? (bn >= 256 ? 6 : 4) : (bn >= 256 ? 4 : (bn >= 128 ? 3 : 4)); }

template <int BN, bool PAIR = false>
struct TcCfg {
  static constexpr int STAGES = tc_stages(BN, PAIR);
  static constexpr int ACC = 2;  // TMEM accumulator stages
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * TC_BK * 2;  // bytes of the weight slab staged in THIS CTA
  static constexpr int B_BYTES_PAD = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES_PAD;
  static constexpr int A_POOL = STAGES * A_BYTES;  // A ring: STAGES slabs of 16 KB, or fewer and larger halo boxes
  static constexpr int ACC_COLS = BN;  // fp32 columns per accumulator
  static constexpr int TMEM_COLS = (ACC * BN) < 32 ? 32 : ACC * BN;  // 32 / 128 / 256 / 512: powers of two
  static constexpr int STATS_BYTES = 4 * BN * 2 * 4;  // [4 epilogue warps][BN][sum, sumsq] fp32
  static constexpr int OUT_STAGE_BYTES = 4 * 32 * 64;  // [4 epilogue warps][32 rows][64 B]: output transpose staging
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STATS_BYTES + OUT_STAGE_BYTES;
  static constexpr int MIN_BLOCKS = (BN >= 256) ? 1 : 2;  // TMEM: 1 x 512 or 2 x <=256 columns per SM
  static constexpr int SMEM_BYTES_XF = TC_XF_POOL + 1024 + TC_XF_BAR_BYTES + STATS_BYTES + OUT_STAGE_BYTES;
};

template <int BN, int CL, bool PAIR, bool HALO, bool XF>
__global__ void __launch_bounds__(XF ? TC_THREADS_XF : TC_THREADS, TcCfg<BN, PAIR>::MIN_BLOCKS)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_s0,
               const __grid_constant__ CUtensorMap map_s1, const TcParams p) {
  static_assert(!PAIR || CL == 2, "cta_group::2 needs a 2-CTA cluster");
  static_assert(!XF || (PAIR && HALO), "the in-kernel GroupNorm transform is built on the CTA-pair halo pipeline");
  using Cfg = TcCfg<BN, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Operand area: [activation pool: STAGES slabs of 16 KB, or fewer and larger halo boxes | STAGES weight slabs]
  // (measured 2-3 % faster than [activation | weight] back to back per slot).  XF: [2 sets x 3 boxes | p.nb weight slabs].
  constexpr int POOL_BYTES = XF ? TC_XF_POOL : Cfg::STAGES * Cfg::STAGE_BYTES;
  constexpr int BAR_BYTES = XF ? TC_XF_BAR_BYTES : 256;
  constexpr int NF_BARS = XF ? TC_XF_SETS : Cfg::STAGES;     // "activation landed" barriers (XF: one per SET of three boxes)
  constexpr int NE_BARS = XF ? TC_XF_SLOTS : Cfg::STAGES;    // "activation buffer free" barriers (XF: one per box slot)
  constexpr int NB_BARS = XF ? TC_XF_MAX_NB : Cfg::STAGES;   // weight ring barriers
  const uint32_t b_base = XF ? static_cast<uint32_t>(p.a_pool_bytes) : static_cast<uint32_t>(Cfg::A_POOL);
  const uint32_t nb = XF ? static_cast<uint32_t>(p.nb) : static_cast<uint32_t>(Cfg::STAGES);
  auto slab_a = [](uint8_t* base, uint32_t s) { return base + s * Cfg::A_BYTES; };
  auto slab_b = [b_base](uint8_t* base, uint32_t s, int) { return base + b_base + s * Cfg::B_BYTES_PAD; };
  // two rings: activation buffers (A) and weight slabs (B), each with full / empty mbarriers
  uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + POOL_BYTES);
  uint64_t* empty_a = full_a + NF_BARS;
  uint64_t* full_b = empty_a + NE_BARS;
  uint64_t* empty_b = full_b + NB_BARS;
  uint64_t* tmem_full_bar = empty_b + NB_BARS;         // [ACC]
  uint64_t* tmem_empty_bar = tmem_full_bar + Cfg::ACC; // [ACC]
  uint64_t* xf_ready = tmem_empty_bar + Cfg::ACC;      // [TC_XF_SETS] (XF only): the set's three boxes are transformed
  uint64_t* full_s = xf_ready + TC_XF_SETS;            // [TC_XF_SLOTS] (XF only): a raw fused-skip slab has landed in the slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(XF ? full_s + TC_XF_SLOTS : xf_ready);
  static_assert((NF_BARS + NE_BARS + 2 * NB_BARS + 2 * Cfg::ACC + (XF ? TC_XF_SETS + TC_XF_SLOTS : 0)) * 8 + 4 <= BAR_BYTES,
                "barrier area");
  float* s_stats = reinterpret_cast<float*>(smem + POOL_BYTES + BAR_BYTES);
  uint8_t* s_out = smem + POOL_BYTES + BAR_BYTES + Cfg::STATS_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CL > 1) ? cluster_ctarank() : 0u;
  const int cluster_id = blockIdx.x / CL, num_clusters = gridDim.x / CL;
  const int main_kb = p.taps * p.c_blks;
  const int num_kb = main_kb + p.skip_blks;
  const int halo_items = p.halo ? p.c_blks * p.n_t : 0;  // (64-channel block, horizontal tap) items served by halo loads

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_w);
    if (p.skip_blks > 0) {
      tma_prefetch_desc(&map_s0);
      tma_prefetch_desc(&map_s1);
    }
    // PAIR: the leader's full barriers collect both CTAs' producers; its single commit frees the slab in both.
    // XF: the raw box is consumed by this CTA's own transform warps -> a local barrier with one producer
    for (int i = 0; i < NF_BARS; ++i) mbar_init(&full_a[i], (PAIR && !XF) ? 2 : 1);
    for (int i = 0; i < NE_BARS; ++i) mbar_init(&empty_a[i], 1);  // activations are never multicast
    for (int i = 0; i < NB_BARS; ++i) {
      mbar_init(&full_b[i], PAIR ? 2 : 1);
      mbar_init(&empty_b[i], PAIR ? 1 : CL);    // multicast: released by the MMA commit of every CTA writing into it
    }
    if constexpr (XF) {
      for (int i = 0; i < TC_XF_SETS; ++i) mbar_init(&xf_ready[i], 8);  // 4 transform warps of each CTA of the pair
      for (int i = 0; i < TC_XF_SLOTS; ++i) mbar_init(&full_s[i], 2);   // both CTAs' activation producers
    }
    for (int i = 0; i < Cfg::ACC; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], PAIR ? 256 : 128);  // every epilogue thread (of both CTAs) arrives
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // peers' barriers must exist before any multicast lands
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // One thread; its per-slab instruction count matters (measured: ~80 % busy at 512 MMA clocks per slab), so the
    // tap / channel-block indices are carried as counters (no divisions in the slab loop) and HALO is compile time.
    if (elect_one()) {
      uint32_t ia = 0, pa = 0, ib = 0, pb = 0;  // slot / phase of the A and B rings, running across tiles
      const int n_t = p.n_t;                    // taps per dimension: 3 (3x3), 2 (sub-pixel phase), 1 (1x1 / GEMM)
      const int dyf = p.tap_mode == 1 ? p.py - 1 : (p.ksize == 3 ? -1 : 0);  // first tap offsets
      const int dxf = p.tap_mode == 1 ? p.px - 1 : (p.ksize == 3 ? -1 : 0);
      const int c_blks = p.c_blks, c0_blks = p.c0_blks;
      // the leader's full barriers (PAIR: both CTAs' loads complete on CTA 0's barrier)
      auto arrive_full = [&](uint64_t* bar, uint32_t bytes) {
        if constexpr (PAIR) mbar_arrive_expect_tx_cluster(bar, bytes, 0);
        else mbar_arrive_expect_tx(bar, bytes);
      };
      auto load_a = [&](uint8_t* dst, const CUtensorMap* m, uint64_t* bar, int ch, int cx, int cy, int bb) {
        if constexpr (PAIR) tma_load_4d_2sm(dst, m, bar, ch, cx, cy, bb);
        else tma_load_4d(dst, m, bar, ch, cx, cy, bb);
      };
      auto load_b = [&](uint32_t slot, int kb, int n0) {
        uint8_t* sb = slab_b(smem, slot, 0);
        if constexpr (PAIR) {  // this CTA's half of the weight slab: rows [rank*BN/2, +BN/2)
          tma_load_2d_2sm(sb, &map_w, &full_b[slot], kb * TC_BK, n0 + static_cast<int>(cta_rank) * (BN / 2));
        } else if (CL == 1) {
          tma_load_2d(sb, &map_w, &full_b[slot], kb * TC_BK, n0);
        } else {  // this CTA fetches rows [rank*BN/CL, +BN/CL) of the weight slab for the whole cluster
          constexpr int ROWS = BN / CL;
          tma_load_2d_mcast(sb + cta_rank * (ROWS * TC_BK * 2), &map_w, &full_b[slot], kb * TC_BK,
                            n0 + static_cast<int>(cta_rank) * ROWS, static_cast<uint16_t>((1u << CL) - 1));
        }
      };
      for (int work = cluster_id; work < p.num_work; work += num_clusters) {
        const int tile_id = work / p.ksplit, split = work - tile_id * p.ksplit;
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        const int n0 = (tile_id % p.n_tiles) * BN;
        int x0, y0, b0;
        tc_tile_origin(p, (tile_id / p.n_tiles) * CL + static_cast<int>(cta_rank), b0, y0, x0);
        const int b1 = (p.x1_batch > 0) ? (b0 % p.x1_batch) : b0;
        const int bs1 = (p.skip_x1_batch > 0) ? (b0 % p.skip_x1_batch) : b0;
        if constexpr (XF) {
          // weights only (the raw activation boxes have their own producer, warp 10, so that their prefetch distance
          // is not tied to the depth of the weight ring): nine slabs per 64-channel block in the order the MMA consumes
          // them — horizontal tap bx, vertical tap g — then the fused skip slabs
          for (int cb = 0; cb < c_blks; ++cb) {
            for (int bx = 0; bx < 3; ++bx) {
              int kb = bx * c_blks + cb;
              for (int g = 0; g < 3; ++g, kb += 3 * c_blks) {
                mbar_wait(&empty_b[ib], pb ^ 1);
                arrive_full(&full_b[ib], Cfg::B_BYTES);
                load_b(ib, kb, n0);
                if (++ib == nb) { ib = 0; pb ^= 1; }
              }
            }
          }
          for (int sb = 0; sb < p.skip_blks; ++sb) {
            mbar_wait(&empty_b[ib], pb ^ 1);
            arrive_full(&full_b[ib], Cfg::B_BYTES);
            load_b(ib, main_kb + sb, n0);
            if (++ib == nb) { ib = 0; pb ^= 1; }
          }
        } else if constexpr (HALO) {
          // items (64-channel block cb, horizontal tap bx): one activation box of the tile's rows + halo, then the
          // n_t weight slabs of the vertical taps a: K slab (a * n_t + bx) * c_blks + cb
          for (int cb = 0; cb < c_blks; ++cb) {
            const bool first = cb < c0_blks;
            const CUtensorMap* amap = first ? &map_a0 : &map_a1;
            const int ach = (first ? cb : cb - c0_blks) * TC_BK, ab = first ? b0 : b1;
            for (int bx = 0; bx < n_t; ++bx) {
              mbar_wait(&empty_a[ia], pa ^ 1);
              arrive_full(&full_a[ia], static_cast<uint32_t>(p.a_buf_bytes));
              load_a(smem + ia * p.a_buf_bytes, amap, &full_a[ia], ach, x0 + dxf + bx, y0 + dyf, ab);
              if (++ia == static_cast<uint32_t>(p.na)) { ia = 0; pa ^= 1; }
              int kb = bx * c_blks + cb;
              for (int g = 0; g < n_t; ++g, kb += n_t * c_blks) {
                mbar_wait(&empty_b[ib], pb ^ 1);
                arrive_full(&full_b[ib], Cfg::B_BYTES);
                load_b(ib, kb, n0);
                if (++ib == static_cast<uint32_t>(Cfg::STAGES)) { ib = 0; pb ^= 1; }
              }
            }
          }
          // fused 1x1 skip input: one plain slab of the tile's own pixels per 64 channels
          for (int sb = 0; sb < p.skip_blks; ++sb) {
            const bool first = sb < p.skip_c0_blks;
            mbar_wait(&empty_a[ia], pa ^ 1);
            arrive_full(&full_a[ia], Cfg::A_BYTES);
            load_a(smem + ia * p.a_buf_bytes, first ? &map_s0 : &map_s1, &full_a[ia],
                   (first ? sb : sb - p.skip_c0_blks) * TC_BK, x0, y0, first ? b0 : bs1);
            if (++ia == static_cast<uint32_t>(p.na)) { ia = 0; pa ^= 1; }
            mbar_wait(&empty_b[ib], pb ^ 1);
            arrive_full(&full_b[ib], Cfg::B_BYTES);
            load_b(ib, main_kb + sb, n0);
            if (++ib == static_cast<uint32_t>(Cfg::STAGES)) { ib = 0; pb ^= 1; }
          }
        } else {
          // one activation slab + one weight slab per K slab, sharing slot and barriers; K slab kb = tap * c_blks + cb
          int tap = 0, cb = kb0;
          if (kb0 >= c_blks) { tap = kb0 / c_blks; cb = kb0 - tap * c_blks; }      // split-K ranges only
          int ta = tap / n_t, tb = tap - ta * n_t;                                // tap (ta, tb): dy = dyf + ta, dx = dxf + tb
          for (int kb = kb0; kb < kb1; ++kb) {
            const CUtensorMap* amap;
            int ach, ab, cx = p.cs * x0, cy = p.cs * y0;
            if (kb < main_kb) {
              const bool first = cb < c0_blks;
              amap = first ? &map_a0 : &map_a1;
              ach = (first ? cb : cb - c0_blks) * TC_BK;
              ab = first ? b0 : b1;
              cx += dxf + tb;
              cy += dyf + ta;
              if (++cb == c_blks) {
                cb = 0;
                if (++tb == n_t) { tb = 0; ++ta; }
              }
            } else {
              const int sb = kb - main_kb;
              const bool first = sb < p.skip_c0_blks;
              amap = first ? &map_s0 : &map_s1;
              ach = (first ? sb : sb - p.skip_c0_blks) * TC_BK;
              ab = first ? b0 : bs1;
            }
            mbar_wait(&empty_b[ib], pb ^ 1);
            arrive_full(&full_b[ib], Cfg::A_BYTES + Cfg::B_BYTES);
            load_a(slab_a(smem, ib), amap, &full_b[ib], ach, cx, cy, ab);
            load_b(ib, kb, n0);
            if (++ib == static_cast<uint32_t>(Cfg::STAGES)) { ib = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if ((!PAIR || cta_rank == 0) && elect_one()) {  // PAIR: only the leader CTA issues (for both)
      constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * TC_BM : TC_BM, BN);
      uint32_t ia = 0, pa = 0, ib = 0, pb = 0, tile = 0, xf_par = 0, sk_par = 0;
      // the 4 MMAs (UMMA_K = 16 bf16 = 32 B -> start address field += 2) of one K slab, then the commit that frees
      // the weight slot (in every CTA that multicasts into it / of the pair) once they have read it
      auto mma_slab = [&](uint32_t tmem_d, uint32_t a_addr, uint32_t slot, uint32_t started) {
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(smem_u32(slab_b(smem, slot, 0)));
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          const uint32_t accumulate = (started | k) != 0 ? 1u : 0u;
          if constexpr (PAIR) umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
          else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
        }
        if constexpr (PAIR) umma_commit_2sm_mcast(&empty_b[slot], 3);
        else if (CL == 1) umma_commit(&empty_b[slot]);
        else umma_commit_mcast(&empty_b[slot], static_cast<uint16_t>((1u << CL) - 1));
      };
      auto release_a = [&](uint32_t slot) {  // the activation box after its last tap (in both CTAs of a pair)
        if constexpr (PAIR) umma_commit_2sm_mcast(&empty_a[slot], 3);
        else umma_commit(&empty_a[slot]);
      };
      for (int work = cluster_id; work < p.num_work; work += num_clusters, ++tile) {
        const uint32_t acc = tile % Cfg::ACC, acc_ph = (tile / Cfg::ACC) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * Cfg::ACC_COLS;
        uint32_t started = 0;
        if constexpr (XF) {
          // ia = set (0 / 1), pa = the phase of its slot-free barriers (one completion per use of the set); bit `ia` of
          // xf_par / sk_par = the phase of the set's "transformed" / "skip slab landed" barriers, which complete only on
          // the uses of that kind.  A box slot is handed back as soon as its three vertical taps are issued.
          for (int cb = 0; cb < p.c_blks; ++cb) {
            mbar_wait_cluster(&xf_ready[ia], (xf_par >> ia) & 1u);  // both CTAs' transform warps have written the three boxes
            xf_par ^= 1u << ia;
            tc_fence_after();
            for (int bx = 0; bx < 3; ++bx) {
              const uint32_t a_addr = smem_u32(smem + (ia * 3 + bx) * p.a_buf_bytes);
              for (int g = 0; g < 3; ++g) {
                mbar_wait(&full_b[ib], pb);
                tc_fence_after();
                mma_slab(tmem_d, a_addr + g * p.a_row_bytes, ib, started);
                started = 1;
                if (++ib == nb) { ib = 0; pb ^= 1; }
              }
              release_a(ia * 3 + bx);
            }
            if (++ia == TC_XF_SETS) { ia = 0; pa ^= 1; }
          }
          // fused skip slabs: raw, one per box slot, each with its own "landed" barrier (six in flight)
          for (int sb = 0; sb < p.skip_blks; sb += 3) {
            const int n = min(3, p.skip_blks - sb);
            for (int i = 0; i < 3; ++i) {
              if (i < n) {
                mbar_wait(&full_s[ia * 3 + i], (sk_par >> ia) & 1u);
                mbar_wait(&full_b[ib], pb);
                tc_fence_after();
                mma_slab(tmem_d, smem_u32(smem + (ia * 3 + i) * p.a_buf_bytes), ib, started);
                started = 1;
                if (++ib == nb) { ib = 0; pb ^= 1; }
              }
              release_a(ia * 3 + i);   // every slot barrier completes once per use of the set, used or not
            }
            sk_par ^= 1u << ia;
            if (++ia == TC_XF_SETS) { ia = 0; pa ^= 1; }
          }
        } else if constexpr (HALO) {
          const int n_t = p.n_t;
          for (int i = 0; i < halo_items; ++i) {
            mbar_wait(&full_a[ia], pa);
            const uint32_t a_addr = smem_u32(smem + ia * p.a_buf_bytes);
            for (int g = 0; g < n_t; ++g) {  // vertical tap g reads the same box g image rows further down
              mbar_wait(&full_b[ib], pb);
              tc_fence_after();
              mma_slab(tmem_d, a_addr + g * p.a_row_bytes, ib, started);
              started = 1;
              if (++ib == static_cast<uint32_t>(Cfg::STAGES)) { ib = 0; pb ^= 1; }
            }
            release_a(ia);
            if (++ia == static_cast<uint32_t>(p.na)) { ia = 0; pa ^= 1; }
          }
          for (int sb = 0; sb < p.skip_blks; ++sb) {
            mbar_wait(&full_a[ia], pa);
            mbar_wait(&full_b[ib], pb);
            tc_fence_after();
            mma_slab(tmem_d, smem_u32(smem + ia * p.a_buf_bytes), ib, started);
            started = 1;
            if (++ib == static_cast<uint32_t>(Cfg::STAGES)) { ib = 0; pb ^= 1; }
            release_a(ia);
            if (++ia == static_cast<uint32_t>(p.na)) { ia = 0; pa ^= 1; }
          }
        } else {
          const int split = work % p.ksplit;
          const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_b[ib], pb);
            tc_fence_after();
            mma_slab(tmem_d, smem_u32(slab_a(smem, ib)), ib, started);
            started = 1;
            if (++ib == static_cast<uint32_t>(Cfg::STAGES)) { ib = 0; pb ^= 1; }
          }
        }
        // accumulator complete -> epilogue (of both CTAs in PAIR mode)
        if constexpr (PAIR) umma_commit_2sm_mcast(&tmem_full_bar[acc], 3);
        else umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (XF && warp == 10) {
    // ============================ raw activation producer (XF only) =================================
    // ONE raw box (the tile's rows + 2 halo rows, horizontal offset 0) per 64-channel block into the MIDDLE slot of a
    // set, as soon as the MMA has released that slot (after the previous use's second horizontal tap); the fused skip
    // input as raw 16 KB slabs of the tile's own pixels, up to three per set.
    if constexpr (XF) {
      if (elect_one()) {
        uint32_t is = 0, ps = 0;
        for (int work = cluster_id; work < p.num_work; work += num_clusters) {
          const int tile_id = work / p.ksplit;
          const int m0 = ((tile_id / p.n_tiles) * CL + static_cast<int>(cta_rank)) * TC_BM;
          const int x0 = m0 % p.W, y0 = (m0 / p.W) % p.H, b0 = m0 / p.HW;
          const int b1 = (p.x1_batch > 0) ? (b0 % p.x1_batch) : b0;
          const int bs1 = (p.skip_x1_batch > 0) ? (b0 % p.skip_x1_batch) : b0;
          for (int cb = 0; cb < p.c_blks; ++cb) {
            const bool first = cb < p.c0_blks;
            mbar_wait(&empty_a[is * 3 + 1], ps ^ 1);
            mbar_arrive_expect_tx(&full_a[is], static_cast<uint32_t>(p.a_buf_bytes));
            tma_load_4d(smem + (is * 3 + 1) * p.a_buf_bytes, first ? &map_a0 : &map_a1, &full_a[is],
                        (first ? cb : cb - p.c0_blks) * TC_BK, x0, y0 - 1, first ? b0 : b1);
            if (++is == TC_XF_SETS) { is = 0; ps ^= 1; }
          }
          for (int sb = 0; sb < p.skip_blks; sb += 3) {
            const int n = min(3, p.skip_blks - sb);
            for (int i = 0; i < 3; ++i) {
              const uint32_t slot = is * 3 + i;
              mbar_wait(&empty_a[slot], ps ^ 1);   // also: the slot's previous "landed" phase is complete and consumed
              if (i < n) {                         // straight to the leader's MMA thread: no transform on this path
                const bool first = sb + i < p.skip_c0_blks;
                mbar_arrive_expect_tx_cluster(&full_s[slot], Cfg::A_BYTES, 0);
                tma_load_4d_2sm(smem + slot * p.a_buf_bytes, first ? &map_s0 : &map_s1, &full_s[slot],
                                (first ? sb + i : sb + i - p.skip_c0_blks) * TC_BK, x0, y0, first ? b0 : bs1);
              } else {
                mbar_arrive_cluster(&full_s[slot], 0);   // unused slot of the last group: keep its phase in step
              }
            }
            if (++is == TC_XF_SETS) { is = 0; ps ^= 1; }
          }
        }
      }
    }
  } else if (XF && warp >= 6) {
    // ============================ GroupNorm + SiLU transform (XF only) ==============================
    // 128 threads: thread -> (16-byte chunk j = 8 channels, pixel lane); per set: read the raw box (middle slot, as
    // TMA wrote it: pixel row p at p * 128 B, chunk j at slot j ^ (p & 7)), y = silu(x * a + s) in fp32, bf16 again, and
    // write the three operand boxes: dx = 0 in place, dx = -1 / +1 one pixel to the right / left in the two outer
    // slots, with the column that has no source pixel and the rows outside the image written as ZERO (the
    // convolution's padding — TMA's zero fill would turn into silu(s) otherwise).  Same formula and rounding as
    // gn_apply_kernel, so the result is bit-identical to the unfused path.
    if constexpr (XF) {
      const int tt = static_cast<int>(threadIdx.x) - 192;
      const int j = tt & 7, pl = tt >> 3;
      const int wmask = p.W - 1, log2w = p.log2w, H = p.H, n_it = (p.a_buf_bytes >> 7) / 16;  // box pixels % 16 == 0
      const bool silu = p.gn_silu != 0;
      const uint32_t smem_base = smem_u32(smem), box = static_cast<uint32_t>(p.a_buf_bytes);
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      uint32_t is = 0, ps = 0, xf_par = 0;
      // (scale, shift) rows of this thread's 8 channels for the tile of work item `w` (the all-out-of-bounds tile of an
      // odd tile count reads sample 0's: its results are never stored)
      auto coef_row = [&](int w) {
        const int m0 = (((w / p.ksplit) / p.n_tiles) * CL + static_cast<int>(cta_rank)) * TC_BM;
        const int b0 = m0 < p.M ? m0 / p.HW : 0;
        return reinterpret_cast<const float4*>(p.gn_coef + (static_cast<size_t>(b0) * p.gn_cstride + p.gn_c_off + j * 8) * 2);
      };
      // the next block's coefficients — of this tile or of the first block of the next one — are always in flight while
      // the current block is transformed (an exposed L2 round trip per tile was measured at 5 % of a 512-channel launch)
      const float4* cp = coef_row(cluster_id);
      float4 k0 = make_float4(0.f, 0.f, 0.f, 0.f), k1 = k0, k2 = k0, k3 = k0;
      if (cluster_id < p.num_work) { k0 = __ldg(cp); k1 = __ldg(cp + 1); k2 = __ldg(cp + 2); k3 = __ldg(cp + 3); }
      for (int work = cluster_id; work < p.num_work; work += num_clusters) {
        const int tile_id = work / p.ksplit;
        const int m0 = ((tile_id / p.n_tiles) * CL + static_cast<int>(cta_rank)) * TC_BM;
        const int y0 = (m0 / p.W) % p.H;
        for (int cb = 0; cb < p.c_blks; ++cb) {
          // h = y / 2 = fma(x, a / 2, s / 2) — exact, so silu(y) = h * tanh(h) + h is bit-identical to silu_f(fma(x, a, s))
          const float ca[8] = {0.5f * k0.x, 0.5f * k0.z, 0.5f * k1.x, 0.5f * k1.z, 0.5f * k2.x, 0.5f * k2.z, 0.5f * k3.x, 0.5f * k3.z};
          const float cs[8] = {0.5f * k0.y, 0.5f * k0.w, 0.5f * k1.y, 0.5f * k1.w, 0.5f * k2.y, 0.5f * k2.w, 0.5f * k3.y, 0.5f * k3.w};
          if (cb + 1 < p.c_blks) {
            cp += TC_BK * 2 / 4;
            k0 = __ldg(cp); k1 = __ldg(cp + 1); k2 = __ldg(cp + 2); k3 = __ldg(cp + 3);
          } else if (work + num_clusters < p.num_work) {
            cp = coef_row(work + num_clusters);
            k0 = __ldg(cp); k1 = __ldg(cp + 1); k2 = __ldg(cp + 2); k3 = __ldg(cp + 3);
          }
          mbar_wait(&empty_a[is * 3], ps ^ 1);      // the outer slots: free once the MMA has read their previous use
          mbar_wait(&empty_a[is * 3 + 2], ps ^ 1);
          mbar_wait(&full_a[is], (xf_par >> is) & 1u);   // the raw box has landed in the middle slot
          xf_par ^= 1u << is;
          // px advances by 16 per iteration, so the swizzle phases (px & 7, (px +- 1) & 7) are loop invariants; the
          // wrap-around destinations of the padding columns (px -+ (W - 1), W % 8 == 0) share them
          const uint32_t a1 = smem_base + (is * 3 + 1) * box + pl * 128u + (static_cast<uint32_t>(j ^ (pl & 7)) << 4);
          const uint32_t a0 = smem_base + (is * 3) * box + (static_cast<uint32_t>(j ^ ((pl + 1) & 7)) << 4);
          const uint32_t a2 = smem_base + (is * 3 + 2) * box + (static_cast<uint32_t>(j ^ ((pl - 1) & 7)) << 4);
#pragma unroll 2
          for (int it = 0; it < n_it; ++it) {
            const int px = pl + 16 * it;
            const int r = px >> log2w, x = px & wmask;
            const bool row_ok = static_cast<unsigned>(y0 - 1 + r) < static_cast<unsigned>(H);
            const uint4 u = lds128(a1 + it * 2048u);
            float v[8];
            float2 f;
            f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
            f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
            f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
            f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float h = fmaf(v[i], ca[i], cs[i]);
              float t;
              asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
              v[i] = silu ? fmaf(h, t, h) : h + h;
            }
            uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                 pack_bf16x2(v[6], v[7]));
            if (!row_ok) o = zero;    // rows outside the image are the convolution's padding
            sts128(a1 + it * 2048u, o);
            // dx = -1 box: operand pixel (r, x') = source pixel (r, x' - 1): this value lands at x + 1; column 0 is padding
            sts128_or_zero(x != wmask, a0 + static_cast<uint32_t>(px + 1) * 128u, o, a0 + static_cast<uint32_t>(px - wmask) * 128u);
            // dx = +1 box: this value lands at x - 1; column W - 1 is padding
            sts128_or_zero(x != 0, a2 + static_cast<uint32_t>(px - 1) * 128u, o, a2 + static_cast<uint32_t>(px + wmask) * 128u);
          }
          fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive_release_cluster(&xf_ready[is], 0);
          if (++is == TC_XF_SETS) { is = 0; ps ^= 1; }
        }
        // the fused-skip slabs go from TMA to the MMA untouched: only keep the set rotation in step
        for (int sb = 0; sb < p.skip_blks; sb += 3)
          if (++is == TC_XF_SETS) { is = 0; ps ^= 1; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are accessible to this warp
    const int row = quad * 32 + lane;
    constexpr int CH = BN < 32 ? 16 : 32;
    uint32_t tile = 0;
    for (int work = cluster_id; work < p.num_work; work += num_clusters, ++tile) {
      const uint32_t acc = tile % Cfg::ACC, acc_ph = (tile / Cfg::ACC) & 1;
      const int tile_id = work / p.ksplit, split = work - tile_id * p.ksplit;
      const int n_base = (tile_id % p.n_tiles) * BN;
      const int m_tile_idx = (tile_id / p.n_tiles) * CL + static_cast<int>(cta_rank);
      const int m_px = tc_tile_pixel(p, m_tile_idx, row);   // linear pixel index of this thread's accumulator row
      const bool valid = m_px >= 0;
      const int m = valid ? m_px : p.M;                     // (split-K workspace rows use the linear tile layout)
      const int b = valid ? m / p.HW : 0;
      const float* emb_row = p.emb ? p.emb + static_cast<size_t>(b) * p.emb_stride : nullptr;
      const uint32_t tmem_acc = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(quad * 32) << 16);
      // NHWC offset (elements) of output row mm at channel 0 (tap_mode 1: pixel (y, x) of this sub-pixel phase lands at
      // (2y + py, 2x + px) of the [B, 2H, 2W, C] output)
      auto row_offset = [&](int mm) -> size_t {
        if (p.tap_mode != 1) return static_cast<size_t>(mm) * p.cout;
        const int bb = mm / p.HW, pix = mm - bb * p.HW, y = pix / p.W, x = pix - y * p.W;
        return ((static_cast<size_t>(bb) * 2 * p.H + 2 * y + p.py) * (2 * p.W) + 2 * x + p.px) * p.cout;
      };
      const size_t o_row = valid ? row_offset(m) : 0;
      const uint32_t stg = smem_u32(s_out) + quad * (32 * 64);  // this warp's staging rows (32-bit shared address: LDS / STS)
      // rows this lane stores after the transpose: (lane / 4) + 8 i of the warp's 32, i < 4
      size_t roff[4];
      bool rok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int mm = tc_tile_pixel(p, m_tile_idx, quad * 32 + (lane >> 2) + 8 * i);
        rok[i] = mm >= 0;
        roff[i] = rok[i] ? row_offset(mm) : 0;
      }
      // bf16 residual (identity skip): a thread reads its own pixel row, and the tensor was written layers ago — the
      // chunk's 64 bytes are requested ONE CHUNK AHEAD (the first chunk before the wait for the accumulator), so the
      // DRAM / L2 latency overlaps the previous chunk's work instead of stalling every chunk (ncu source page, round 2:
      // 14 % of all stall samples of the 128 -> 128 launch sat on the first use of the loaded residual)
      constexpr bool RES_PF = CH == 32;
      const bool res_bf16 = RES_PF && valid && p.residual != nullptr && p.res_dtype == DT_BF16 && p.ksplit == 1 && !p.out_nchw;
      const __nv_bfloat16* res_row = nullptr;
      if (res_bf16)
        res_row = static_cast<const __nv_bfloat16*>(p.residual) +
                  (p.res_rows > 0 ? static_cast<size_t>(m % p.res_rows) * p.cout : o_row);
      uint4 rnx[4];
      auto load_res = [&](int nn) {
        if (res_bf16 && nn < p.cout) {
          const uint4* rp = reinterpret_cast<const uint4*>(res_row + nn);
#pragma unroll
          for (int k = 0; k < 4; ++k) rnx[k] = __ldg(rp + k);
        }
      };
      load_res(n_base);
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += CH) {
        const int n = n_base + c;
        if (n >= p.cout) break;  // partially filled last channel tile (cout % BN != 0): nothing to store
        uint4 rcur[4];
        if constexpr (RES_PF) {
#pragma unroll
          for (int k = 0; k < 4; ++k) rcur[k] = rnx[k];
          if (c + CH < BN) load_res(n + CH);
        }
        float v[CH], add[CH];
        if constexpr (CH == 32) {
          uint32_t r[32];
          tmem_ld32(tmem_acc + c, r);
          // bias / embedding rows are fetched while the TMEM load is in flight (they were the epilogue's longest stall)
#pragma unroll
          for (int j = 0; j < CH; ++j) add[j] = 0.f;
          if (p.ksplit == 1) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                add[j] = t.x; add[j + 1] = t.y; add[j + 2] = t.z; add[j + 3] = t.w;
              }
            }
            if (emb_row) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(emb_row + n + j));
                add[j] += t.x; add[j + 1] += t.y; add[j + 2] += t.z; add[j + 3] += t.w;
              }
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
          uint32_t r[16];
          tmem_ld16(tmem_acc + c, r);
#pragma unroll
          for (int j = 0; j < CH; ++j) add[j] = 0.f;
          if (p.ksplit == 1) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
                add[j] = t.x; add[j + 1] = t.y; add[j + 2] = t.z; add[j + 3] = t.w;
              }
            }
            if (emb_row) {
#pragma unroll
              for (int j = 0; j < CH; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(emb_row + n + j));
                add[j] += t.x; add[j + 1] += t.y; add[j + 2] += t.z; add[j + 3] += t.w;
              }
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        }
        if (p.ksplit > 1) {  // split-K: raw fp32 partial tile -> workspace; splitk_finish_kernel applies the epilogue
          float4* wp = reinterpret_cast<float4*>(
              p.ws + (static_cast<size_t>(split) * p.m_pad + m_tile_idx * TC_BM + row) * p.cout + n);
#pragma unroll
          for (int j = 0; j < CH; j += 4) wp[j / 4] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          continue;
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] += add[j];
        if (p.act == STEDM_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = gelu_erf_fast(v[j]);
        }
        const size_t o = o_row + n;
        if (valid && p.residual) {
          // a residual with fewer samples (the part of a sum that is shared by the cond / uncond halves of a guided
          // batch) is read at row m % res_rows
          const size_t o_res = p.res_rows > 0 ? static_cast<size_t>(m % p.res_rows) * p.cout + n : o;
          if (p.res_dtype == DT_BF16) {
            const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + o_res);
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              uint4 u;
              if constexpr (RES_PF) u = res_bf16 ? rcur[j / 8] : rp[j / 8];
              else u = rp[j / 8];
              float2 f;
              f = unpack_bf16x2(u.x); v[j] += f.x; v[j + 1] += f.y;
              f = unpack_bf16x2(u.y); v[j + 2] += f.x; v[j + 3] += f.y;
              f = unpack_bf16x2(u.z); v[j + 4] += f.x; v[j + 5] += f.y;
              f = unpack_bf16x2(u.w); v[j + 6] += f.x; v[j + 7] += f.y;
            }
          } else {
            const float4* rp = reinterpret_cast<const float4*>(static_cast<const float*>(p.residual) + o_res);
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const float4 t = rp[j / 4];
              v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
          }
        }
        if (p.out_nchw) {
          if (valid) {
            // channel-major output (eps / image heads in fp32; V^T for the decoder attention in bf16): consecutive
            // lanes = consecutive pixels -> coalesced per channel plane
            const size_t base = static_cast<size_t>(b) * p.cout_store * p.HW + (m - b * p.HW);
            if (p.out_dtype == DT_F32) {
              float* op = static_cast<float*>(p.out) + base;
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (n + j < p.cout_store) op[static_cast<size_t>(n + j) * p.HW] = v[j];
            } else {
              __nv_bfloat16* op = static_cast<__nv_bfloat16*>(p.out) + base;
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (n + j < p.cout_store) op[static_cast<size_t>(n + j) * p.HW] = __float2bfloat16_rn(v[j]);
            }
          }
        } else if constexpr (CH == 32) {
          // NHWC output through a per-warp shared-memory transpose: a thread owns one pixel row of the accumulator,
          // so direct stores are 32 half-written sectors per instruction; staged, 4 lanes write one row's 64
          // contiguous bytes (2 whole sectors).  16-byte slots are XOR-swizzled: 4 wavefronts per access, the minimum.
          const int jj = lane & 3;
          if (p.out_dtype == DT_BF16) {
            if (p.stats_out != nullptr && !valid) {     // rows past the end of the tensor count as zeros in the statistics
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                     make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7])));
            __syncwarp();
            if (p.stats_out != nullptr) {
              // GroupNorm statistics of the tensor being written, taken from the STAGED bf16 tile: the values the
              // consumer will actually normalise, and 16 conflict-free LDS.32 + 4 shuffles per lane instead of the
              // 62-shuffle transposing reduction (25 % of this warp's stall samples on the 128-wide channel tiles).
              // lane -> (channel pair cp, half of the 32 rows); the halves walk rows of opposite parity (rows r and
              // r + 1 sit on disjoint banks: a 64-byte row covers 16 of the 32 banks)
              const int cp = lane & 15, rh = lane >> 4;
              const uint32_t cbase = stg + static_cast<uint32_t>(cp & 3) * 4;
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int r = rh * 16 + (rh ? (i ^ 1) : i);
                const float2 f = unpack_bf16x2(lds32(cbase + r * 64 + (((cp >> 2) ^ ((r >> 1) & 3)) << 4)));
                s0 += f.x; s1 += f.y;
                q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
              }
              s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
              q0 += __shfl_xor_sync(0xffffffffu, q0, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
              if (lane < 16)
                *reinterpret_cast<float4*>(&s_stats[((quad * BN) + c + 2 * cp) * 2]) = make_float4(s0, q0, s1, q1);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int rr = (lane >> 2) + 8 * i;
              if (rok[i])
                *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + roff[i] + n + jj * 8) =
                    lds128(stg + rr * 64 + ((jj ^ ((rr >> 1) & 3)) << 4));
            }
            __syncwarp();
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {  // 16 fp32 channels = 64 B per pass
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4),
                       make_uint4(__float_as_uint(v[16 * h + 4 * j]), __float_as_uint(v[16 * h + 4 * j + 1]),
                                  __float_as_uint(v[16 * h + 4 * j + 2]), __float_as_uint(v[16 * h + 4 * j + 3])));
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = (lane >> 2) + 8 * i;
                if (rok[i])
                  *reinterpret_cast<uint4*>(static_cast<float*>(p.out) + roff[i] + n + 16 * h + jj * 4) =
                      lds128(stg + rr * 64 + ((jj ^ ((rr >> 1) & 3)) << 4));
              }
              __syncwarp();
            }
          }
        } else if (valid) {
          if (p.out_dtype == DT_BF16) {
            uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + o);
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              uint4 u;
              u.x = pack_bf16x2(v[j], v[j + 1]);
              u.y = pack_bf16x2(v[j + 2], v[j + 3]);
              u.z = pack_bf16x2(v[j + 4], v[j + 5]);
              u.w = pack_bf16x2(v[j + 6], v[j + 7]);
              op[j / 8] = u;
            }
          } else {
            float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + o);
#pragma unroll
            for (int j = 0; j < CH; j += 4) op[j / 4] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        if constexpr (CH == 32) {
          if (p.stats_out != nullptr && !(p.out_dtype == DT_BF16 && !p.out_nchw)) {   // (bf16 NHWC: done from the staged tile)
            // GroupNorm statistics of the tensor being written (K7's statistics pass folded into its producer):
            // per-channel sum and sum of squares over this warp's 32 pixel rows; lane L ends up with channel c + L
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = valid ? v[j] : 0.f;
              sq[j] = v[j] * v[j];
            }
            const float cs = warp_transpose_reduce32(v, lane);
            const float cq = warp_transpose_reduce32(sq, lane);
            *reinterpret_cast<float2*>(&s_stats[((quad * BN) + c + lane) * 2]) = make_float2(cs, cq);
          }
        }
      }
      // all TMEM reads of this accumulator are complete (tcgen05.wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      if constexpr (PAIR) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);  // the leader's MMA thread waits for both CTAs
      else mbar_arrive(&tmem_empty_bar[acc]);
      if constexpr (CH == 32) {
        if (p.stats_out != nullptr) {
          // fold the four warps' partials in a fixed order and publish this tile's per-channel statistics
          named_bar_sync(1, 128);
          const int m_tile = (tile_id / p.n_tiles) * CL + static_cast<int>(cta_rank);
          if (m_tile * TC_BM < p.M) {
            float* dst = p.stats_out + (static_cast<size_t>(p.stats_tile_base + m_tile) * p.cout + n_base) * 2;
            for (int ch = threadIdx.x - 64; ch < BN; ch += 128) {
              float a = 0.f, b2 = 0.f;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const float2 t = *reinterpret_cast<const float2*>(&s_stats[((w * BN) + ch) * 2]);
                a += t.x;
                b2 += t.y;
              }
              *reinterpret_cast<float2*>(dst + ch * 2) = make_float2(a, b2);
            }
          }
          named_bar_sync(1, 128);
        }
      }
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all(); else __syncthreads();  // no CTA may exit while a peer can still write into it
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Split-K second pass: out = epilogue(sum over splits of the fp32 partial tiles, in split order: deterministic).
// Block = one 128-pixel tile x 16 output channels, thread = (pixel row, 8-channel octet): the ksplit partial loads of a
// thread are independent and issued together (this pass is latency-bound: its launches have at most a few dozen tiles).
// When the consumer is a GroupNorm the block also reduces the tile's per-channel sum / sum of squares in a fixed order,
// exactly what the single-pass epilogue publishes — a split launch keeps the statistics pass folded into its producer.
constexpr int SKF_CH = 16;
__global__ void __launch_bounds__(256) splitk_finish_kernel(const TcParams p) {
  __shared__ float s_red[TC_BM][SKF_CH][2];
  const int tile = blockIdx.x, n0 = blockIdx.y * SKF_CH;
  const int r = threadIdx.x >> 1, oct = threadIdx.x & 1, n = n0 + oct * 8;
  const int m = tile * TC_BM + r;
  const bool live = m < p.M && n < p.cout;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  if (live) {
    const float* wp = p.ws + static_cast<size_t>(m) * p.cout + n;
    const size_t ss = static_cast<size_t>(p.m_pad) * p.cout;
    int s = 0;
    for (; s + 4 <= p.ksplit; s += 4) {   // fixed order: deterministic; 8 independent 16-byte loads in flight
      float4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[u] = *reinterpret_cast<const float4*>(wp + (s + u) * ss);
        b[u] = *reinterpret_cast<const float4*>(wp + (s + u) * ss + 4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        v[0] += a[u].x; v[1] += a[u].y; v[2] += a[u].z; v[3] += a[u].w;
        v[4] += b[u].x; v[5] += b[u].y; v[6] += b[u].z; v[7] += b[u].w;
      }
    }
    for (; s < p.ksplit; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(wp + s * ss), b = *reinterpret_cast<const float4*>(wp + s * ss + 4);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    const int b = m / p.HW;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float add = p.bias ? p.bias[n + j] : 0.f;
      if (p.emb) add += p.emb[static_cast<size_t>(b) * p.emb_stride + n + j];
      v[j] += add;
      if (p.act == STEDM_ACT_GELU) v[j] = gelu_erf(v[j]);
    }
    size_t o = static_cast<size_t>(m) * p.cout + n;
    if (p.tap_mode == 1) {
      const int pix = m - b * p.HW, y = pix / p.W, x = pix - y * p.W;
      o = ((static_cast<size_t>(b) * 2 * p.H + 2 * y + p.py) * (2 * p.W) + 2 * x + p.px) * p.cout + n;
    }
    if (p.residual) {
      const size_t o_res = p.res_rows > 0 ? static_cast<size_t>(m % p.res_rows) * p.cout + n : o;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] += (p.res_dtype == DT_F32) ? static_cast<const float*>(p.residual)[o_res + j]
                                        : __bfloat162float(static_cast<const __nv_bfloat16*>(p.residual)[o_res + j]);
    }
    if (p.out_nchw) {
      const size_t base = static_cast<size_t>(b) * p.cout_store * p.HW + (m - b * p.HW);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n + j >= p.cout_store) continue;
        if (p.out_dtype == DT_F32) static_cast<float*>(p.out)[base + static_cast<size_t>(n + j) * p.HW] = v[j];
        else static_cast<__nv_bfloat16*>(p.out)[base + static_cast<size_t>(n + j) * p.HW] = __float2bfloat16_rn(v[j]);
      }
    } else if (p.out_dtype == DT_BF16) {
      *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + o) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    } else {
      float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + o);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  if (p.stats_out == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // (the statistics describe the STORED tensor: bf16-rounded values when the output is bf16, as in the single-pass epilogue)
    const float x = live ? (p.out_dtype == DT_BF16 ? __bfloat162float(__float2bfloat16_rn(v[j])) : v[j]) : 0.f;
    s_red[r][oct * 8 + j][0] = x;
    s_red[r][oct * 8 + j][1] = x * x;
  }
  __syncthreads();
  // 32 outputs (16 channels x {sum, sumsq}) x 8 threads: fixed strided partition + fixed shuffle tree = deterministic
  const int o = threadIdx.x >> 3, sub = threadIdx.x & 7;
  const int ch = o >> 1, which = o & 1;
  float acc = 0.f;
#pragma unroll 4
  for (int l = sub; l < TC_BM; l += 8) acc += s_red[l][ch][which];
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0 && n0 + ch < p.cout && tile * TC_BM < p.M)
    p.stats_out[(static_cast<size_t>(p.stats_tile_base + tile) * p.cout + n0 + ch) * 2 + which] = acc;
}

int tc_num_sms() {
  static int cached[64] = {0};  // per device (a process may drive several GPUs)
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cached[dev & 63];
  if (n == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n = v;
  }
  return n;
}

template <int BN, int CL, bool PAIR, bool HALO, bool XF = false>
int launch_tc_impl(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mw, const CUtensorMap& ms0,
              const CUtensorMap& ms1, TcParams p, cudaStream_t stream) {
  using Cfg = TcCfg<BN, PAIR>;
  constexpr int SMEM = XF ? Cfg::SMEM_BYTES_XF : Cfg::SMEM_BYTES;
  static DeviceOnce configured;  // the attribute is per function AND per device
  if (configured.needed()) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, CL, PAIR, HALO, XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("conv_tc: cudaFuncSetAttribute(%d B smem): %s", SMEM, cudaGetErrorString(e));
      return ERR_CUDA;
    }
    configured.done();
  }
  if (XF) {
    // 2 sets x 3 boxes, the weight ring in what is left of the pool
    p.a_pool_bytes = TC_XF_SETS * 3 * p.a_buf_bytes;
    p.nb = (TC_XF_POOL - p.a_pool_bytes) / Cfg::B_BYTES_PAD;
    if (p.nb > TC_XF_MAX_NB) p.nb = TC_XF_MAX_NB;
    static const int nb_cap = [] { const char* e = getenv("STEDM_XF_NB"); return e ? atoi(e) : 0; }();  // A/B: shallower ring
    if (nb_cap >= 3 && p.nb > nb_cap) p.nb = nb_cap;
    if (p.nb < 3 || p.a_buf_bytes < Cfg::A_BYTES) {
      set_error("conv_tc: halo box of %d B leaves no room for the weight ring in GroupNorm-fused mode", p.a_buf_bytes);
      return ERR_ARG;
    }
    p.na = TC_XF_SETS;
  } else if (!HALO) {
    p.a_buf_bytes = Cfg::A_BYTES;
    p.na = Cfg::STAGES;
    p.a_row_bytes = 0;
  } else if (p.na > Cfg::STAGES || p.na * p.a_buf_bytes > Cfg::A_POOL) {
    set_error("conv_tc: halo ring (%d x %d B) does not fit the %d B pool", p.na, p.a_buf_bytes, Cfg::A_POOL);
    return ERR_ARG;
  }
  const int m_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int m_groups = (m_tiles + CL - 1) / CL;  // an odd tile count gets one all-out-of-bounds tile (zero-filled loads, no stores)
  p.num_work = m_groups * p.n_tiles * p.ksplit;
  const int max_clusters = tc_num_sms() * Cfg::MIN_BLOCKS / CL;   // persistent: one resident wave
  const int clusters = p.num_work < max_clusters ? p.num_work : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters) * CL);
  cfg.blockDim = dim3(XF ? TC_THREADS_XF : TC_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TcParams pk = p;
  if (p.ksplit > 1) pk.stats_out = nullptr;   // split launches publish the tile statistics from the finish pass
  const cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, CL, PAIR, HALO, XF>, ma0, ma1, mw, ms0, ms1, pk);
  if (e != cudaSuccess) {
    set_error("conv_tc: launch failed: %s", cudaGetErrorString(e));
    return ERR_CUDA;
  }
  int rc = check_launch("conv_tc");
  if (rc == 0 && p.ksplit > 1) {
    splitk_finish_kernel<<<dim3(static_cast<unsigned>(m_tiles), static_cast<unsigned>((p.cout + SKF_CH - 1) / SKF_CH)), 256, 0, stream>>>(p);
    rc = check_launch("conv_tc split-K finish");
  }
  return rc;
}

template <int BN, int CL, bool PAIR = false>
int launch_tc(const CUtensorMap& ma0, const CUtensorMap& ma1, const CUtensorMap& mw, const CUtensorMap& ms0,
              const CUtensorMap& ms1, const TcParams& p, cudaStream_t stream) {
  if constexpr (BN == 256 && PAIR) {
    if (p.gn_coef != nullptr) return launch_tc_impl<BN, CL, PAIR, true, true>(ma0, ma1, mw, ms0, ms1, p, stream);
  }
  return p.halo ? launch_tc_impl<BN, CL, PAIR, true>(ma0, ma1, mw, ms0, ms1, p, stream)
                : launch_tc_impl<BN, CL, PAIR, false>(ma0, ma1, mw, ms0, ms1, p, stream);
}

// Tile / cluster / split-K plan of one launch: shared by stedm_conv_tc and stedm_conv_tc_workspace_bytes.
// Split-K is chosen when the output tiles would occupy less than half of the persistent grid and K is deep
// (small batches: the deep layers have a handful of tiles but hundreds of K slabs).
struct TcPlan {
  int bn, cl, ksplit, kb_per_split, m_pad;
  bool pair;
  size_t ws_bytes;
};

TcPlan tc_plan(long long M, int cout, int num_kb, bool want_stats) {
  TcPlan t;
  t.bn = (cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : (cout % 64 == 0 ? 64 : 16));
  if (t.bn == 16 && cout % 32 == 0 && cout > 64) {
    // widths like 96 / 288 (Swin-V2-T stage 1): a partially filled last tile beats 16-wide tiles; least padding wins
    const int pad128 = (cout + 127) / 128 * 128, pad64 = (cout + 63) / 64 * 64;
    t.bn = pad128 <= pad64 ? 128 : 64;
  }
  t.cl = (t.bn >= 128 && M > TC_BM && g_tc_cluster_enabled) ? 2 : 1;
  t.pair = t.cl == 2 && g_tc_pair_enabled;
  const int m_tiles = static_cast<int>((M + TC_BM - 1) / TC_BM);
  const int m_groups = (m_tiles + t.cl - 1) / t.cl;
  const int tiles = m_groups * ((cout + t.bn - 1) / t.bn);
  const int max_clusters = tc_num_sms() * (t.bn >= 256 ? 1 : 2) / t.cl;
  t.m_pad = m_groups * t.cl * TC_BM;
  t.ksplit = 1;
  t.kb_per_split = num_kb;
  t.ws_bytes = 0;
  (void)want_stats;   // the split-K finish pass publishes the GroupNorm tile statistics too
  if (g_tc_splitk_enabled && t.bn >= 64 && tiles * 2 <= max_clusters && num_kb >= 8) {
    int ks = max_clusters / tiles;
    if (ks > num_kb / 4) ks = num_kb / 4;
    if (ks >= 2) {
      t.kb_per_split = (num_kb + ks - 1) / ks;
      t.ksplit = (num_kb + t.kb_per_split - 1) / t.kb_per_split;
      t.ws_bytes = static_cast<size_t>(t.ksplit) * t.m_pad * cout * sizeof(float);
    }
  }
  return t;
}

}  // namespace

extern "C" long long stedm_conv_tc_workspace_bytes(const stedm_conv_desc* d) {
  if (d == nullptr || d->cout < 16 || d->cout % 16 != 0 || d->batch <= 0) return 0;
  const int st = d->stride == 2 ? 2 : 1;
  const long long M = static_cast<long long>(d->batch) * (d->in_h / st) * (d->in_w / st);
  const int taps = d->tap_mode == 1 ? 4 : d->ksize * d->ksize;
  const int num_kb = taps * ((d->c0 + d->c1 + TC_BK - 1) / TC_BK) + (d->skip_x0 ? (d->skip_c0 + d->skip_c1) / TC_BK : 0);
  return static_cast<long long>(tc_plan(M, d->cout, num_kb, d->stats_out != nullptr).ws_bytes);
}

namespace {

// Everything stedm_conv_tc decides before it touches the device: argument checks, tile geometry, channel tile /
// cluster / split-K plan and the halo-mode choice (also exported as stedm_conv_tc_plan for CPU-side tests).
struct TcLaunch {
  int tw, th, tb, x1b, tile2d;
  long long M;
  int ctot, taps, c_blks, skip_c, skip_blks, n_t, halo, halo_na, halo_bytes;
  TcPlan plan;
};

int tc_prepare(const stedm_conv_desc* d, TcLaunch* L) {
  STEDM_REQUIRE(d != nullptr, "conv_tc: null descriptor");
  STEDM_REQUIRE(d->in_dtype == DT_BF16, "conv_tc: operands must be bf16");
  STEDM_REQUIRE(d->upsample == 0, "conv_tc: nearest upsampling runs as four sub-pixel phase convolutions (tap_mode 1)");
  STEDM_REQUIRE(d->stride == 1 || (d->stride == 2 && d->ksize == 3 && d->tap_mode == 0 && d->c1 == 0 && d->skip_x0 == nullptr &&
                                   d->gn_coef == nullptr && d->in_h % 2 == 0 && d->in_w % 2 == 0 && d->x0_pix_stride == 0),
                "conv_tc: stride 2 (openaimodel.py:164-166) takes one dense 3x3 source with even height and width");
  STEDM_REQUIRE(!d->out_nchw || (d->residual == nullptr && d->cout_store >= 0 && d->cout_store <= d->cout),
                "conv_tc: NCHW output takes no residual and needs cout_store <= cout");
  STEDM_REQUIRE(d->ksize == 1 || d->ksize == 3, "conv_tc: ksize %d unsupported", d->ksize);
  STEDM_REQUIRE(d->tap_mode == 0 || (d->tap_mode == 1 && d->phase >= 0 && d->phase < 4 && d->residual == nullptr &&
                                     d->out_nchw == 0 && d->c1 == 0),
                "conv_tc: sub-pixel phase convolution takes one source, no residual, NHWC output, phase in [0,4)");
  // plain GEMM (1 tap, one source): the last K slab may be partial — TMA zero-fills both operands beyond c0
  const bool plain_gemm = d->ksize == 1 && d->tap_mode == 0 && d->c1 == 0;
  STEDM_REQUIRE(d->c0 > 0 && (d->c0 % TC_BK == 0 || (plain_gemm && d->c0 % 8 == 0)) && d->c1 % TC_BK == 0 &&
                    (d->c1 == 0 || d->x1),
                "conv_tc: channel counts must be multiples of 64 (8 for a plain GEMM) (%d, %d)", d->c0, d->c1);
  STEDM_REQUIRE(d->cout >= 16 && d->cout % 16 == 0, "conv_tc: cout %d must be a multiple of 16", d->cout);
  STEDM_REQUIRE(d->act == STEDM_ACT_NONE || d->act == STEDM_ACT_GELU, "conv_tc: unknown activation %d", d->act);
  // GEMM rows = OUTPUT pixels: for the stride-2 convolution the maps below are tiled over the half-resolution output and
  // the TMA boxes walk the input with a traversal stride of 2 (the gather happens in the TMA unit, no im2col tensor)
  const int H = d->in_h / d->stride, W = d->in_w / d->stride, B = d->batch;
  STEDM_REQUIRE(B > 0 && H > 0 && W > 0, "conv_tc: bad shape");
  // tile geometry: 128 consecutive pixels of the flattened (b, y, x) index must form a TMA box
  int tw, th, tb;
  if (W >= TC_BM) {
    STEDM_REQUIRE(W % TC_BM == 0, "conv_tc: width %d must divide or be a multiple of 128", W);
    tw = TC_BM; th = 1; tb = 1;
  } else {
    STEDM_REQUIRE(TC_BM % W == 0, "conv_tc: width %d must divide or be a multiple of 128", W);
    tw = W;
    const int rows = TC_BM / W;
    if (H >= rows) {
      STEDM_REQUIRE(H % rows == 0, "conv_tc: height %d incompatible with the 128-pixel tile", H);
      th = rows; tb = 1;
    } else {
      STEDM_REQUIRE(rows % H == 0, "conv_tc: height %d incompatible with the 128-pixel tile", H);
      th = H; tb = rows / H;
    }
  }
  const int x1b = (d->c1 > 0 && d->x1_batch > 0 && d->x1_batch != B) ? d->x1_batch : 0;
  if (x1b > 0)
    STEDM_REQUIRE((static_cast<long long>(x1b) * H * W) % TC_BM == 0 || tb == 1,
                  "conv_tc: broadcast source batch %d not tile aligned", x1b);
  const long long M = static_cast<long long>(B) * H * W;
  STEDM_REQUIRE(M < (1LL << 31), "conv_tc: too many pixels");

  const int ctot = d->c0 + d->c1, taps = d->tap_mode == 1 ? 4 : d->ksize * d->ksize;
  // channel tile, 2-CTA cluster (cta_group::2 pair or weight multicast) and split-K plan
  const int c_blks = (ctot + TC_BK - 1) / TC_BK;
  // fused 1x1 skip convolution: a second input [skip_x0 | skip_x1] of the output's spatial size, extra K slabs
  const int skip_c = d->skip_x0 ? d->skip_c0 + d->skip_c1 : 0;
  const int skip_blks = skip_c / TC_BK;
  TcPlan plan = tc_plan(M, d->cout, taps * c_blks + skip_blks, d->stats_out != nullptr);
  if (plan.ksplit > 1 && (d->workspace == nullptr || static_cast<size_t>(d->workspace_bytes) < plan.ws_bytes)) {
    plan.ksplit = 1;                        // no (or too small a) workspace: single pass
    plan.kb_per_split = taps * c_blks + skip_blks;
  }
  // halo mode: tiles of th >= 2 whole image rows of one sample; the activation box grows by the n_t - 1 halo rows and
  // is loaded once per (channel block, horizontal tap).  The ring holds as many boxes as fit the STAGES * 16 KB pool.
  const int n_t = d->tap_mode == 1 ? 2 : d->ksize;  // taps per dimension
  int halo = 0, halo_na = 0, halo_bytes = 0;
  // (fused-skip slabs take a whole box-sized ring slot each: with many of them the shallower ring costs more than the
  //  halo saves — measured break-even near one skip slab per five tap slabs)
  // 2-D pixel tiles (16 columns x 8 rows) for the 3x3 convolutions on maps at least 32 wide: the halo box then carries
  // 2 extra rows per 8 whatever the width (full-row tiles: 2 per 4 at W = 32, 2 per 2 at W = 64, no halo mode from W = 128)
  // OPT-IN (STEDM_TC_TILE2D=1, read per call so tests can toggle it): measured on the bench workload it does not pay —
  // the 64 x 64 layers are unchanged (0.526 / 0.241 ms vs 0.530 / 0.236), the sub-pixel phase convolutions at 32 x 32 and
  // the decoder's 128 -> 128 at 256 x 256 lose 5-8 % (a tile's stores and residual reads become eight 4 KB runs instead
  // of one 32 KB run), only the 256 -> 128 at 256 x 256 gains 7 % (profiles/r02_tile2d_ab.txt)
  const char* t2e = getenv("STEDM_TC_TILE2D");
  const bool tile2d_enabled = t2e != nullptr && t2e[0] == '1';
  int tile2d = 0;
  if (tile2d_enabled && g_tc_halo_enabled && d->ksize == 3 && d->stride == 1 && W >= 32 && W % 16 == 0 && H % 8 == 0 &&
      plan.ksplit == 1 && d->gn_coef == nullptr && skip_blks * 5 <= taps * c_blks) {
    tile2d = 1;
    tw = 16; th = 8; tb = 1;
  }
  if (d->gn_coef != nullptr) {
    // GroupNorm (+ SiLU) in the operand path: built on the CTA-pair halo pipeline with 256-wide channel tiles
    plan.ksplit = 1;
    plan.kb_per_split = taps * c_blks + skip_blks;
    halo_bytes = (th + 2) * W * TC_BK * 2;
    const bool geometry = d->ksize == 3 && d->tap_mode == 0 && tb == 1 && th >= 2 && tw == W && W >= 8 && H >= th + 2 &&
                          halo_bytes % 1024 == 0;
    STEDM_REQUIRE(geometry && plan.bn == 256 && plan.pair && ctot % TC_BK == 0 &&
                      TC_XF_SETS * 3 * halo_bytes + 3 * (256 / 2) * TC_BK * 2 <= TC_XF_POOL,
                  "conv_tc: gn_coef (GroupNorm in the operand path) needs a 3x3 convolution over whole-row tiles of one "
                  "sample (W in 8..64, H >= rows per tile + 2), cout %% 256 == 0 and more than one pixel tile");
    STEDM_REQUIRE(d->gn_cstride >= d->gn_c_off + ctot && d->gn_c_off % 8 == 0 && d->gn_cstride % 4 == 0,
                  "conv_tc: gn_coef row of %d channels does not cover channels [%d, %d)", d->gn_cstride, d->gn_c_off,
                  d->gn_c_off + ctot);
    halo = 1;
    halo_na = TC_XF_SETS;
  } else if (g_tc_halo_enabled && d->stride == 1 && d->ksize == 3 && tb == 1 && th >= 2 && (tw == W || tile2d) && W >= 8 &&
             H >= th + n_t - 1 && plan.ksplit == 1 && skip_blks * 5 <= taps * c_blks) {
    halo_bytes = (th + n_t - 1) * tw * TC_BK * 2;
    const int stages = tc_stages(plan.bn, plan.pair);
    halo_na = stages * TC_BM * TC_BK * 2 / halo_bytes;
    if (halo_na > stages) halo_na = stages;
    halo = halo_na >= 2 && halo_bytes % 1024 == 0;
  }
  L->tw = tw; L->th = th; L->tb = tb; L->x1b = x1b; L->M = M; L->tile2d = tile2d;
  L->ctot = ctot; L->taps = taps; L->c_blks = c_blks; L->skip_c = skip_c; L->skip_blks = skip_blks;
  L->n_t = n_t; L->halo = halo; L->halo_na = halo_na; L->halo_bytes = halo_bytes; L->plan = plan;
  return 0;
}

}  // namespace

// out[m][n] = round(src[m % src_rows][n] + emb[(m / hw) * emb_stride + n]) (+ the per-(128-row tile, channel) statistics
// a GroupNorm consumer folds).  Used where a convolution's input is shared by the G halves of a guided batch and only the
// per-sample embedding differs (ResBlockStyle's first convolution, openaimodel.py:291-297 after :277-286): the convolution
// runs once per distinct input, this pass expands it.  A streaming pass (HBM / L2 bound): block = one 128-row tile x 256
// channels, thread = (8-channel octet, row group) so that a warp reads 1 KB of one row; the statistics describe the STORED
// values (bf16-rounded when the output is bf16, as in the convolution's epilogue) and are summed in a fixed order.
namespace {
constexpr int RAE_CH = 256, RAE_GROUPS = 8;
__global__ void __launch_bounds__(256) rows_add_emb_kernel(const float* __restrict__ src, int src_rows,
                                                           const float* __restrict__ emb, int emb_stride,
                                                           void* __restrict__ out, int out_bf16, int M, int hw, int C,
                                                           float* __restrict__ stats) {
  __shared__ float s_red[RAE_GROUPS][RAE_CH][2];
  const int tile = blockIdx.x, oct = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int n = blockIdx.y * RAE_CH + oct * 8;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (n < C) {
#pragma unroll 4
    for (int i = 0; i < TC_BM / RAE_GROUPS; ++i) {
      const int m = tile * TC_BM + rg + i * RAE_GROUPS;
      if (m >= M) break;
      const float* sp = src + static_cast<size_t>(m % src_rows) * C + n;
      const float* ep = emb + static_cast<size_t>(m / hw) * emb_stride + n;
      const float4 a = *reinterpret_cast<const float4*>(sp), b = *reinterpret_cast<const float4*>(sp + 4);
      const float4 ea = *reinterpret_cast<const float4*>(ep), eb = *reinterpret_cast<const float4*>(ep + 4);
      float v[8] = {a.x + ea.x, a.y + ea.y, a.z + ea.z, a.w + ea.w, b.x + eb.x, b.y + eb.y, b.z + eb.z, b.w + eb.w};
      const size_t o = static_cast<size_t>(m) * C + n;
      if (out_bf16) {
        const uint4 u = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                   pack_bf16x2(v[6], v[7]));
        *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + o) = u;
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
      } else {
        float4* op = reinterpret_cast<float4*>(static_cast<float*>(out) + o);
        op[0] = make_float4(v[0], v[1], v[2], v[3]);
        op[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        q[j] = fmaf(v[j], v[j], q[j]);
      }
    }
  }
  if (stats == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_red[rg][oct * 8 + j][0] = s[j];
    s_red[rg][oct * 8 + j][1] = q[j];
  }
  __syncthreads();
  const int c = blockIdx.y * RAE_CH + threadIdx.x;
  if (c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int g = 0; g < RAE_GROUPS; ++g) {
      a += s_red[g][threadIdx.x][0];
      b += s_red[g][threadIdx.x][1];
    }
    *reinterpret_cast<float2*>(stats + (static_cast<size_t>(tile) * C + c) * 2) = make_float2(a, b);
  }
}
}  // namespace

extern "C" int stedm_rows_add_emb(const float* src, long long src_rows, const float* emb, int emb_stride, void* out,
                                  int out_dtype, long long rows_out, int hw, int c, float* stats_out, void* stream) {
  STEDM_REQUIRE(src && emb && out && src_rows > 0 && rows_out > 0 && hw > 0 && c > 0 && c % 8 == 0 && rows_out % hw == 0 &&
                    src_rows % hw == 0 && rows_out < (1LL << 31) && src_rows < (1LL << 31) && emb_stride % 4 == 0,
                "rows_add_emb: bad argument (channels must be a multiple of 8, row counts multiples of hw, emb rows 16-byte "
                "aligned)");
  STEDM_REQUIRE(out_dtype == DT_BF16 || out_dtype == DT_F32, "rows_add_emb: bad output dtype");
  STEDM_REQUIRE(stats_out == nullptr || hw % TC_BM == 0, "rows_add_emb: tile statistics need hw %% 128 == 0");
  const int m_tiles = static_cast<int>((rows_out + TC_BM - 1) / TC_BM);
  rows_add_emb_kernel<<<dim3(static_cast<unsigned>(m_tiles), static_cast<unsigned>((c + RAE_CH - 1) / RAE_CH)), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(src, static_cast<int>(src_rows), emb, emb_stride, out,
                                                             out_dtype == DT_BF16, static_cast<int>(rows_out), hw, c, stats_out);
  return check_launch("rows_add_emb");
}

extern "C" int stedm_conv_tc_plan(const stedm_conv_desc* d, int32_t* out8) {
  STEDM_REQUIRE(out8 != nullptr, "conv_tc_plan: null output");
  TcLaunch L;
  const int rc = tc_prepare(d, &L);
  if (rc) return rc;
  const int stages = tc_stages(L.plan.bn, L.plan.pair);
  out8[0] = L.plan.bn; out8[1] = L.plan.cl; out8[2] = L.plan.pair ? 1 : 0; out8[3] = L.halo;
  out8[4] = L.halo ? L.halo_na : stages; out8[5] = L.halo ? L.halo_bytes : TC_BM * TC_BK * 2;
  out8[6] = L.plan.ksplit; out8[7] = L.taps * L.c_blks + L.skip_blks;
  return 0;
}

extern "C" int stedm_conv_tc(const stedm_conv_desc* d, void* stream) {
  STEDM_REQUIRE(d && d->x0 && d->weight && d->out, "conv_tc: null pointer");
  TcLaunch L;
  {
    const int rc = tc_prepare(d, &L);
    if (rc) return rc;
  }
  const int cs = d->stride;                                  // 2: strided boxes over the full-resolution input
  const int H = d->in_h / cs, W = d->in_w / cs, B = d->batch;  // output map = GEMM rows
  const int tw = L.tw, th = L.th, tb = L.tb, x1b = L.x1b;
  const long long M = L.M;
  const int ctot = L.ctot, taps = L.taps, c_blks = L.c_blks, skip_c = L.skip_c, skip_blks = L.skip_blks;
  const TcPlan plan = L.plan;
  const int n_t = L.n_t, halo = L.halo, halo_na = L.halo_na, halo_bytes = L.halo_bytes;
  const uint32_t abox_h = static_cast<uint32_t>(halo ? th + n_t - 1 : th);

  CUtensorMap ma0, ma1, mw;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(d->c0), static_cast<uint64_t>(d->in_w), static_cast<uint64_t>(d->in_h),
                              static_cast<uint64_t>(B)};
    // x0 may be a channel slice of a wider NHWC tensor: its pixels are x0_pix_stride channels apart
    const uint64_t ps0 = d->x0_pix_stride > 0 ? d->x0_pix_stride : d->c0;
    STEDM_REQUIRE(ps0 >= static_cast<uint64_t>(d->c0) && ps0 % 8 == 0, "conv_tc: bad x0 pixel stride %d", d->x0_pix_stride);
    const uint64_t str[3] = {ps0 * 2, static_cast<uint64_t>(d->in_w) * ps0 * 2,
                             static_cast<uint64_t>(d->in_h) * d->in_w * ps0 * 2};
    // stride 2: a box of 2 tw x 2 th input pixels traversed with stride 2 = the tile's tw x th sampling positions
    const uint32_t box[4] = {TC_BK, static_cast<uint32_t>(tw * cs), abox_h * cs, static_cast<uint32_t>(tb)};
    const uint32_t est[4] = {1, static_cast<uint32_t>(cs), static_cast<uint32_t>(cs), 1};
    int rc = make_tmap_bf16(&ma0, d->x0, 4, dims, str, box, est);
    if (rc) return rc;
  }
  if (d->c1 > 0) {
    const int b1 = x1b > 0 ? x1b : B;
    const uint64_t dims[4] = {static_cast<uint64_t>(d->c1), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(b1)};
    const uint64_t ps1 = d->x1_pix_stride > 0 ? d->x1_pix_stride : d->c1;   // x1 as a channel slice, like x0
    STEDM_REQUIRE(ps1 >= static_cast<uint64_t>(d->c1) && ps1 % 8 == 0, "conv_tc: bad x1 pixel stride %d", d->x1_pix_stride);
    const uint64_t str[3] = {ps1 * 2, static_cast<uint64_t>(W) * ps1 * 2, static_cast<uint64_t>(H) * W * ps1 * 2};
    const uint32_t box[4] = {TC_BK, static_cast<uint32_t>(tw), abox_h, static_cast<uint32_t>(tb)};
    int rc = make_tmap_bf16(&ma1, d->x1, 4, dims, str, box);
    if (rc) return rc;
  } else {
    ma1 = ma0;
  }
  if (d->skip_x0) {
    STEDM_REQUIRE(d->tap_mode == 0 && ctot % TC_BK == 0 && d->skip_c0 > 0 && d->skip_c0 % TC_BK == 0 &&
                      d->skip_c1 % TC_BK == 0 && (d->skip_c1 == 0 || d->skip_x1),
                  "conv_tc: fused skip input needs dense taps and channel counts that are multiples of 64 (%d, %d)",
                  d->skip_c0, d->skip_c1);
  }
  const int skip_x1b = (skip_c > 0 && d->skip_c1 > 0 && d->skip_x1_batch > 0 && d->skip_x1_batch != B) ? d->skip_x1_batch : 0;
  if (skip_x1b > 0)
    STEDM_REQUIRE((static_cast<long long>(skip_x1b) * H * W) % TC_BM == 0 || tb == 1,
                  "conv_tc: broadcast skip batch %d not tile aligned", skip_x1b);
  CUtensorMap ms0 = ma0, ms1 = ma0;
  if (skip_c > 0) {
    const uint32_t box[4] = {TC_BK, static_cast<uint32_t>(tw), static_cast<uint32_t>(th), static_cast<uint32_t>(tb)};
    {
      const uint64_t dims[4] = {static_cast<uint64_t>(d->skip_c0), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                static_cast<uint64_t>(B)};
      const uint64_t str[3] = {static_cast<uint64_t>(d->skip_c0) * 2, static_cast<uint64_t>(W) * d->skip_c0 * 2,
                               static_cast<uint64_t>(H) * W * d->skip_c0 * 2};
      int rc = make_tmap_bf16(&ms0, d->skip_x0, 4, dims, str, box);
      if (rc) return rc;
    }
    ms1 = ms0;
    if (d->skip_c1 > 0) {
      const int bs1 = skip_x1b > 0 ? skip_x1b : B;
      const uint64_t dims[4] = {static_cast<uint64_t>(d->skip_c1), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                static_cast<uint64_t>(bs1)};
      const uint64_t str[3] = {static_cast<uint64_t>(d->skip_c1) * 2, static_cast<uint64_t>(W) * d->skip_c1 * 2,
                               static_cast<uint64_t>(H) * W * d->skip_c1 * 2};
      int rc = make_tmap_bf16(&ms1, d->skip_x1, 4, dims, str, box);
      if (rc) return rc;
    }
  }
  const int bn = plan.bn, cl = plan.cl;
  const bool pair = plan.pair;
  {
    const uint64_t ktot = static_cast<uint64_t>(taps) * ctot + skip_c;
    const uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->cout)};
    const uint64_t str[1] = {ktot * 2};
    const uint32_t box[2] = {TC_BK, static_cast<uint32_t>(bn / cl)};  // multicast half / pair half / whole slab
    int rc = make_tmap_bf16(&mw, d->weight, 2, dims, str, box);
    if (rc) return rc;
  }
  TcParams p;
  p.bias = d->bias; p.emb = d->emb; p.residual = d->residual; p.out = d->out;
  p.M = static_cast<int>(M); p.H = H; p.W = W; p.HW = H * W; p.cout = d->cout;
  p.taps = taps; p.ksize = d->ksize;
  p.c0_blks = (d->c0 + TC_BK - 1) / TC_BK; p.c_blks = c_blks;
  p.x1_batch = x1b;
  p.n_tiles = (d->cout + bn - 1) / bn;
  STEDM_REQUIRE(d->cout % bn == 0 || (d->cout % 32 == 0 && d->out_nchw == 0),
                "conv_tc: cout %d with a partial last channel tile needs cout %% 32 == 0 and NHWC output", d->cout);
  p.ksplit = plan.ksplit; p.kb_per_split = plan.kb_per_split; p.m_pad = plan.m_pad;
  p.ws = static_cast<float*>(d->workspace);
  p.emb_stride = d->emb_stride; p.res_dtype = d->res_dtype; p.out_dtype = d->out_dtype;
  p.out_nchw = d->out_nchw; p.cout_store = d->cout_store > 0 ? d->cout_store : d->cout;
  p.tap_mode = d->tap_mode; p.py = d->phase >> 1; p.px = d->phase & 1;
  p.cs = cs;
  p.act = d->act;
  p.skip_blks = skip_blks; p.skip_c0_blks = skip_c > 0 ? d->skip_c0 / TC_BK : 0; p.skip_x1_batch = skip_x1b;
  p.halo = halo; p.n_t = n_t; p.na = halo_na; p.a_buf_bytes = halo_bytes; p.a_row_bytes = tw * TC_BK * 2;
  p.tile2d = L.tile2d; p.txs = W / 16; p.tps = H * W / TC_BM;
  p.res_rows = 0;
  if (d->residual != nullptr && d->res_batch > 0 && d->res_batch != B) {
    STEDM_REQUIRE(d->tap_mode == 0 && B % d->res_batch == 0, "conv_tc: residual batch %d does not divide the batch %d",
                  d->res_batch, B);
    p.res_rows = d->res_batch * H * W;
  }
  p.gn_coef = d->gn_coef; p.gn_cstride = d->gn_cstride; p.gn_c_off = d->gn_c_off; p.gn_silu = d->gn_silu;
  p.nb = 0; p.a_pool_bytes = 0; p.log2w = 0;
  while ((1 << p.log2w) < W) ++p.log2w;
  p.stats_out = nullptr; p.stats_tile_base = 0;
  if (d->stats_out != nullptr) {
    STEDM_REQUIRE(bn >= 64 && d->cout % bn == 0 && d->out_nchw == 0 && (H * W) % TC_BM == 0,
                  "conv_tc: fused GroupNorm statistics need cout %% 64 == 0, NHWC output and H*W %% 128 == 0");
    const int m_tiles = static_cast<int>((M + TC_BM - 1) / TC_BM);
    p.stats_out = d->stats_out;
    p.stats_tile_base = d->tap_mode == 1 ? d->phase * m_tiles : 0;
  }
  auto s = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256:
      if (pair) return launch_tc<256, 2, true>(ma0, ma1, mw, ms0, ms1, p, s);
      return cl == 2 ? launch_tc<256, 2>(ma0, ma1, mw, ms0, ms1, p, s) : launch_tc<256, 1>(ma0, ma1, mw, ms0, ms1, p, s);
    case 128:
      if (pair) return launch_tc<128, 2, true>(ma0, ma1, mw, ms0, ms1, p, s);
      return cl == 2 ? launch_tc<128, 2>(ma0, ma1, mw, ms0, ms1, p, s) : launch_tc<128, 1>(ma0, ma1, mw, ms0, ms1, p, s);
    case 64: return launch_tc<64, 1>(ma0, ma1, mw, ms0, ms1, p, s);
    default: return launch_tc<16, 1>(ma0, ma1, mw, ms0, ms1, p, s);
  }
}
