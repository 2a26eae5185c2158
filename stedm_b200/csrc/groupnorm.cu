// K7: GroupNorm(32 groups) over NHWC tensors, with the channel concat of skip tensors and the SiLU that follow
// it in the reference folded in.  HBM-bound: the statistics pass reads the input once, the apply pass reads it
// once and writes the (usually bf16) operand of the consumer convolution.  fp32 statistics per thread, double
// precision across threads/blocks (GroupNorm32 computes in fp32: ldm/modules/diffusionmodules/util.py:214-216);
// no atomics anywhere, so results are deterministic and independent of the batch a sample is launched in.
//
// Layout: one "item" = 8 consecutive channels of one pixel (16 B of bf16 / 32 B of fp32).  A 256-thread block
// owns one sample and a slice of its pixels; thread -> (pixel lane, channel item) so that a warp touches
// consecutive 16 B items of the same pixel rows (fully coalesced).
#include "../../include/stedm_b200.h"
#include <stdlib.h>

#include "common.cuh"

using namespace stedm;

namespace {

constexpr int GN_GROUPS = 32;
constexpr int GN_THREADS = 256;
constexpr int GN_MLP = 4;  // independent loads in flight per thread

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]);
  u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]);
  u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// Pointer to the 8-channel item `item` of pixel `pix` of sample b in the virtual concat [x0 | x1].
template <typename T>
__device__ __forceinline__ const T* item_ptr(const T* x0, const T* x1, int b, int b1, size_t pix, int hw, int c0,
                                             int c1, int item) {
  const int c = item * 8;
  if (c < c0) return x0 + (static_cast<size_t>(b) * hw + pix) * c0 + c;
  return x1 + (static_cast<size_t>(b1) * hw + pix) * c1 + (c - c0);
}

// Deterministic, batch-size-independent reduction: block (chunk, b) reduces a fixed pixel range; per-thread fp32
// partials go to shared memory WITHOUT atomics ([lane][channel]), are folded over lanes and the group's channels
// in a fixed order in double precision, and written to partials[b][chunk][group][{sum, sumsq}].  The apply kernel
// folds the chunks in index order.  Hence a sample's statistics do not depend on which other samples share the
// launch (a rank-sharded run is bit-identical to the single-GPU run) nor on scheduling.
template <typename T>
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const T* __restrict__ x0, const T* __restrict__ x1,
                                                              int x1_batch, int hw, int c0, int c1, int pix_per_block,
                                                              double* __restrict__ partials) {
  extern __shared__ float s_acc[];  // [2][lanes][C]
  const int C = c0 + c1, items = C / 8;
  const int b = blockIdx.y, b1 = (x1_batch > 0) ? (b % x1_batch) : b;
  const int tpi = min(items, GN_THREADS);       // threads along the item axis
  const int lanes = GN_THREADS / tpi;           // pixel lanes
  const int lane = threadIdx.x / tpi, it0 = threadIdx.x % tpi;
  const int p_begin = blockIdx.x * pix_per_block, p_end = min(hw, p_begin + pix_per_block);
  float* s_sum = s_acc;
  float* s_sq = s_acc + static_cast<size_t>(lanes) * C;
  if (lane < lanes) {
    for (int item = it0; item < items; item += tpi) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
      const T* src = item_ptr<T>(x0, x1, b, b1, 0, hw, c0, c1, item);
      const size_t stride = (item * 8 < c0) ? c0 : c1;
      // batches of GN_MLP independent 16/32-byte loads are issued before any is consumed (memory-level parallelism)
      int p = p_begin + lane;
      for (; p + (GN_MLP - 1) * lanes < p_end; p += GN_MLP * lanes) {
        float v[GN_MLP][8];
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u) load8<T>(src + static_cast<size_t>(p + u * lanes) * stride, v[u]);
#pragma unroll
        for (int u = 0; u < GN_MLP; ++u)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += v[u][j];
            q[j] += v[u][j] * v[u][j];
          }
      }
      for (; p < p_end; p += lanes) {
        float v[8];
        load8<T>(src + static_cast<size_t>(p) * stride, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[j];
          q[j] += v[j] * v[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_sum[lane * C + item * 8 + j] = s[j];
        s_sq[lane * C + item * 8 + j] = q[j];
      }
    }
  }
  __syncthreads();
  {
    // 64 outputs (32 groups x {sum, sumsq}), 4 threads each: fixed strided partition + fixed shuffle tree =
    // deterministic, and 4x shorter than one thread per output
    const int o = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const int g = o % GN_GROUPS, which = o / GN_GROUPS;
    const int cpg = C / GN_GROUPS;
    const float* src = (which ? s_sq : s_sum) + g * cpg;
    double acc = 0.0;
    const int n = lanes * cpg;
    for (int i = sub; i < n; i += 4) {
      const int l = i / cpg, c = i - l * cpg;
      acc += static_cast<double>(src[l * C + c]);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0) partials[((static_cast<size_t>(b) * gridDim.x + blockIdx.x) * GN_GROUPS + g) * 2 + which] = acc;
  }
}

template <typename TI, typename TO, bool kPrecise>
__global__ void __launch_bounds__(GN_THREADS, kPrecise ? 1 : 6) gn_apply_kernel(const TI* __restrict__ x0, const TI* __restrict__ x1,
                                                              int x1_batch, int hw, int c0, int c1, int pix_per_block,
                                                              const double* __restrict__ partials, int n_chunks,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, int apply_silu,
                                                              TO* __restrict__ out, int tpi_lo, int split_c,
                                                              int batch_lo, TO* __restrict__ out_hi, int tpi_hi) {
  extern __shared__ float s_ab[];  // scale[C], shift[C]
  __shared__ double s_tot[2 * GN_GROUPS];
  __shared__ float s_mr[2 * GN_GROUPS];  // per group: mean, rstd
  const int C = c0 + c1, cpg = C / GN_GROUPS;
  // split mode (stedm_gn_apply_split): rows y < batch_lo of the grid write channels [0, split_c) of sample y to `out`,
  // rows above write channels [split_c, C) of sample y - batch_lo to `out_hi`; both outputs are dense
  int b = blockIdx.y, c_begin = 0, c_end = C;
  if (split_c > 0) {
    if (b < batch_lo) {
      c_end = split_c;
    } else {
      b -= batch_lo;
      c_begin = split_c;
      out = out_hi;
    }
  }
  const int items = (c_end - c_begin) / 8, item_base = c_begin / 8, c_out = c_end - c_begin;
  const int b1 = (x1_batch > 0) ? (b % x1_batch) : b;
  {
    // fold the per-chunk partials: 64 outputs x 4 threads, fixed strided partition + fixed shuffle tree
    // (deterministic); the loads of one thread are independent, so they are issued in batches
    const int o = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const double* src = partials + static_cast<size_t>(b) * n_chunks * GN_GROUPS * 2 + o;
    double acc = 0.0;
#pragma unroll 8
    for (int k = sub; k < n_chunks; k += 4) acc += src[static_cast<size_t>(k) * GN_GROUPS * 2];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0) s_tot[o] = acc;  // index = g*2 + which
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {  // the only double-precision sqrt/div: 32 per block
    const double n = static_cast<double>(hw) * cpg;
    const double mean = s_tot[threadIdx.x * 2] / n;
    double var = s_tot[threadIdx.x * 2 + 1] / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mr[threadIdx.x * 2] = static_cast<float>(mean);
    s_mr[threadIdx.x * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int c = c_begin + threadIdx.x; c < c_end; c += GN_THREADS) {
    const int g = c / cpg;
    const float a = s_mr[g * 2 + 1] * gamma[c];
    s_ab[c] = a;
    s_ab[C + c] = fmaf(-s_mr[g * 2], a, beta[c]);
  }
  __syncthreads();
  // thread -> (pixel lane, channel item): no integer division in the streaming loop, the item's scale/shift
  // live in registers, and the unrolled pixel loop keeps several independent 16 B loads in flight per thread
  const int p_begin = blockIdx.x * pix_per_block, p_end = min(hw, p_begin + pix_per_block);
  const int tpi = (split_c > 0 && c_begin > 0) ? tpi_hi : tpi_lo;   // threads along the item axis (apply_threads_per_item)
  const int lanes = GN_THREADS / tpi;
  const int lane = threadIdx.x / tpi, it0 = threadIdx.x % tpi;
  if (lane >= lanes) return;
  for (int item = it0; item < items; item += tpi) {
    const int gi = item_base + item;          // item of the concat's channel axis
    float a[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = s_ab[gi * 8 + j];
      sh[j] = s_ab[C + gi * 8 + j];
    }
    const TI* src = item_ptr<TI>(x0, x1, b, b1, 0, hw, c0, c1, gi);
    const size_t src_stride = (gi * 8 < c0) ? c0 : c1;
    TO* dst = out + static_cast<size_t>(b) * hw * c_out + item * 8;
    int p = p_begin + lane;
    for (; p + (GN_MLP - 1) * lanes < p_end; p += GN_MLP * lanes) {
      float v[GN_MLP][8];
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) load8<TI>(src + static_cast<size_t>(p + u * lanes) * src_stride, v[u]);
#pragma unroll
      for (int u = 0; u < GN_MLP; ++u) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float y = fmaf(v[u][j], a[j], sh[j]);
          if (apply_silu) y = kPrecise ? silu_precise(y) : silu_f(y);
          v[u][j] = y;
        }
        store8<TO>(dst + static_cast<size_t>(p + u * lanes) * c_out, v[u]);
      }
    }
    for (; p < p_end; p += lanes) {
      float v[8];
      load8<TI>(src + static_cast<size_t>(p) * src_stride, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(v[j], a[j], sh[j]);
        if (apply_silu) y = kPrecise ? silu_precise(y) : silu_f(y);
        v[j] = y;
      }
      store8<TO>(dst + static_cast<size_t>(p) * c_out, v);
    }
  }
}

// bf16 -> bf16 throughput variant of the apply pass: the input is staged by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier) through a 4 x 16 KB shared-memory ring, so ~48 KB per block (x3 blocks per SM) are in
// flight without costing a register, instead of 4 x 16 B per thread; the two sources of a concat are streamed one
// after the other (each is contiguous per pixel range).  Same arithmetic as gn_apply_kernel<bf16, bf16, false>.
constexpr int GB_STAGES = 4;
constexpr int GB_STAGE_BYTES = 16384;

__global__ void __launch_bounds__(GN_THREADS) gn_apply_bulk_kernel(
    const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1, int x1_batch, int hw, int c0, int c1,
    int pix_per_block, const double* __restrict__ partials, int n_chunks, const float* __restrict__ gamma,
    const float* __restrict__ beta, float eps, int apply_silu, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(128) uint8_t s_raw[];
  uint8_t* s_data = s_raw;                                                 // [GB_STAGES][GB_STAGE_BYTES]
  float* s_ab = reinterpret_cast<float*>(s_raw + GB_STAGES * GB_STAGE_BYTES);  // scale[C], shift[C]
  __shared__ double s_tot[2 * GN_GROUPS];
  __shared__ float s_mr[2 * GN_GROUPS];
  __shared__ uint64_t s_full[GB_STAGES];
  const int C = c0 + c1, cpg = C / GN_GROUPS;
  const int b = blockIdx.y, b1 = (x1_batch > 0) ? (b % x1_batch) : b;
  if (threadIdx.x == 0) {
    for (int i = 0; i < GB_STAGES; ++i) mbar_init(&s_full[i], 1);
    fence_barrier_init();
  }
  {
    const int o = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const double* src = partials + static_cast<size_t>(b) * n_chunks * GN_GROUPS * 2 + o;
    double acc = 0.0;
#pragma unroll 8
    for (int k = sub; k < n_chunks; k += 4) acc += src[static_cast<size_t>(k) * GN_GROUPS * 2];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0) s_tot[o] = acc;
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    const double n = static_cast<double>(hw) * cpg;
    const double mean = s_tot[threadIdx.x * 2] / n;
    double var = s_tot[threadIdx.x * 2 + 1] / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mr[threadIdx.x * 2] = static_cast<float>(mean);
    s_mr[threadIdx.x * 2 + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    const int g = c / cpg;
    const float a = s_mr[g * 2 + 1] * gamma[c];
    s_ab[c] = a;
    s_ab[C + c] = fmaf(-s_mr[g * 2], a, beta[c]);
  }
  __syncthreads();

  const int p_begin = blockIdx.x * pix_per_block, p_end = min(hw, p_begin + pix_per_block);
  uint32_t kk = 0;  // chunk counter across both sources: ring slot = kk % GB_STAGES, parity = (kk / GB_STAGES) & 1
#pragma unroll 1
  for (int s = 0; s < 2; ++s) {
    const int cs = s == 0 ? c0 : c1;
    if (cs == 0) continue;
    const int coff = s == 0 ? 0 : c0;
    const __nv_bfloat16* base = (s == 0 ? x0 + static_cast<size_t>(b) * hw * c0 : x1 + static_cast<size_t>(b1) * hw * c1);
    const int ipp = cs / 8;                     // 16-byte items per pixel (<= 256, checked by the host)
    const int lanes = GN_THREADS / ipp;
    const int lane = threadIdx.x / ipp, ci = threadIdx.x % ipp;
    const bool active = lane < lanes;
    const int row_bytes = cs * 2;
    const int pps = max(lanes, (GB_STAGE_BYTES / row_bytes) / lanes * lanes);  // pixels per ring stage
    const int n_ch = (p_end - p_begin + pps - 1) / pps;
    float a[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = s_ab[coff + ci * 8 + j];
      sh[j] = s_ab[C + coff + ci * 8 + j];
    }
    auto issue = [&](int k, uint32_t slot) {
      const int ps = p_begin + k * pps, np = min(pps, p_end - ps);
      const uint32_t bytes = static_cast<uint32_t>(np) * row_bytes;
      mbar_arrive_expect_tx(&s_full[slot], bytes);
      bulk_load_1d(s_data + slot * GB_STAGE_BYTES, base + static_cast<size_t>(ps) * cs, bytes, &s_full[slot]);
    };
    if (threadIdx.x == 0)
      for (int k = 0; k < min(GB_STAGES, n_ch); ++k) issue(k, (kk + k) % GB_STAGES);
    for (int k = 0; k < n_ch; ++k, ++kk) {
      const uint32_t slot = kk % GB_STAGES;
      mbar_wait(&s_full[slot], (kk / GB_STAGES) & 1);
      const int ps = p_begin + k * pps, np = min(pps, p_end - ps);
      if (active) {
        const uint4* sd = reinterpret_cast<const uint4*>(s_data + slot * GB_STAGE_BYTES);
        __nv_bfloat16* dst = out + (static_cast<size_t>(b) * hw + ps) * C + coff + ci * 8;
#pragma unroll 2
        for (int q = lane; q < np; q += lanes) {
          const uint4 u = sd[q * ipp + ci];
          float v[8];
          float2 f;
          f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
          f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
          f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
          f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y = fmaf(v[j], a[j], sh[j]);
            v[j] = apply_silu ? silu_f(y) : y;
          }
          store8<__nv_bfloat16>(dst + static_cast<size_t>(q) * C, v);
        }
      }
      __syncthreads();  // every thread is done reading this slot: it may be refilled
      if (threadIdx.x == 0 && k + GB_STAGES < n_ch) issue(k + GB_STAGES, slot);
    }
  }
}

// Pixel range per statistics block: a function of (hw, C) ONLY (never of the batch), at most 128 chunks per sample.
int stats_pix_per_block(int hw, int C) {
  static const int kStatsBytes = [] { const char* e = getenv("STEDM_GN_STATS_ELEMS"); return e ? atoi(e) : 65536; }();
  const int by_bytes = max(2, kStatsBytes / C);  // elements per block (~128 KB of bf16 by default)
  const int by_count = (hw + 127) / 128;
  return min(hw, max(by_bytes, by_count));
}

// Fold the per-(128-pixel tile, channel) statistics that stedm_conv_tc's epilogue wrote for the producer(s) of a
// GroupNorm input into the per-sample, per-group (sum, sumsq) the apply kernel consumes — the statistics pass over
// the activation itself disappears.  One block per sample; 64 outputs x 4 threads, fixed partition + fixed shuffle
// tree in double precision (deterministic, independent of the batch).
struct FoldSrc {
  const float* tiles;   // [reps][rep_stride tiles][c][2]
  int c, reps, tps, batch;
  long long rep_stride;  // in tiles
};

// One block per (sample, group): thread <-> (tile lane, channel of the group): (sum, sumsq) over the sample's tiles with
// float2 loads accumulated in double in tile order, then a fixed-order fold over tile lanes and the group's channels.
// (History: 16 lanes per output walking tiles with 4-byte loads 1 KB apart took 18 us per launch; one block per SAMPLE
// with coalesced loads 8 us at batch 64+ but 25 us at batch 1-2 — two blocks on the whole GPU; per (sample, group) blocks
// keep the launch short at every batch.)
constexpr int FOLD_THREADS = 128;
__global__ void __launch_bounds__(FOLD_THREADS) gn_fold_tiles_kernel(FoldSrc s0, FoldSrc s1, double* __restrict__ out,
                                                                     const float* __restrict__ gamma,
                                                                     const float* __restrict__ beta, float eps, int hw,
                                                                     float2* __restrict__ coef) {
  __shared__ double s_part[FOLD_THREADS][2];
  __shared__ float s_mr[2];
  const int b = blockIdx.x, g = blockIdx.y;
  const int C = s0.c + s1.c, cpg = C / GN_GROUPS;
  const int cl = cpg < FOLD_THREADS ? cpg : FOLD_THREADS;   // channel lanes (cpg <= 128 on the path: C <= 4096)
  const int TL = FOLD_THREADS / cl;                         // tile lanes
  const int tl = threadIdx.x / cl, c_lane = threadIdx.x - tl * cl;
  double a0 = 0.0, a1 = 0.0;
  if (tl < TL) {
    for (int cc = c_lane; cc < cpg; cc += cl) {
      const int ch = g * cpg + cc;
      const FoldSrc& s = (ch < s0.c) ? s0 : s1;
      const int lc = (ch < s0.c) ? ch : ch - s0.c;
      const int bs = b % s.batch;
      for (int r = 0; r < s.reps; ++r) {
        const float* base = s.tiles + ((static_cast<size_t>(r) * s.rep_stride + static_cast<size_t>(bs) * s.tps) * s.c + lc) * 2;
#pragma unroll 4
        for (int j = tl; j < s.tps; j += TL) {
          const float2 v = *reinterpret_cast<const float2*>(base + static_cast<size_t>(j) * s.c * 2);
          a0 += static_cast<double>(v.x);
          a1 += static_cast<double>(v.y);
        }
      }
    }
  }
  s_part[threadIdx.x][0] = a0;
  s_part[threadIdx.x][1] = a1;
  __syncthreads();
  if (threadIdx.x < 2) {   // fixed order over the block's partials: deterministic
    double acc = 0.0;
    for (int i = 0; i < FOLD_THREADS; ++i) acc += s_part[i][threadIdx.x];
    s_part[0][threadIdx.x] = acc;   // (thread t only reads column t: no hazard with the other thread's store)
    if (out != nullptr) out[(static_cast<size_t>(b) * GN_GROUPS + g) * 2 + threadIdx.x] = acc;
  }
  if (coef == nullptr) return;
  // Per-(sample, channel) scale / shift for consumers that apply the normalisation in their own operand path
  // (stedm_conv_desc.gn_coef): the SAME arithmetic, in the same precision, as gn_apply_kernel's prologue, so a fused
  // consumer is bit-identical to gn_apply + plain consumer.
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = static_cast<double>(hw) * cpg;
    const double mean = s_part[0][0] / n;
    double var = s_part[0][1] / n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mr[0] = static_cast<float>(mean);
    s_mr[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int cc = threadIdx.x; cc < cpg; cc += FOLD_THREADS) {
    const int c = g * cpg + cc;
    const float a = s_mr[1] * gamma[c];
    coef[static_cast<size_t>(b) * C + c] = make_float2(a, fmaf(-s_mr[0], a, beta[c]));
  }
}

// Threads along the item (8-channel) axis of an apply block; the other GN_THREADS / tpi are pixel lanes and a thread loops
// over items it0, it0 + tpi, ...  The split that keeps the most threads busy: e.g. 192 items -> 64 threads x 4 lanes x 3
// items each instead of 192 threads x 1 lane with a quarter of the block idle — as long as every lane keeps GN_MLP pixels
// for the unrolled loop.
int apply_threads_per_item(int items, int pix_per_block) {
  int tpi = items < GN_THREADS ? items : GN_THREADS, best = 0;
  const int max_lanes = max(GN_THREADS / tpi, pix_per_block / GN_MLP);
  for (int k = 1; k <= 8; ++k) {
    const int t = (items + k - 1) / k;
    if (t > GN_THREADS || GN_THREADS / t > max_lanes) continue;
    const int score = items * (GN_THREADS / t) / k;   // busy thread-iterations per unit of block time
    if (score > best) { best = score; tpi = t; }
  }
  return tpi;
}

int apply_pix_per_block(int batch, int hw, int C) {
  // >= 64 KB of input per block (amortises the per-block scale/shift prologue) unless that leaves the GPU short of
  // ~4 blocks per SM, then smaller down to 8 KB
  static const int kApplyElems = [] { const char* e = getenv("STEDM_GN_APPLY_ELEMS"); return e ? atoi(e) : 32768; }();
  int ppb = max(1, kApplyElems / C);
  const long long blocks = static_cast<long long>(batch) * ((hw + ppb - 1) / ppb);
  if (blocks < 148 * 4) {
    const int blocks_per_sample = max(1, (148 * 4 + batch - 1) / batch);
    ppb = max(max(1, 4096 / C), (hw + blocks_per_sample - 1) / blocks_per_sample);
  }
  return min(ppb, hw);
}

}  // namespace

extern "C" int stedm_gn_num_chunks(int hw, int channels) {
  if (hw <= 0 || channels <= 0) return ERR_ARG;
  const int ppb = stats_pix_per_block(hw, channels);
  return (hw + ppb - 1) / ppb;
}

extern "C" int stedm_gn_stats(const void* x0, const void* x1, int in_dtype, int batch, int x1_batch, int hw, int c0,
                              int c1, double* partials, void* stream) {
  const int C = c0 + c1;
  STEDM_REQUIRE(x0 && partials && (c1 == 0 || x1), "gn_stats: null pointer");
  STEDM_REQUIRE(batch > 0 && hw > 0 && c0 > 0 && c0 % 8 == 0 && c1 % 8 == 0 && C % GN_GROUPS == 0,
                "gn_stats: channels (%d + %d) must be multiples of 8 and sum to a multiple of 32", c0, c1);
  STEDM_REQUIRE(C <= 4096, "gn_stats: too many channels");
  const int ppb = stats_pix_per_block(hw, C);
  dim3 grid((hw + ppb - 1) / ppb, batch);
  const int items = C / 8, tpi = items < GN_THREADS ? items : GN_THREADS, lanes = GN_THREADS / tpi;
  const size_t smem = static_cast<size_t>(2) * lanes * C * sizeof(float);
  auto s = static_cast<cudaStream_t>(stream);
  if (in_dtype == DT_BF16)
    gn_stats_kernel<__nv_bfloat16><<<grid, GN_THREADS, smem, s>>>(static_cast<const __nv_bfloat16*>(x0),
                                                                 static_cast<const __nv_bfloat16*>(x1), x1_batch, hw,
                                                                 c0, c1, ppb, partials);
  else
    gn_stats_kernel<float><<<grid, GN_THREADS, smem, s>>>(static_cast<const float*>(x0), static_cast<const float*>(x1),
                                                         x1_batch, hw, c0, c1, ppb, partials);
  return check_launch("gn_stats");
}

extern "C" int stedm_gn_fold_tiles(const float* tiles0, int c0, int reps0, long long rep_stride0, int tps0, int batch0,
                                   const float* tiles1, int c1, int reps1, long long rep_stride1, int tps1, int batch1,
                                   int batch, double* out, const float* gamma, const float* beta, float eps, int hw,
                                   float* coef, void* stream) {
  STEDM_REQUIRE(tiles0 && (out || coef) && (c1 == 0 || tiles1), "gn_fold_tiles: null pointer");
  STEDM_REQUIRE(coef == nullptr || (gamma && beta && hw > 0), "gn_fold_tiles: coefficients need gamma, beta and hw");
  STEDM_REQUIRE(batch > 0 && c0 > 0 && (c0 + c1) % GN_GROUPS == 0 && reps0 > 0 && tps0 > 0 && batch0 > 0 &&
                    (c1 == 0 || (reps1 > 0 && tps1 > 0 && batch1 > 0)),
                "gn_fold_tiles: bad shape");
  FoldSrc s0{tiles0, c0, reps0, tps0, batch0, rep_stride0};
  FoldSrc s1{tiles1, c1, c1 ? reps1 : 1, c1 ? tps1 : 1, c1 ? batch1 : 1, rep_stride1};
  gn_fold_tiles_kernel<<<dim3(batch, GN_GROUPS), FOLD_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      s0, s1, out, gamma, beta, eps, hw, reinterpret_cast<float2*>(coef));
  return check_launch("gn_fold_tiles");
}

extern "C" int stedm_gn_apply(const void* x0, const void* x1, int in_dtype, int batch, int x1_batch, int hw, int c0,
                              int c1, const double* partials, int n_chunks_in, const float* gamma, const float* beta,
                              float eps, int apply_silu, void* out, int out_dtype, void* stream) {
  const int C = c0 + c1;
  STEDM_REQUIRE(x0 && partials && gamma && beta && out && (c1 == 0 || x1), "gn_apply: null pointer");
  STEDM_REQUIRE(batch > 0 && hw > 0 && c0 > 0 && c0 % 8 == 0 && c1 % 8 == 0 && C % GN_GROUPS == 0 && C <= 4096,
                "gn_apply: bad channel counts (%d + %d)", c0, c1);
  const int sppb = stats_pix_per_block(hw, C);
  const int n_chunks = n_chunks_in > 0 ? n_chunks_in : (hw + sppb - 1) / sppb;
  const int ppb = apply_pix_per_block(batch, hw, C);
  dim3 grid((hw + ppb - 1) / ppb, batch);
  const size_t smem = static_cast<size_t>(2) * C * sizeof(float);
  const int tpi = apply_threads_per_item(C / 8, ppb < hw ? ppb : hw);
  auto s = static_cast<cudaStream_t>(stream);
#define LAUNCH(TI, TO, PREC)                                                                                       \
  gn_apply_kernel<TI, TO, PREC><<<grid, GN_THREADS, smem, s>>>(static_cast<const TI*>(x0), static_cast<const TI*>(x1), \
                                                              x1_batch, hw, c0, c1, ppb, partials, n_chunks, gamma, \
                                                              beta, eps, apply_silu, static_cast<TO*>(out), tpi, 0, 0, nullptr, 0)
  static const bool kBulk = [] { const char* e = getenv("STEDM_GN_BULK"); return !(e && e[0] == '0'); }();
  if (in_dtype == DT_BF16 && out_dtype == DT_BF16 && kBulk && c0 <= 2048 && c1 == 0) {  // measured: +10-17 % for one source, slower for a concat (two short pipelines)
    // throughput path: TMA-bulk staged input (needs 16-byte items per pixel <= 256 per source)
    const int ppb_bulk = max(ppb, 4 * (GB_STAGE_BYTES / (C * 2) > 0 ? GB_STAGE_BYTES / (C * 2) : 1));
    dim3 grid_b((hw + ppb_bulk - 1) / ppb_bulk, batch);
    const size_t smem_b = static_cast<size_t>(GB_STAGES) * GB_STAGE_BYTES + static_cast<size_t>(2) * C * sizeof(float);
    static DeviceOnce configured;
    if (configured.needed()) {
      cudaFuncSetAttribute(gn_apply_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      configured.done();
    }
    gn_apply_bulk_kernel<<<grid_b, GN_THREADS, smem_b, s>>>(
        static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), x1_batch, hw, c0, c1, ppb_bulk,
        partials, n_chunks, gamma, beta, eps, apply_silu, static_cast<__nv_bfloat16*>(out));
  } else if (in_dtype == DT_BF16 && out_dtype == DT_BF16)
    LAUNCH(__nv_bfloat16, __nv_bfloat16, false);
  else if (in_dtype == DT_F32 && out_dtype == DT_BF16)
    LAUNCH(float, __nv_bfloat16, false);
  else if (in_dtype == DT_F32 && out_dtype == DT_F32)
    LAUNCH(float, float, true);
  else if (in_dtype == DT_BF16 && out_dtype == DT_F32)
    LAUNCH(__nv_bfloat16, float, true);
  else {
    set_error("gn_apply: unsupported dtype combination");
    return ERR_UNSUPPORTED;
  }
#undef LAUNCH
  return check_launch("gn_apply");
}

// GroupNorm (+ SiLU) of [x0 | x1] where x1 holds `x1_batch` < `batch` DISTINCT samples broadcast as b % x1_batch (guided
// sampling: the encoder skips are shared by the cond / uncond halves).  Channels [split_c, C) lie in groups made of x1
// channels only, so their normalised values repeat with period x1_batch: they are written ONCE per distinct sample to
// `out_hi` [x1_batch][hw][C - split_c]; channels [0, split_c) go to `out_lo` [batch][hw][split_c].  Same arithmetic per
// element as stedm_gn_apply (whose output is [out_lo | out_hi broadcast]); one launch.
extern "C" int stedm_gn_apply_split(const void* x0, const void* x1, int batch, int x1_batch, int hw, int c0, int c1,
                                    const double* partials, int n_chunks_in, const float* gamma, const float* beta,
                                    float eps, int apply_silu, int split_c, void* out_lo, void* out_hi, void* stream) {
  const int C = c0 + c1;
  STEDM_REQUIRE(x0 && x1 && partials && gamma && beta && out_lo && out_hi, "gn_apply_split: null pointer");
  STEDM_REQUIRE(batch > 0 && hw > 0 && c0 > 0 && c1 > 0 && c0 % 8 == 0 && c1 % 8 == 0 && C % GN_GROUPS == 0 && C <= 4096,
                "gn_apply_split: bad channel counts (%d + %d)", c0, c1);
  STEDM_REQUIRE(x1_batch > 0 && x1_batch < batch && batch % x1_batch == 0, "gn_apply_split: x1_batch %d must divide batch %d",
                x1_batch, batch);
  const int cpg = C / GN_GROUPS;
  STEDM_REQUIRE(split_c % 8 == 0 && split_c < C && split_c >= (c0 + cpg - 1) / cpg * cpg,
                "gn_apply_split: channels from %d on must lie in groups without x0 channels (c0 %d, %d per group)", split_c,
                c0, cpg);
  const int sppb = stats_pix_per_block(hw, C);
  const int n_chunks = n_chunks_in > 0 ? n_chunks_in : (hw + sppb - 1) / sppb;
  const int ppb = apply_pix_per_block(batch, hw, split_c);   // sized by the bytes a block of the wide part streams
  dim3 grid((hw + ppb - 1) / ppb, batch + x1_batch);
  gn_apply_kernel<__nv_bfloat16, __nv_bfloat16, false><<<grid, GN_THREADS, static_cast<size_t>(2) * C * sizeof(float),
                                                         static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x0), static_cast<const __nv_bfloat16*>(x1), x1_batch, hw, c0, c1, ppb, partials,
      n_chunks, gamma, beta, eps, apply_silu, static_cast<__nv_bfloat16*>(out_lo),
      apply_threads_per_item(split_c / 8, ppb < hw ? ppb : hw), split_c, batch, static_cast<__nv_bfloat16*>(out_hi),
      apply_threads_per_item((C - split_c) / 8, ppb < hw ? ppb : hw));
  return check_launch("gn_apply_split");
}
