// Style encoder (torchvision swin_v2_t behind networks/s_zss_dm.py:19-20 and networks/agg_blocks.py:66-75): the
// kernels around the token-major linear layers, which run on the tcgen05 implicit-GEMM kernel (conv_tc.cu, ksize 1).
//
//   patch_embed_ln     Conv2d(3, E, 4, stride 4) + LayerNorm on NHWC images            (features.0)
//   layernorm          y = [residual +] LN(x) * gamma + beta  (Swin-V2 res-post-norm: x = x + norm(f(x)))
//   window_attention   shifted-window cosine attention with the continuous position bias (ShiftedWindowAttentionV2)
//   patch_merge_gather 2x2 space-to-depth in torchvision's x0|x1|x2|x3 order            (PatchMergingV2)
//   ln_meanpool        final LayerNorm + AdaptiveAvgPool2d(1)                             (norm, permute, avgpool)
//   set_reduce         mean / max over the n style images of a sample                     (Agg_Mean / Agg_Max)
//
// All of these are HBM-bound (or, for the 64-token windows, CUDA-core bound: 64x64x32 problems are far below a UMMA
// tile); reductions use warp shuffles in a fixed order, no atomics, so results do not depend on the batch.
#include "../../include/stedm_b200.h"
#include <stdlib.h>

#include "common.cuh"

using namespace stedm;

namespace {

// ------------------------------------------------------------------------------------------ 8-wide vector access
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
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
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ------------------------------------------------------------------------------------------ patch embedding
// One LANE per token: the lane keeps its 4x4x3 patch (48 floats, 12 16-byte loads: consecutive lanes = consecutive
// 48-byte runs of an image row) and all E accumulators in registers; the [48][E] weight sits in shared memory and is
// read as broadcast float4s (one LDS.128 per 4 FMAs), and the LayerNorm over E is thread-local.  (The first version
// used a warp per token with a lane per channel: 7 instructions per FMA-triple and a shuffle tree per token made it
// issue-bound at 15 % of the HBM roofline.)
template <int E>
__global__ void __launch_bounds__(128) patch_embed_ln_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             float* __restrict__ out_f32,
                                                             __nv_bfloat16* __restrict__ out_bf16, long long tokens,
                                                             int P) {
  constexpr int K = 48;
  __shared__ __align__(16) float ws[K * E];
  __shared__ __align__(16) float s_bias[E], s_gamma[E], s_beta[E];
  for (int i = threadIdx.x; i < K * E; i += blockDim.x) ws[i] = w[i];
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    s_bias[i] = bias ? bias[i] : 0.f;
    s_gamma[i] = gamma[i];
    s_beta[i] = beta[i];
  }
  __syncthreads();
  const int T = P / 4;
  for (long long tok = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; tok < tokens;
       tok += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tx = static_cast<int>(tok % T), ty = static_cast<int>((tok / T) % T);
    const long long b = tok / (static_cast<long long>(T) * T);
    const float4* src = reinterpret_cast<const float4*>(img + ((b * P + 4 * ty) * P + 4 * tx) * 3);
    float x[K];
#pragma unroll
    for (int dy = 0; dy < 4; ++dy)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4 v = __ldg(src + static_cast<long long>(dy) * (P * 3 / 4) + q);
        x[dy * 12 + q * 4] = v.x; x[dy * 12 + q * 4 + 1] = v.y; x[dy * 12 + q * 4 + 2] = v.z; x[dy * 12 + q * 4 + 3] = v.w;
      }
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; e += 4) {
      const float4 bv = *reinterpret_cast<const float4*>(&s_bias[e]);
      acc[e] = bv.x; acc[e + 1] = bv.y; acc[e + 2] = bv.z; acc[e + 3] = bv.w;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int e = 0; e < E; e += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(&ws[k * E + e]);
        acc[e] = fmaf(x[k], wv.x, acc[e]);
        acc[e + 1] = fmaf(x[k], wv.y, acc[e + 1]);
        acc[e + 2] = fmaf(x[k], wv.z, acc[e + 2]);
        acc[e + 3] = fmaf(x[k], wv.w, acc[e + 3]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) s += acc[e];
    const float mean = s * (1.0f / E);
    float q2 = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float d = acc[e] - mean;
      q2 = fmaf(d, d, q2);
    }
    const float rstd = 1.0f / sqrtf(q2 * (1.0f / E) + eps);
#pragma unroll
    for (int e = 0; e < E; e += 8) {
      float y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = (acc[e + j] - mean) * rstd * s_gamma[e + j] + s_beta[e + j];
      if (out_f32) store8(out_f32 + tok * E + e, y);
      if (out_bf16) store8(out_bf16 + tok * E + e, y);
    }
  }
}

// ------------------------------------------------------------------------------------------ row LayerNorm
// A row of C channels (C % 8 == 0, C <= 2048) is owned by G lanes (power of two <= 32); a lane holds up to PER (4 or 8)
// 8-channel items (item j of the row -> lane j % G), so every global access is a 16/32-byte vector.
__device__ __forceinline__ float group_sum(float v, int G) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename TI, int PER>
__global__ void __launch_bounds__(256) layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ residual,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float* __restrict__ out_f32,
                                                        __nv_bfloat16* __restrict__ out_bf16, long long rows, int C,
                                                        int G) {
  const int lane = threadIdx.x & 31, sub = lane % G;
  const int rows_per_warp = 32 / G;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long row = warp_global * rows_per_warp + lane / G;
  const bool live = row < rows;
  const int items = C / 8;
  float v[PER][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int it = sub + i * G;
    if (live && it < items) {
      load8<TI>(x + row * C + it * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  const float mean = group_sum(s, G) / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    if (sub + i * G < items) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = 1.0f / sqrtf(group_sum(q, G) / static_cast<float>(C) + eps);
  if (!live) return;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int it = sub + i * G;
    if (it >= items) continue;
    float g[8], b[8], y[8];
    load8<float>(gamma + it * 8, g);
    load8<float>(beta + it * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = (v[i][j] - mean) * rstd * g[j] + b[j];
    if (residual) {
      float r[8];
      load8<float>(residual + row * C + it * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] += r[j];
    }
    if (out_f32) store8(out_f32 + row * C + it * 8, y);
    if (out_bf16) store8(out_bf16 + row * C + it * 8, y);
  }
}

// ------------------------------------------------------------------------------------------ window attention
// Block = one (window, head); thread i = query token i of the 8x8 window (64 threads).  K and V of the window are
// staged in shared memory as fp32 (K already L2-normalised); the thread keeps its normalised query and the 64 logits
// of its row in registers, so the softmax needs no communication and every shared-memory read is a broadcast.
// Cyclic shift and its reverse are index arithmetic: window cell (ys, xs) of the rolled map is token
// ((ys + shift) % H, (xs + shift) % W) of the stored map; the region mask (-100 between different regions,
// swin_transformer.py shifted_window_attention) is recomputed from (ys, xs).
constexpr int WA_TOK = 64, WA_D = 32, WA_LD = 36;  // row stride 36 floats: 16-byte aligned rows

template <typename T>
__device__ __forceinline__ void load_row32(const T* p, float (&v)[WA_D]) {
#pragma unroll
  for (int i = 0; i < WA_D / 8; ++i) {
    float t[8];
    load8<T>(p + i * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[i * 8 + j] = t[j];
  }
}

template <typename T>
__global__ void __launch_bounds__(WA_TOK) window_attention_kernel(const T* __restrict__ qkv,
                                                                  const float* __restrict__ logit_scale,
                                                                  const float* __restrict__ rel_bias,
                                                                  const float* __restrict__ qkv_bias,
                                                                  T* __restrict__ out, int H, int W, int PH, int PW,
                                                                  int heads, int shift_y, int shift_x) {
  __shared__ __align__(16) float ks[WA_TOK * WA_LD];
  __shared__ __align__(16) float vs[WA_TOK * WA_LD];
  __shared__ int region[WA_TOK];
  const int i = threadIdx.x;
  const int head = blockIdx.y;
  // (PH, PW) = the map zero-padded to whole windows (F.pad before the qkv Linear): a padded token's q, k, v are the
  // qkv biases (k bias = 0 in V2), it takes part as a key and its own output row is dropped
  const int wpr = PW / 8, wpi = wpr * (PH / 8);
  const int b = blockIdx.x / wpi, wrem = blockIdx.x % wpi;
  const int ys = (wrem / wpr) * 8 + (i >> 3), xs = (wrem % wpr) * 8 + (i & 7);
  const int y = (ys + shift_y) % PH, x = (xs + shift_x) % PW;
  const bool pad = y >= H || x >= W;
  const int C = heads * WA_D;
  const size_t tok = pad ? 0 : (static_cast<size_t>(b) * H + y) * W + x;
  const T* src = qkv + tok * 3 * C + head * WA_D;
  {
    int ry = 0, rx = 0;
    if (shift_y > 0) ry = ys < PH - 8 ? 0 : (ys < PH - shift_y ? 1 : 2);
    if (shift_x > 0) rx = xs < PW - 8 ? 0 : (xs < PW - shift_x ? 1 : 2);
    region[i] = ry * 3 + rx;
  }
  float q[WA_D];
  if (pad) {
#pragma unroll
    for (int d = 0; d < WA_D; ++d) {
      ks[i * WA_LD + d] = 0.f;
      vs[i * WA_LD + d] = qkv_bias ? qkv_bias[2 * C + head * WA_D + d] : 0.f;
      q[d] = 0.f;  // the row is never stored
    }
  } else {
    float t[WA_D];
    load_row32<T>(src + C, t);  // k
    float n2 = 0.f;
#pragma unroll
    for (int d = 0; d < WA_D; ++d) n2 = fmaf(t[d], t[d], n2);
    const float inv = 1.0f / fmaxf(sqrtf(n2), 1e-12f);  // F.normalize
#pragma unroll
    for (int d = 0; d < WA_D; d += 4)
      *reinterpret_cast<float4*>(&ks[i * WA_LD + d]) = make_float4(t[d] * inv, t[d + 1] * inv, t[d + 2] * inv, t[d + 3] * inv);
    load_row32<T>(src + 2 * C, t);  // v
#pragma unroll
    for (int d = 0; d < WA_D; d += 4)
      *reinterpret_cast<float4*>(&vs[i * WA_LD + d]) = make_float4(t[d], t[d + 1], t[d + 2], t[d + 3]);
    load_row32<T>(src, q);
    n2 = 0.f;
#pragma unroll
    for (int d = 0; d < WA_D; ++d) n2 = fmaf(q[d], q[d], n2);
    const float qs = logit_scale[head] / fmaxf(sqrtf(n2), 1e-12f);  // fold the (clamped, exponentiated) logit scale
#pragma unroll
    for (int d = 0; d < WA_D; ++d) q[d] *= qs;
  }
  __syncthreads();
  const int my_region = region[i];
  const float* bias_row = rel_bias + (static_cast<size_t>(head) * WA_TOK + i) * WA_TOK;
  const bool masked = (shift_y | shift_x) != 0;
  float s[WA_TOK];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < WA_TOK; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < WA_D; d += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&ks[j * WA_LD + d]);
      a = fmaf(q[d], k4.x, a);
      a = fmaf(q[d + 1], k4.y, a);
      a = fmaf(q[d + 2], k4.z, a);
      a = fmaf(q[d + 3], k4.w, a);
    }
    a += __ldg(bias_row + j);
    if (masked && region[j] != my_region) a -= 100.0f;
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < WA_TOK; ++j) {
    s[j] = sizeof(T) == 4 ? expf(s[j] - mx) : __expf(s[j] - mx);
    sum += s[j];
  }
  const float inv = 1.0f / sum;
  float o[WA_D];
#pragma unroll
  for (int d = 0; d < WA_D; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < WA_TOK; ++j) {
    const float pj = s[j];
#pragma unroll
    for (int d = 0; d < WA_D; d += 4) {
      const float4 v4 = *reinterpret_cast<const float4*>(&vs[j * WA_LD + d]);
      o[d] = fmaf(pj, v4.x, o[d]);
      o[d + 1] = fmaf(pj, v4.y, o[d + 1]);
      o[d + 2] = fmaf(pj, v4.z, o[d + 2]);
      o[d + 3] = fmaf(pj, v4.w, o[d + 3]);
    }
  }
  if (pad) return;
  T* dst = out + tok * C + head * WA_D;
#pragma unroll
  for (int d = 0; d < WA_D; d += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = o[d + j] * inv;
    store8(dst + d, t);
  }
}

// bf16 throughput path: the same (window, head) problem on the warp-level tensor-core MMA (mma.sync m16n8k16, bf16 in,
// fp32 accumulate).  A 64x64x32 problem is far below a tcgen05 tile (and TMEM allocation / commit latency would
// dominate it), so this is the right-sized instruction; the kernel is then bound by the qkv read + output write.
// Block = 4 warps = one (window, head); warp w owns query rows 16w..16w+15.  S = Q.K^T accumulates in registers, the
// softmax runs on the accumulator fragments (a row lives in 4 lanes), and the same registers, packed to bf16, are the
// A fragments of O = P.V (V fragments through ldmatrix.trans).  The output tile is staged through shared memory so
// the global stores are 16-byte vectors.
constexpr int WM_LD = 40;  // smem row stride in bf16 (80 B): conflict-free fragment loads, 16-byte aligned rows

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void load16(const __nv_bfloat16* p, float (&v)[16]) {
  float a[8], b[8];
  load8<__nv_bfloat16>(p, a);
  load8<__nv_bfloat16>(p + 8, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] = a[j];
    v[8 + j] = b[j];
  }
}
__device__ __forceinline__ void store16_smem(__nv_bfloat16* p, const float (&v)[16], float mul) {
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = v[j] * mul;
    b[j] = v[8 + j] * mul;
  }
  store8(p, a);
  store8(p + 8, b);
}

constexpr int WM_WPB = 4;  // windows per block: the head's bias fragments (32 registers) are loaded once per block

struct WinRaw {  // one thread's share of a window: 16 dims of token (tid / 2)'s q, k, v, still bf16
  uint4 q0, q1, k0, k1, v0, v1;
  long long tok;  // token index in the stored map, -1 for a padded cell
  int region;
};

__global__ void __launch_bounds__(128, 4) window_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                      const float* __restrict__ logit_scale,
                                                                      const float* __restrict__ rel_bias,
                                                                      const float* __restrict__ qkv_bias,
                                                                      __nv_bfloat16* __restrict__ out, int H, int W,
                                                                      int PH, int PW, int heads, int shift_y,
                                                                      int shift_x, int num_windows) {
  __shared__ __align__(16) __nv_bfloat16 qs[WA_TOK * WM_LD];
  __shared__ __align__(16) __nv_bfloat16 ks[WA_TOK * WM_LD];
  __shared__ __align__(16) __nv_bfloat16 vs[WA_TOK * WM_LD];
  __shared__ int region[WA_TOK];
  __shared__ long long tok_of[WA_TOK];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int C = heads * WA_D;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16, i0 = r0 + g, i1 = i0 + 8;
  const bool masked = (shift_y | shift_x) != 0;
  const float lscale = logit_scale[head];
  // relative position bias of this thread's accumulator cells: rows i0 / i1, columns nt*8 + 2t (+1)
  float bf[8][4];
  {
    const float* bias0 = rel_bias + (static_cast<size_t>(head) * WA_TOK + i0) * WA_TOK;
    const float* bias1 = bias0 + 8 * WA_TOK;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 b0 = __ldg(reinterpret_cast<const float2*>(bias0 + nt * 8 + 2 * t));
      const float2 b1 = __ldg(reinterpret_cast<const float2*>(bias1 + nt * 8 + 2 * t));
      bf[nt][0] = b0.x; bf[nt][1] = b0.y; bf[nt][2] = b1.x; bf[nt][3] = b1.y;
    }
  }
  const int ti = tid >> 1, half = tid & 1;  // staging role: thread pair (2i, 2i+1) = the two halves of token i
  const int wpr = PW / 8, wpi = wpr * (PH / 8);
  auto fetch = [&](int win, WinRaw& r) {
    const int b = win / wpi, wrem = win % wpi;
    const int ys = (wrem / wpr) * 8 + (ti >> 3), xs = (wrem % wpr) * 8 + (ti & 7);
    const int y = (ys + shift_y) % PH, x = (xs + shift_x) % PW;
    const bool pad = y >= H || x >= W;
    r.tok = pad ? -1 : (static_cast<long long>(b) * H + y) * W + x;
    int ry = 0, rx = 0;
    if (shift_y > 0) ry = ys < PH - 8 ? 0 : (ys < PH - shift_y ? 1 : 2);
    if (shift_x > 0) rx = xs < PW - 8 ? 0 : (xs < PW - shift_x ? 1 : 2);
    r.region = ry * 3 + rx;
    if (!pad) {
      const uint4* src = reinterpret_cast<const uint4*>(qkv + r.tok * 3 * C + head * WA_D + half * 16);
      const int cs = C / 8;  // uint4 per C channels
      r.q0 = __ldg(src); r.q1 = __ldg(src + 1);
      r.k0 = __ldg(src + cs); r.k1 = __ldg(src + cs + 1);
      r.v0 = __ldg(src + 2 * cs); r.v1 = __ldg(src + 2 * cs + 1);
    }
  };
  auto unpack16 = [](const uint4& a, const uint4& b, float (&v)[16]) {
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      v[2 * j] = f.x;
      v[2 * j + 1] = f.y;
    }
  };
  const int win0 = blockIdx.x * WM_WPB;
  WinRaw raw;
  fetch(win0, raw);
  for (int w = 0; w < WM_WPB && win0 + w < num_windows; ++w) {
    {
      // ---- stage the fetched window: normalise q (logit scale folded in) and k in fp32, round to bf16
      float q[16], k[16], v[16];
      if (raw.tok >= 0) {
        unpack16(raw.q0, raw.q1, q);
        unpack16(raw.k0, raw.k1, k);
        unpack16(raw.v0, raw.v1, v);
      } else {  // padded cell: q, k, v = the qkv biases (k bias is zero); its own output row is dropped
#pragma unroll
        for (int d = 0; d < 16; ++d) {
          q[d] = 0.f;
          k[d] = 0.f;
          v[d] = qkv_bias ? qkv_bias[2 * C + head * WA_D + half * 16 + d] : 0.f;
        }
      }
      float q2 = 0.f, k2 = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        q2 = fmaf(q[d], q[d], q2);
        k2 = fmaf(k[d], k[d], k2);
      }
      q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
      k2 += __shfl_xor_sync(0xffffffffu, k2, 1);
      const float qmul = lscale / fmaxf(sqrtf(q2), 1e-12f);  // F.normalize
      const float kmul = 1.0f / fmaxf(sqrtf(k2), 1e-12f);
      store16_smem(&qs[ti * WM_LD + half * 16], q, qmul);
      store16_smem(&ks[ti * WM_LD + half * 16], k, kmul);
      store16_smem(&vs[ti * WM_LD + half * 16], v, 1.0f);
      if (half == 0) {
        region[ti] = raw.region;
        tok_of[ti] = raw.tok;
      }
    }
    __syncthreads();
    if (w + 1 < WM_WPB && win0 + w + 1 < num_windows) fetch(win0 + w + 1, raw);  // in flight during the MMAs below
    uint32_t a[2][4];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      a[kk][0] = *reinterpret_cast<const uint32_t*>(&qs[i0 * WM_LD + kk * 16 + 2 * t]);
      a[kk][1] = *reinterpret_cast<const uint32_t*>(&qs[i1 * WM_LD + kk * 16 + 2 * t]);
      a[kk][2] = *reinterpret_cast<const uint32_t*>(&qs[i0 * WM_LD + kk * 16 + 2 * t + 8]);
      a[kk][3] = *reinterpret_cast<const uint32_t*>(&qs[i1 * WM_LD + kk * 16 + 2 * t + 8]);
    }
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = bf[nt][e];  // accumulate on top of the position bias
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&ks[(nt * 8 + g) * WM_LD + kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&ks[(nt * 8 + g) * WM_LD + kk * 16 + 2 * t + 8]);
        mma_bf16_16816(s[nt], a[kk], b0, b1);
      }
    }
    // ---- shift mask; fp32 softmax on the accumulator fragments (a row lives in the 4 lanes of a quad)
    const int reg0 = region[i0], reg1 = region[i1];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (masked) {
        const int j = nt * 8 + 2 * t;
        const int rj0 = region[j], rj1 = region[j + 1];
        if (rj0 != reg0) s[nt][0] -= 100.0f;
        if (rj1 != reg0) s[nt][1] -= 100.0f;
        if (rj0 != reg1) s[nt][2] -= 100.0f;
        if (rj1 != reg1) s[nt][3] -= 100.0f;
      }
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = __expf(s[nt][0] - m0);
      s[nt][1] = __expf(s[nt][1] - m0);
      s[nt][2] = __expf(s[nt][2] - m1);
      s[nt][3] = __expf(s[nt][3] - m1);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // ---- O = P.V: the S accumulator fragments of key tiles (2kk, 2kk+1) are the A fragment of key step kk
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, smem_u32(&vs[(kk * 16 + (lane & 15)) * WM_LD + nt * 8]));
        mma_bf16_16816(o[nt], pa, b0, b1);
      }
    }
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    // ---- stage the warp's 16 x 32 output rows in its own (already consumed) rows of qs, then 16-byte global stores
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<uint32_t*>(&qs[i0 * WM_LD + nt * 8 + 2 * t]) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
      *reinterpret_cast<uint32_t*>(&qs[i1 * WM_LD + nt * 8 + 2 * t]) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
    __syncwarp();
#pragma unroll
    for (int c = lane; c < 64; c += 32) {
      const int row = r0 + (c >> 2), part = c & 3;
      const long long tok = tok_of[row];
      if (tok >= 0)
        *reinterpret_cast<uint4*>(out + tok * C + head * WA_D + part * 8) =
            *reinterpret_cast<const uint4*>(&qs[row * WM_LD + part * 8]);
    }
    __syncthreads();  // every warp is done with this window's K / V / region before the next one is staged
  }
}

// ------------------------------------------------------------------------------------------ patch merging gather
// out[b, y, x, q*C + c] = in[b, 2y + (q & 1), 2x + (q >> 1), c]   (x0 | x1 | x2 | x3 of _patch_merging_pad)
__global__ void __launch_bounds__(256) patch_merge_gather_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                                 long long total, int H, int W, int vec_per_pix) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % vec_per_pix);
  long long t = idx / vec_per_pix;
  const int q = static_cast<int>(t % 4);
  t /= 4;
  const int ox = static_cast<int>(t % (W / 2));
  t /= (W / 2);
  const int oy = static_cast<int>(t % (H / 2));
  const long long b = t / (H / 2);
  out[idx] = in[((b * H + 2 * oy + (q & 1)) * W + 2 * ox + (q >> 1)) * vec_per_pix + v];
}

// ------------------------------------------------------------------------------------------ final norm + pool
// One block per image: warp w normalises tokens w, w+8, ... and accumulates the normalised rows per lane; the eight
// partial sums are folded in a fixed order.
__global__ void __launch_bounds__(256) ln_meanpool_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float eps,
                                                          float* __restrict__ out, int tokens, int C) {
  extern __shared__ float part[];  // [8][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = C / 32;          // <= 32
  const float* xb = x + static_cast<size_t>(blockIdx.x) * tokens * C;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int t = warp; t < tokens; t += 8) {
    float v[32];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      v[i] = i < per ? xb[static_cast<size_t>(t) * C + lane + 32 * i] : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float d = i < per ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / static_cast<float>(C) + eps);
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < per) acc[i] += (v[i] - mean) * rstd;
  }
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < per) part[warp * C + lane + 32 * i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += part[w * C + c];
    // mean_t(norm(x_t) * gamma + beta) = mean_t(norm(x_t)) * gamma + beta
    out[static_cast<size_t>(blockIdx.x) * C + c] = a / static_cast<float>(tokens) * gamma[c] + beta[c];
  }
}

// ------------------------------------------------------------------------------------------ set reduction
__global__ void __launch_bounds__(256) set_reduce_kernel(const float* __restrict__ x, float* __restrict__ out, int n,
                                                         int F, int mode, long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long b = idx / F;
  const int f = static_cast<int>(idx % F);
  const float* p = x + b * n * F + f;
  float a = p[0];
  for (int i = 1; i < n; ++i) {
    const float v = p[static_cast<size_t>(i) * F];
    a = mode == 1 ? fmaxf(a, v) : a + v;
  }
  out[idx] = mode == 1 ? a : a / static_cast<float>(n);
}

// ------------------------------------------------------------------------------------------ sViT (style_agg=svit)
// SPT patch tokens (networks/vit_set.py:97-107): the ns style images of a sample are stacked along channels
// (channel c*ns + s) and cut into p x p patches flattened as (p1 p2 c).  One thread per output element (coalesced
// writes; the gathered reads are 12-byte pixel triples that stay in L1/L2).
__global__ void __launch_bounds__(256) spt_patchify_kernel(const float* __restrict__ img, float* __restrict__ out,
                                                           long long total, int ns, int P, int p) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cn = 3 * ns, pd = p * p * cn, g = P / p;
  const int e = static_cast<int>(idx % pd);
  const long long tokidx = idx / pd;
  const int tok = static_cast<int>(tokidx % (g * g));
  const long long b = tokidx / (g * g);
  const int cs = e % cn, pp = e / cn;
  const int c = cs / ns, s = cs % ns;
  const int y = (tok / g) * p + pp / p, x = (tok % g) * p + pp % p;
  out[idx] = img[(((b * ns + s) * P + y) * P + x) * 3 + c];
}

// x[b, 0] = cls + pos[0];  x[b, 1] = t_emb (zeros when None) + pos[1];  x[b, 2 + i] = patch_i + pos[2 + i]
// (vit_set.py:182-190); rows >= T of the padded token buffer are zero.
template <typename TI>
__global__ void __launch_bounds__(256) svit_assemble_kernel(const TI* __restrict__ patches, const float* __restrict__ cls,
                                                            const float* __restrict__ pos, float* __restrict__ out,
                                                            long long total, int n_patches, int t_pad, int dim) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int d = static_cast<int>(idx % dim);
  const long long r = idx / dim;
  const int t = static_cast<int>(r % t_pad);
  const long long b = r / t_pad;
  float v = 0.f;
  if (t < n_patches + 2) {
    v = pos[static_cast<size_t>(t) * dim + d];
    if (t == 0) v += cls[d];
    else if (t >= 2) v += to_f32<TI>(patches[(b * n_patches + (t - 2)) * dim + d]);
  }
  out[idx] = v;
}

// mean over the first `tokens` rows of each sample of a [batch][t_pad][c] fp32 buffer (pool = 'mean', vit_set.py:196)
__global__ void __launch_bounds__(256) token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int tokens,
                                                         int t_pad, int C) {
  __shared__ float part[8][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 32 + lane;
  const float* xb = x + static_cast<size_t>(blockIdx.y) * t_pad * C;
  float a = 0.f;
  if (c < C)
    for (int t = warp; t < tokens; t += 8) a += xb[static_cast<size_t>(t) * C + c];
  part[warp][lane] = a;
  __syncthreads();
  if (warp == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][lane];
    out[static_cast<size_t>(blockIdx.y) * C + c] = s / static_cast<float>(tokens);
  }
}

}  // namespace

extern "C" int stedm_spt_patchify(const float* img, float* out, int batch, int ns, int p_img, int patch, void* stream) {
  STEDM_REQUIRE(img && out && batch > 0 && ns > 0 && patch > 0 && p_img > 0 && p_img % patch == 0,
                "spt_patchify: bad argument");
  const long long total = static_cast<long long>(batch) * ns * p_img * p_img * 3;
  const long long blocks = (total + 255) / 256;
  STEDM_REQUIRE(blocks < (1LL << 31), "spt_patchify: too large");
  spt_patchify_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, out, total, ns,
                                                                                                    p_img, patch);
  return check_launch("spt_patchify");
}

extern "C" int stedm_svit_assemble(const void* patches, int dtype, const float* cls, const float* pos, float* out,
                                   int batch, int n_patches, int t_pad, int dim, void* stream) {
  STEDM_REQUIRE(patches && cls && pos && out && batch > 0 && n_patches > 0 && t_pad >= n_patches + 2 && dim > 0,
                "svit_assemble: bad argument");
  const long long total = static_cast<long long>(batch) * t_pad * dim;
  const long long blocks = (total + 255) / 256;
  STEDM_REQUIRE(blocks < (1LL << 31), "svit_assemble: too large");
  auto s = static_cast<cudaStream_t>(stream);
  if (dtype == DT_BF16)
    svit_assemble_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(patches), cls, pos, out, total, n_patches, t_pad, dim);
  else
    svit_assemble_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const float*>(patches), cls,
                                                                               pos, out, total, n_patches, t_pad, dim);
  return check_launch("svit_assemble");
}

extern "C" int stedm_token_mean(const float* x, float* out, int batch, int tokens, int t_pad, int c, void* stream) {
  STEDM_REQUIRE(x && out && batch > 0 && tokens > 0 && t_pad >= tokens && c > 0 && batch <= 65535,
                "token_mean: bad argument");
  dim3 grid((c + 31) / 32, batch);
  token_mean_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, tokens, t_pad, c);
  return check_launch("token_mean");
}

extern "C" int stedm_patch_embed_ln(const float* img, const float* w, const float* bias, const float* gamma,
                                    const float* beta, float eps, float* out_f32, void* out_bf16, int batch, int p,
                                    int patch, int embed, void* stream) {
  STEDM_REQUIRE(img && w && gamma && beta && (out_f32 || out_bf16), "patch_embed_ln: null pointer");
  STEDM_REQUIRE(patch == 4 && p > 0 && p % 4 == 0 && batch > 0, "patch_embed_ln: patch must be 4 and divide the image");
  STEDM_REQUIRE(embed == 32 || embed == 64 || embed == 96 || embed == 128, "patch_embed_ln: embed %d unsupported", embed);
  STEDM_REQUIRE(p % 16 == 0 || (p * 3) % 4 == 0, "patch_embed_ln: image rows must be 16-byte aligned");
  const long long tokens = static_cast<long long>(batch) * (p / 4) * (p / 4);
  const long long want = (tokens + 127) / 128;
  const int grid = static_cast<int>(want < 148 * 16 ? want : 148 * 16);
  auto s = static_cast<cudaStream_t>(stream);
  auto ob = static_cast<__nv_bfloat16*>(out_bf16);
  switch (embed) {
    case 32: patch_embed_ln_kernel<32><<<grid, 128, 0, s>>>(img, w, bias, gamma, beta, eps, out_f32, ob, tokens, p); break;
    case 64: patch_embed_ln_kernel<64><<<grid, 128, 0, s>>>(img, w, bias, gamma, beta, eps, out_f32, ob, tokens, p); break;
    case 96: patch_embed_ln_kernel<96><<<grid, 128, 0, s>>>(img, w, bias, gamma, beta, eps, out_f32, ob, tokens, p); break;
    default: patch_embed_ln_kernel<128><<<grid, 128, 0, s>>>(img, w, bias, gamma, beta, eps, out_f32, ob, tokens, p); break;
  }
  return check_launch("patch_embed_ln");
}

extern "C" int stedm_layernorm(const void* x, int x_dtype, const float* residual, const float* gamma, const float* beta,
                               float eps, float* out_f32, void* out_bf16, long long rows, int c, void* stream) {
  STEDM_REQUIRE(x && gamma && beta && (out_f32 || out_bf16), "layernorm: null pointer");
  STEDM_REQUIRE(rows > 0 && c >= 8 && c % 8 == 0 && c <= 2048, "layernorm: C = %d must be a multiple of 8, <= 2048", c);
  const int items = c / 8;
  const int per = items > 128 ? 8 : 4;      // items per lane
  int G = 1;
  while (G < 32 && G * per < items) G <<= 1;
  if (G < 4) G = 4;
  const int rows_per_block = 8 * (32 / G);
  const long long blocks = (rows + rows_per_block - 1) / rows_per_block;
  STEDM_REQUIRE(blocks < (1LL << 31), "layernorm: too many rows");
  auto s = static_cast<cudaStream_t>(stream);
  auto ob = static_cast<__nv_bfloat16*>(out_bf16);
  const unsigned nb = static_cast<unsigned>(blocks);
  if (x_dtype == DT_BF16) {
    auto xp = static_cast<const __nv_bfloat16*>(x);
    if (per == 4) layernorm_kernel<__nv_bfloat16, 4><<<nb, 256, 0, s>>>(xp, residual, gamma, beta, eps, out_f32, ob, rows, c, G);
    else layernorm_kernel<__nv_bfloat16, 8><<<nb, 256, 0, s>>>(xp, residual, gamma, beta, eps, out_f32, ob, rows, c, G);
  } else {
    auto xp = static_cast<const float*>(x);
    if (per == 4) layernorm_kernel<float, 4><<<nb, 256, 0, s>>>(xp, residual, gamma, beta, eps, out_f32, ob, rows, c, G);
    else layernorm_kernel<float, 8><<<nb, 256, 0, s>>>(xp, residual, gamma, beta, eps, out_f32, ob, rows, c, G);
  }
  return check_launch("layernorm");
}

extern "C" int stedm_window_attention(const void* qkv, int dtype, const float* logit_scale, const float* rel_bias,
                                      const float* qkv_bias, void* out, int batch, int h, int w, int heads,
                                      int head_dim, int window, int shift, void* stream) {
  STEDM_REQUIRE(qkv && logit_scale && rel_bias && out, "window_attention: null pointer");
  STEDM_REQUIRE(window == 8 && head_dim == WA_D, "window_attention: window 8 and head_dim 32 only (swin_v2_t)");
  STEDM_REQUIRE(batch > 0 && heads > 0 && h > 0 && w > 0, "window_attention: bad shape");
  STEDM_REQUIRE(shift >= 0 && shift < 8, "window_attention: bad shift");
  const int ph = (h + 7) / 8 * 8, pw = (w + 7) / 8 * 8;  // pad the map to whole windows
  // swin_transformer.py: no shift along an axis the window already covers
  const int sy = ph <= 8 ? 0 : shift, sx = pw <= 8 ? 0 : shift;
  const long long windows = static_cast<long long>(batch) * (ph / 8) * (pw / 8);
  STEDM_REQUIRE(windows < (1LL << 31) && heads <= 65535, "window_attention: grid too large");
  dim3 grid(static_cast<unsigned>(windows), static_cast<unsigned>(heads));
  dim3 grid_mma(static_cast<unsigned>((windows + WM_WPB - 1) / WM_WPB), static_cast<unsigned>(heads));
  auto s = static_cast<cudaStream_t>(stream);
  static const bool simt_bf16 = [] {  // STEDM_WINATTN_SIMT=1: CUDA-core kernel for bf16 too (A/B measurements)
    const char* e = getenv("STEDM_WINATTN_SIMT");
    return e && e[0] == '1';
  }();
  if (dtype == DT_BF16 && !simt_bf16)
    window_attention_mma_kernel<<<grid_mma, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), logit_scale, rel_bias,
                                                         qkv_bias, static_cast<__nv_bfloat16*>(out), h, w, ph, pw,
                                                         heads, sy, sx, static_cast<int>(windows));
  else if (dtype == DT_BF16)
    window_attention_kernel<__nv_bfloat16><<<grid, WA_TOK, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), logit_scale,
                                                                   rel_bias, qkv_bias, static_cast<__nv_bfloat16*>(out),
                                                                   h, w, ph, pw, heads, sy, sx);
  else
    window_attention_kernel<float><<<grid, WA_TOK, 0, s>>>(static_cast<const float*>(qkv), logit_scale, rel_bias,
                                                           qkv_bias, static_cast<float*>(out), h, w, ph, pw, heads, sy,
                                                           sx);
  return check_launch("window_attention");
}

extern "C" int stedm_patch_merge_gather(const void* x, void* out, int dtype, int batch, int h, int w, int c,
                                        void* stream) {
  STEDM_REQUIRE(x && out, "patch_merge_gather: null pointer");
  const int es = dtype_size(dtype);
  STEDM_REQUIRE(batch > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && (c * es) % 16 == 0,
                "patch_merge_gather: even map and 16-byte channel rows required");
  const int vpp = c * es / 16;
  const long long total = static_cast<long long>(batch) * (h / 2) * (w / 2) * 4 * vpp;
  const long long blocks = (total + 255) / 256;
  STEDM_REQUIRE(blocks < (1LL << 31), "patch_merge_gather: too large");
  patch_merge_gather_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), total, h, w, vpp);
  return check_launch("patch_merge_gather");
}

extern "C" int stedm_ln_meanpool(const float* x, const float* gamma, const float* beta, float eps, float* out,
                                 int batch, int tokens, int c, void* stream) {
  STEDM_REQUIRE(x && gamma && beta && out, "ln_meanpool: null pointer");
  STEDM_REQUIRE(batch > 0 && tokens > 0 && c % 32 == 0 && c >= 32 && c <= 1024, "ln_meanpool: C = %d unsupported", c);
  ln_meanpool_kernel<<<batch, 256, 8 * c * sizeof(float), static_cast<cudaStream_t>(stream)>>>(x, gamma, beta, eps, out,
                                                                                               tokens, c);
  return check_launch("ln_meanpool");
}

extern "C" int stedm_set_reduce(const float* x, float* out, int b, int n, int f, int mode, void* stream) {
  STEDM_REQUIRE(x && out && b > 0 && n > 0 && f > 0 && (mode == 0 || mode == 1), "set_reduce: bad argument");
  const long long total = static_cast<long long>(b) * f;
  set_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, out, n, f, mode, total);
  return check_launch("set_reduce");
}
