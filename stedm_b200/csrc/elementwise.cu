// Memory-bound kernels of the sampling path: fused CFG + DDIM update (K11), layout packing, nearest upsample,
// stride-2 gather, timestep embedding + small linears (K8), VQ nearest code (K12), spatial rescaler (K13),
// uint8 image conversion, row softmax.  All are coalesced along the innermost dimension and vectorised
// (16 B per thread) where the layout allows; reductions use warp shuffles.
#include <math.h>

#include "../../include/stedm_b200.h"
#include "common.cuh"

using namespace stedm;

// =====================================================================================================
// K11: CFG combine + (C,H)-std rescale + DDIM update.  One block = 32 columns (w) of one sample; 8 row lanes.
// Reference: ldm/models/diffusion/ddim.py:177-209.  dims=(1,2) -> statistics per (b, w) over C*H values, unbiased.
// =====================================================================================================
__global__ void __launch_bounds__(256) cfg_ddim_kernel(const float* __restrict__ e_c, const float* __restrict__ e_u,
                                                       const float* __restrict__ x, const float* __restrict__ noise,
                                                       float* __restrict__ x_prev, float* __restrict__ pred_x0, int CH,
                                                       int W, int guided, float scale, float phi, float a_t,
                                                       float a_prev, float sigma, float s1m) {
  __shared__ float red[4][8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int w = blockIdx.x * 32 + tx;
  const bool active = w < W;
  const size_t base = static_cast<size_t>(blockIdx.y) * CH * W + (active ? w : 0);
  float ratio = 1.0f;
  if (guided) {
    float sc = 0.f, sw = 0.f;
    if (active)
      for (int r = ty; r < CH; r += 8) {
        const float ec = e_c[base + static_cast<size_t>(r) * W], eu = e_u[base + static_cast<size_t>(r) * W];
        sc += ec;
        sw += eu + scale * (ec - eu);
      }
    red[0][ty][tx] = sc;
    red[1][ty][tx] = sw;
    __syncthreads();
    float mc = 0.f, mw = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mc += red[0][i][tx];
      mw += red[1][i][tx];
    }
    mc /= static_cast<float>(CH);
    mw /= static_cast<float>(CH);
    float qc = 0.f, qw = 0.f;
    if (active)
      for (int r = ty; r < CH; r += 8) {
        const float ec = e_c[base + static_cast<size_t>(r) * W], eu = e_u[base + static_cast<size_t>(r) * W];
        const float dc = ec - mc, dw = (eu + scale * (ec - eu)) - mw;
        qc += dc * dc;
        qw += dw * dw;
      }
    red[2][ty][tx] = qc;
    red[3][ty][tx] = qw;
    __syncthreads();
    float vc = 0.f, vw = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      vc += red[2][i][tx];
      vw += red[3][i][tx];
    }
    const float denom = static_cast<float>(CH - 1);
    ratio = __fdiv_rn(__fsqrt_rn(vc / denom), __fsqrt_rn(vw / denom));
  }
  if (!active) return;
  const float inv_sqrt_at = __fsqrt_rn(a_t);
  const float sqrt_aprev = __fsqrt_rn(a_prev);
  const float dir_coef = __fsqrt_rn(1.0f - a_prev - sigma * sigma);
  const float one_m_phi = 1.0f - phi;
  for (int r = ty; r < CH; r += 8) {
    const size_t o = base + static_cast<size_t>(r) * W;
    float e = e_c[o];
    if (guided) {
      const float eu = e_u[o];
      const float ew = eu + scale * (e - eu);
      e = (ew * ratio) * phi + one_m_phi * e;
    }
    const float p0 = __fdiv_rn(x[o] - s1m * e, inv_sqrt_at);
    float xp = sqrt_aprev * p0 + dir_coef * e;
    if (noise != nullptr) xp += sigma * noise[o];
    pred_x0[o] = p0;
    x_prev[o] = xp;
  }
}

extern "C" int stedm_cfg_ddim_step(const float* e_c, const float* e_u, const float* x, const float* noise,
                                   float* x_prev, float* pred_x0, int batch, int channels, int height, int width,
                                   int guided, float cfg_scale, float phi, float a_t, float a_prev, float sigma_t,
                                   float sqrt_one_minus_at, void* stream) {
  STEDM_REQUIRE(e_c && x && x_prev && pred_x0 && (!guided || e_u), "cfg_ddim_step: null pointer");
  STEDM_REQUIRE(batch > 0 && channels > 0 && height > 0 && width > 0, "cfg_ddim_step: bad shape");
  STEDM_REQUIRE(!guided || channels * height > 1, "cfg_ddim_step: std over a single value");
  dim3 grid((width + 31) / 32, batch), block(32, 8);
  cfg_ddim_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      e_c, e_u, x, noise, x_prev, pred_x0, channels * height, width, guided, cfg_scale, phi, a_t, a_prev, sigma_t,
      sqrt_one_minus_at);
  return check_launch("cfg_ddim_step");
}

// =====================================================================================================
// NCHW fp32 [x0 | x1] -> NHWC (bf16|fp32), channels zero-padded to c_pad.  One thread = one pixel; reads are
// coalesced along the pixel index for every channel, the c_pad-wide row is written with 16 B stores.
// =====================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256) pack_nchw_to_nhwc_kernel(const float* __restrict__ x0, int c0,
                                                                const float* __restrict__ x1, int c1,
                                                                TO* __restrict__ out, int hw, int c_pad) {
  // one thread = one 16-byte chunk of one pixel's channel row: consecutive threads write consecutive chunks
  // (fully coalesced 16 B stores; the zero padding is most of the row), the few real channels are gathered
  constexpr int VEC = 16 / sizeof(TO);
  const int chunks = c_pad / VEC;
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= static_cast<size_t>(hw) * chunks) return;
  const int p = static_cast<int>(i / chunks), ch0 = static_cast<int>(i % chunks) * VEC;
  float v[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int c = ch0 + j;
    float t = 0.f;
    if (c < c0)
      t = x0[(static_cast<size_t>(b) * c0 + c) * hw + p];
    else if (c < c0 + c1)
      t = x1[(static_cast<size_t>(b) * c1 + (c - c0)) * hw + p];
    v[j] = t;
  }
  TO* o = out + (static_cast<size_t>(b) * hw + p) * c_pad + ch0;
  if constexpr (sizeof(TO) == 2) {
    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                              pack_bf16x2(v[6], v[7]));
  } else {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

extern "C" int stedm_pack_nchw_to_nhwc(const float* x0, int c0, const float* x1, int c1, void* out, int out_dtype,
                                       int batch, int hw, int c_pad, void* stream) {
  STEDM_REQUIRE(x0 && out && (c1 == 0 || x1), "pack_nchw_to_nhwc: null pointer");
  STEDM_REQUIRE(c0 + c1 <= c_pad && batch > 0 && hw > 0, "pack_nchw_to_nhwc: bad shape");
  const int vec = out_dtype == DT_BF16 ? 8 : 4;
  STEDM_REQUIRE(c_pad % vec == 0, "pack_nchw_to_nhwc: c_pad must be a multiple of %d", vec);
  const size_t items = static_cast<size_t>(hw) * (c_pad / vec);
  dim3 grid(static_cast<unsigned>((items + 255) / 256), batch);
  auto s = static_cast<cudaStream_t>(stream);
  if (out_dtype == DT_BF16)
    pack_nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x0, c0, x1, c1, static_cast<__nv_bfloat16*>(out), hw, c_pad);
  else
    pack_nchw_to_nhwc_kernel<float><<<grid, 256, 0, s>>>(x0, c0, x1, c1, static_cast<float*>(out), hw, c_pad);
  return check_launch("pack_nchw_to_nhwc");
}

// NHWC -> NCHW fp32 through a 32x32 shared-memory transpose (coalesced on both sides).
template <typename TI>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ x, float* __restrict__ out, int hw, int c) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int p = p0 + i, cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < hw && cc < c) ? to_f32<TI>(x[(static_cast<size_t>(b) * hw + p) * c + cc]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int cc = c0 + i, p = p0 + threadIdx.x;
    if (p < hw && cc < c) out[(static_cast<size_t>(b) * c + cc) * hw + p] = tile[threadIdx.x][i];
  }
}

extern "C" int stedm_nhwc_to_nchw_f32(const void* x, int dtype, float* out, int batch, int hw, int c, void* stream) {
  STEDM_REQUIRE(x && out && batch > 0 && hw > 0 && c > 0, "nhwc_to_nchw: bad argument");
  dim3 grid((hw + 31) / 32, (c + 31) / 32, batch), block(32, 8);
  auto s = static_cast<cudaStream_t>(stream);
  if (dtype == DT_BF16)
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(static_cast<const __nv_bfloat16*>(x), out, hw, c);
  else
    nhwc_to_nchw_kernel<float><<<grid, block, 0, s>>>(static_cast<const float*>(x), out, hw, c);
  return check_launch("nhwc_to_nchw");
}

// =====================================================================================================
// Nearest x2 upsample and stride-2 3x3 gather, NHWC, 16 B vectors (c*sizeof(T) % 16 == 0).
// =====================================================================================================
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int h, int w, int cv,
                                  size_t total) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t p = i / cv;
    const int ox = static_cast<int>(p % (2 * w));
    p /= (2 * w);
    const int oy = static_cast<int>(p % (2 * h));
    const size_t b = p / (2 * h);
    out[i] = x[((b * h + (oy >> 1)) * w + (ox >> 1)) * cv + v];
  }
}

extern "C" int stedm_upsample_nearest2x(const void* x, void* out, int dtype, int batch, int h, int w, int c,
                                        void* stream) {
  const int es = dtype_size(dtype);
  STEDM_REQUIRE(x && out && (c * es) % 16 == 0, "upsample_nearest2x: channels*elsize must be a multiple of 16 B");
  const int cv = c * es / 16;
  const size_t total = static_cast<size_t>(batch) * 4 * h * w * cv;
  const int blocks = static_cast<int>(min(static_cast<size_t>(148 * 16), (total + 255) / 256));
  upsample2x_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), h, w, cv, total);
  return check_launch("upsample_nearest2x");
}

__global__ void im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int h, int w, int cv,
                                 size_t total) {
  const int ho = h / 2, wo = w / 2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % cv);
    size_t p = i / cv;
    const int tap = static_cast<int>(p % 9);
    p /= 9;
    const int ox = static_cast<int>(p % wo);
    p /= wo;
    const int oy = static_cast<int>(p % ho);
    const size_t b = p / ho;
    const int iy = oy * 2 + tap / 3 - 1, ix = ox * 2 + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) val = x[((b * h + iy) * w + ix) * cv + v];
    out[i] = val;
  }
}

extern "C" int stedm_im2col_3x3_s2(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream) {
  const int es = dtype_size(dtype);
  STEDM_REQUIRE(x && out && (c * es) % 16 == 0 && h % 2 == 0 && w % 2 == 0, "im2col_3x3_s2: bad shape");
  const int cv = c * es / 16;
  const size_t total = static_cast<size_t>(batch) * (h / 2) * (w / 2) * 9 * cv;
  const int blocks = static_cast<int>(min(static_cast<size_t>(148 * 16), (total + 255) / 256));
  im2col_s2_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(x),
                                                                          static_cast<uint4*>(out), h, w, cv, total);
  return check_launch("im2col_3x3_s2");
}

// =====================================================================================================
// K8: sinusoidal timestep embedding and small row-vector linears (fp32; weights are read once per launch).
// =====================================================================================================
template <typename TT>
__global__ void timestep_embedding_kernel(const TT* __restrict__ t, float* __restrict__ out, int dim) {
  const int b = blockIdx.x, half = dim / 2;
  const float tv = static_cast<float>(t[b]);
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    // util.py:163-166: freqs = exp(-ln(10000) * i / half) in fp32, args = t * freqs, [cos | sin]
    const float f = expf(-9.210340371976184f * static_cast<float>(i) / static_cast<float>(half));
    const float a = tv * f;
    out[static_cast<size_t>(b) * dim + i] = cosf(a);
    out[static_cast<size_t>(b) * dim + half + i] = sinf(a);
  }
}

extern "C" int stedm_timestep_embedding(const long long* t, float* out, int batch, int dim, void* stream) {
  STEDM_REQUIRE(t && out && batch > 0 && dim > 0 && dim % 2 == 0, "timestep_embedding: bad argument");
  timestep_embedding_kernel<long long><<<batch, 64, 0, static_cast<cudaStream_t>(stream)>>>(t, out, dim);
  return check_launch("timestep_embedding");
}

extern "C" int stedm_timestep_embedding_f32(const float* t, float* out, int batch, int dim, void* stream) {
  STEDM_REQUIRE(t && out && batch > 0 && dim > 0 && dim % 2 == 0, "timestep_embedding_f32: bad argument");
  timestep_embedding_kernel<float><<<batch, 64, 0, static_cast<cudaStream_t>(stream)>>>(t, out, dim);
  return check_launch("timestep_embedding_f32");
}

// One warp per output feature n; the weight row is read once (float4) and reused for up to 8 batch rows held in
// registers; activations (batch x k) are staged in shared memory with SiLU applied once.
constexpr int LIN_BT = 8;
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ out,
                                                     int batch, int k, int n, int act) {
  extern __shared__ float s_in[];  // [LIN_BT][k]
  const int b0 = blockIdx.y * LIN_BT;
  const int nb = min(LIN_BT, batch - b0);
  for (int i = threadIdx.x; i < nb * k; i += blockDim.x) {
    float v = in[static_cast<size_t>(b0) * k + i];
    s_in[i] = (act & 1) ? silu_precise(v) : ((act & 2) ? fmaxf(v, 0.f) : v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 8 + warp;
  if (col >= n) return;
  float acc[LIN_BT];
#pragma unroll
  for (int j = 0; j < LIN_BT; ++j) acc[j] = 0.f;
  const float* wr = w + static_cast<size_t>(col) * k;
  for (int i = lane; i < k; i += 32) {
    const float wv = wr[i];
#pragma unroll
    for (int j = 0; j < LIN_BT; ++j)
      if (j < nb) acc[j] += wv * s_in[j * k + i];
  }
#pragma unroll
  for (int j = 0; j < LIN_BT; ++j) {
    const float v = warp_sum(acc[j]);
    if (lane == 0 && j < nb) {
      const float y = v + (bias ? bias[col] : 0.f);
      out[static_cast<size_t>(b0 + j) * n + col] = (act & 4) ? fmaxf(y, 0.f) : y;
    }
  }
}

extern "C" int stedm_linear(const float* in, const float* w, const float* bias, float* out, int batch, int k, int n,
                            int act, void* stream) {
  STEDM_REQUIRE(in && w && out && batch > 0 && k > 0 && n > 0, "linear: bad argument");
  const size_t smem = static_cast<size_t>(LIN_BT) * k * 4;
  STEDM_REQUIRE(smem <= 200 * 1024, "linear: k = %d too large for the staging buffer", k);
  if (smem > 48 * 1024) {
    static DeviceOnce configured;
    if (configured.needed()) {
      if (cudaFuncSetAttribute(linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
        set_error("linear: cannot raise the dynamic shared memory limit");
        return ERR_CUDA;
      }
      configured.done();
    }
  }
  dim3 grid((n + 7) / 8, (batch + LIN_BT - 1) / LIN_BT);
  linear_kernel<<<grid, 256, static_cast<size_t>(LIN_BT) * k * 4, static_cast<cudaStream_t>(stream)>>>(
      in, w, bias, out, batch, k, n, act);
  return check_launch("linear");
}

// =====================================================================================================
// K12: VQ nearest code.  One warp per latent pixel group: codebook staged in shared memory as (e, |e|^2);
// each thread scans codes lane, lane+32, ... for its pixel set, then a warp argmin (ties -> lowest index,
// matching torch.argmin's first-minimum rule).  Distances use the reference's expanded form
// |z|^2 + |e|^2 - 2 z.e in fp32.
// =====================================================================================================
// Round 2: the codebook is staged channel-planar ([C + 1][n_codes]: conflict-free when lane i reads code i + 32 k — the
// [code][C + 1] layout of round 1 put the 32 lanes on 8 banks) and a warp scans the codes for VQ_PIX pixels at once, so
// every code component read from shared memory feeds VQ_PIX distance updates (1.96 -> 0.90 ms per 64 latents of 64 x 64:
// ~10 instructions per (pixel, code) pair, i.e. near the issue limit of this exact-fp32 formulation).
constexpr int VQ_PIX = 4;
template <int C>
__global__ void __launch_bounds__(256) vq_nearest_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                         float* __restrict__ zq, int* __restrict__ idx, int hw,
                                                         int n_codes, size_t n_pix) {
  extern __shared__ float s_cb[];  // [C + 1][n_codes]: components, then |e|^2
  for (int i = threadIdx.x; i < n_codes; i += blockDim.x) {
    float n2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float e = cb[static_cast<size_t>(i) * C + c];
      s_cb[c * n_codes + i] = e;
      n2 += e * e;
    }
    s_cb[C * n_codes + i] = n2;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t warps_total = static_cast<size_t>(gridDim.x) * (blockDim.x >> 5);
  const size_t n_grp = (n_pix + VQ_PIX - 1) / VQ_PIX;
  for (size_t g = static_cast<size_t>(blockIdx.x) * (blockDim.x >> 5) + warp; g < n_grp; g += warps_total) {
    float zv[VQ_PIX][C], z2[VQ_PIX], best[VQ_PIX];
    int besti[VQ_PIX];
#pragma unroll
    for (int q = 0; q < VQ_PIX; ++q) {
      const size_t p = min(g * VQ_PIX + q, n_pix - 1);     // a ragged last group repeats its last pixel
      const size_t b = p / hw, pix = p % hw;
      z2[q] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        zv[q][c] = z[(b * C + c) * hw + pix];
        z2[q] += zv[q][c] * zv[q][c];
      }
      best[q] = INFINITY;
      besti[q] = 0x7fffffff;
    }
    for (int i = lane; i < n_codes; i += 32) {
      float e[C];
#pragma unroll
      for (int c = 0; c < C; ++c) e[c] = s_cb[c * n_codes + i];
      const float n2 = s_cb[C * n_codes + i];
#pragma unroll
      for (int q = 0; q < VQ_PIX; ++q) {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) dot += zv[q][c] * e[c];
        const float d = (z2[q] + n2) - 2.0f * dot;        // the reference's expanded form, same operation order
        if (d < best[q]) {
          best[q] = d;
          besti[q] = i;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < VQ_PIX; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best[q], o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti[q], o);
        if (ob < best[q] || (ob == best[q] && oi < besti[q])) {
          best[q] = ob;
          besti[q] = oi;
        }
      }
      const size_t p = g * VQ_PIX + q;
      if (p >= n_pix) continue;
      // a pixel whose every distance is NaN (diverged sampling) wins no comparison: code 0, like torch.argmin on an
      // all-NaN row, instead of an out-of-bounds gather
      int bi = besti[q];
      if (static_cast<unsigned>(bi) >= static_cast<unsigned>(n_codes)) bi = 0;
      const size_t b = p / hw, pix = p % hw;
      if (lane < C) zq[(b * C + lane) * hw + pix] = s_cb[lane * n_codes + bi];
      if (lane == 0 && idx != nullptr) idx[p] = bi;
    }
  }
}

extern "C" int stedm_vq_nearest(const float* z, const float* codebook, float* zq, int* idx, int batch, int c, int hw,
                                int n_codes, void* stream) {
  STEDM_REQUIRE(z && codebook && zq && batch > 0 && hw > 0 && n_codes > 0, "vq_nearest: bad argument");
  STEDM_REQUIRE(c == 3 || c == 4, "vq_nearest: embed_dim %d unsupported (3 or 4)", c);
  const size_t smem = static_cast<size_t>(n_codes) * (c + 1) * 4;
  STEDM_REQUIRE(smem <= 200 * 1024, "vq_nearest: codebook too large for shared memory");
  const size_t n_pix = static_cast<size_t>(batch) * hw;
  const int blocks = static_cast<int>(min(static_cast<size_t>(148), (n_pix + 8 * VQ_PIX - 1) / (8 * VQ_PIX)));
  auto s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (c == 3) {
    e = cudaFuncSetAttribute(vq_nearest_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) vq_nearest_kernel<3><<<blocks, 256, smem, s>>>(z, codebook, zq, idx, hw, n_codes, n_pix);
  } else {
    e = cudaFuncSetAttribute(vq_nearest_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) vq_nearest_kernel<4><<<blocks, 256, smem, s>>>(z, codebook, zq, idx, hw, n_codes, n_pix);
  }
  if (e != cudaSuccess) {
    set_error("vq_nearest: %s", cudaGetErrorString(e));
    return ERR_CUDA;
  }
  return check_launch("vq_nearest");
}

// =====================================================================================================
// K13: SpatialRescaler = n_stages x (bilinear 0.5, align_corners=False) == nested 2x2 means, then 1x1 conv.
// The nested-mean order ((a+b)/2 pairs, stage by stage) follows the reference's two interpolate calls.
// =====================================================================================================
__device__ float nested_mean(const float* __restrict__ img, int p, int y0, int x0, int size) {
  if (size == 1) return img[static_cast<size_t>(y0) * p + x0];
  const int h = size / 2;
  // one bilinear-0.5 stage: out = 0.5*(0.5*a + 0.5*b) + 0.5*(0.5*c + 0.5*d) (lerp along x, then y)
  const float a = nested_mean(img, p, y0, x0, h), b = nested_mean(img, p, y0, x0 + h, h);
  const float c = nested_mean(img, p, y0 + h, x0, h), d = nested_mean(img, p, y0 + h, x0 + h, h);
  return 0.5f * (0.5f * a + 0.5f * b) + 0.5f * (0.5f * c + 0.5f * d);
}

__global__ void spatial_rescale_kernel(const float* __restrict__ seg, const float* __restrict__ w,
                                       float* __restrict__ out, int cin, int cout, int p, int n_stages) {
  const int l = p >> n_stages, f = 1 << n_stages;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= l * l) return;
  const int oy = i / l, ox = i % l;
  float m[8];
  for (int c = 0; c < cin; ++c)
    m[c] = nested_mean(seg + (static_cast<size_t>(b) * cin + c) * p * p, p, oy * f, ox * f, f);
  for (int o = 0; o < cout; ++o) {
    float acc = 0.f;
    for (int c = 0; c < cin; ++c) acc += w[o * cin + c] * m[c];
    out[(static_cast<size_t>(b) * cout + o) * l * l + i] = acc;
  }
}

extern "C" int stedm_spatial_rescale(const float* seg, const float* w, float* out, int batch, int cin, int cout, int p,
                                     int n_stages, void* stream) {
  STEDM_REQUIRE(seg && w && out && batch > 0 && cin > 0 && cin <= 8 && cout > 0, "spatial_rescale: bad argument");
  STEDM_REQUIRE(n_stages >= 0 && n_stages <= 4 && p % (1 << n_stages) == 0, "spatial_rescale: bad size");
  const int l = p >> n_stages;
  dim3 grid((l * l + 127) / 128, batch);
  spatial_rescale_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(seg, w, out, cin, cout, p, n_stages);
  return check_launch("spatial_rescale");
}

// =====================================================================================================
// predict_step tail: clip, (x+1)*127.5, truncate to uint8, NCHW -> NHWC.
// =====================================================================================================
__global__ void image_to_uint8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int c, int hw) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= hw) return;
  for (int ch = 0; ch < c; ++ch) {
    float v = img[(static_cast<size_t>(b) * c + ch) * hw + p];
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    v = __fmul_rn(__fadd_rn(v, 1.0f), 127.5f);  // numpy: (x + 1) * 127.5 in fp32, then astype(uint8) truncates
    out[(static_cast<size_t>(b) * hw + p) * c + ch] = static_cast<uint8_t>(static_cast<int>(v));
  }
}

extern "C" int stedm_image_to_uint8(const float* img, uint8_t* out, int batch, int c, int hw, void* stream) {
  STEDM_REQUIRE(img && out && batch > 0 && c > 0 && hw > 0, "image_to_uint8: bad argument");
  dim3 grid((hw + 255) / 256, batch);
  image_to_uint8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, out, c, hw);
  return check_launch("image_to_uint8");
}

// =====================================================================================================
// Row softmax (fp32, in place), one warp per row; used by the fp32-mode attention.
// =====================================================================================================
template <typename TO>
__global__ void __launch_bounds__(256) softmax_rows_kernel(float* __restrict__ x, TO* __restrict__ out, long long rows,
                                                           int cols, float scale, int mask_diag_period) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  float* r = x + row * cols;
  TO* o = out + row * cols;
  // mask_diag_period = T > 0: rows are queries of [.., T, cols = T] score matrices; a token does not attend to itself
  const int self_col = mask_diag_period > 0 ? static_cast<int>(row % mask_diag_period) : -1;
  float m = -INFINITY;
  for (int i = lane; i < cols; i += 32)
    if (i != self_col) m = fmaxf(m, r[i]);
  m = warp_max(m) * scale;
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) {
    const float e = i == self_col ? 0.f : expf(r[i] * scale - m);
    r[i] = e;                      // each lane re-reads only what it wrote
    s += e;
  }
  s = warp_sum(s);
  for (int i = lane; i < cols; i += 32) o[i] = from_f32<TO>(__fdiv_rn(r[i], s));
}

// Long rows (the VAE decoder's T = L^2 = 4096+ keys, model.py:189-190): one 256-thread block per row keeps the row
// in registers — one global read, one write — instead of the warp kernel's three passes over global memory.
template <typename TO, int PER>
__global__ void __launch_bounds__(256) softmax_rows_block_kernel(const float* __restrict__ x, TO* __restrict__ out,
                                                                 int cols, float scale, int mask_diag_period) {
  __shared__ float red[8];
  const long long row = blockIdx.x;
  const float* r = x + row * cols;
  const int self_col = mask_diag_period > 0 ? static_cast<int>(row % mask_diag_period) : -1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float v[PER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = (c < cols && c != self_col) ? r[c] * scale : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = expf(v[i] - m);   // exp(-inf) = 0 for masked / out-of-range columns
    s += v[i];
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  TO* o = out + row * cols;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < cols) o[c] = from_f32<TO>(__fdiv_rn(v[i], s));
  }
}

extern "C" int stedm_softmax_rows(float* x, void* out, int out_dtype, long long rows, int cols, float scale,
                                  int mask_diag_period, void* stream) {
  STEDM_REQUIRE(x && rows > 0 && cols > 0 && scale > 0.f, "softmax_rows: bad argument");
  STEDM_REQUIRE((rows + 7) / 8 < 0x7fffffffLL, "softmax_rows: too many rows");
  auto s = static_cast<cudaStream_t>(stream);
  if (cols >= 2048 && cols <= 8192 && rows < 0x7fffffffLL) {
    const unsigned g = static_cast<unsigned>(rows);
    const bool f32 = out == nullptr || out_dtype == DT_F32;
    float* of = out ? static_cast<float*>(out) : x;
    auto ob = static_cast<__nv_bfloat16*>(out);
    if (cols <= 4096) {
      if (f32) softmax_rows_block_kernel<float, 16><<<g, 256, 0, s>>>(x, of, cols, scale, mask_diag_period);
      else softmax_rows_block_kernel<__nv_bfloat16, 16><<<g, 256, 0, s>>>(x, ob, cols, scale, mask_diag_period);
    } else {
      if (f32) softmax_rows_block_kernel<float, 32><<<g, 256, 0, s>>>(x, of, cols, scale, mask_diag_period);
      else softmax_rows_block_kernel<__nv_bfloat16, 32><<<g, 256, 0, s>>>(x, ob, cols, scale, mask_diag_period);
    }
    return check_launch("softmax_rows");
  }
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (out == nullptr || out_dtype == DT_F32)
    softmax_rows_kernel<float><<<grid, 256, 0, s>>>(x, out ? static_cast<float*>(out) : x, rows, cols, scale,
                                                    mask_diag_period);
  else
    softmax_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, static_cast<__nv_bfloat16*>(out), rows, cols, scale,
                                                            mask_diag_period);
  return check_launch("softmax_rows");
}

// =====================================================================================================
// GEGLU (ldm/modules/attention.py:37-44): out = x * gelu(gate) with [x | gate] = the two halves of the projection's
// output row.  HBM-bound: 8 channels (16 / 32 bytes) per thread, exact erf GELU.
// =====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ in, T* __restrict__ out, long long items,
                                                    int f) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= items) return;
  const int per_row = f / 8;
  const long long row = idx / per_row;
  const int c = static_cast<int>(idx % per_row) * 8;
  const T* xp = in + row * 2 * f + c;
  const T* gp = xp + f;
  T* op = out + row * f + c;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = to_f32<T>(xp[j]), g = to_f32<T>(gp[j]);
    op[j] = from_f32<T>(x * gelu_erf(g));
  }
}

extern "C" int stedm_geglu(const void* in, void* out, int dtype, long long rows, int f, void* stream) {
  STEDM_REQUIRE(in && out && rows > 0 && f > 0 && f % 8 == 0, "geglu: bad argument");
  const long long items = rows * (f / 8);
  const long long blocks = (items + 255) / 256;
  STEDM_REQUIRE(blocks < (1LL << 31), "geglu: too large");
  auto s = static_cast<cudaStream_t>(stream);
  if (dtype == DT_BF16)
    geglu_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in),
                                                                             static_cast<__nv_bfloat16*>(out), items, f);
  else
    geglu_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, s>>>(static_cast<const float*>(in),
                                                                      static_cast<float*>(out), items, f);
  return check_launch("geglu");
}
