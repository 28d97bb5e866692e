// VAE decoder (AutoencoderKL.decode, ldm/models/autoencoder.py:88-91 -> Decoder.forward,
// ldm/modules/diffusionmodules/model.py:622-655) on NHWC bf16 activations:
//   * 3x3 convolutions and 1x1 convolutions run on the tcgen05 implicit-GEMM / GEMM kernel (gemm.cu),
//     bias and the residual add fused into the epilogue;
//   * GroupNorm(32, eps 1e-6): statistics from the producing conv / GEMM epilogue (standalone two-stage reduction for
//     conv_in's output only), normalise + SiLU -> bf16 as one HBM-bound pass; conv_in (K = 36) is a fused kernel here;
//   * Upsample.forward (nearest x2 + 3x3 conv) runs as four 2x2 phase convs on the low-resolution input with pre-summed
//     weights (4/9 of the FLOPs, no upsampled tensor); conv_out (N = 3) as a tap-response GEMM (N = 27) + 9-neighbour
//     gather-sum;
//   * the mid-block single-head attention (d = 512) materialises the score matrix through the GEMM kernel
//     (B200 has the HBM for it) with a row softmax in fp32.
#include "vae.cuh"

#include <cmath>
#include <cstdlib>

namespace ir {

static inline int div_up_l(long a, long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------------ conv_in
// post_quant_conv (1x1, zc->zc) folded into conv_in (3x3, zc->Cout, pad 1): the zero padding applies to the
// post_quant output, so out-of-image taps contribute nothing. z: (B, zc, H, W) fp32, scaled by in_scale first
// (the caller's `latents / scaling_factor`, test_scripts/inference.py:116,141). out: NHWC bf16.
// PQ = false: plain 3x3 conv of an NCHW fp32 image (Encoder.conv_in, model.py:456-460).
template <int ZC, bool PQ = true>
__global__ void __launch_bounds__(256) conv_in_kernel(const float* __restrict__ z, const float* __restrict__ pq_w,
                                                      const float* __restrict__ pq_b, const float* __restrict__ w_t,
                                                      const float* __restrict__ bias, bf16* __restrict__ out, int B,
                                                      int H, int W, int Cout, float in_scale) {
  const int tpp = Cout / 8;  // threads per pixel
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long pix = idx / tpp;
  if (pix >= (long)B * H * W) return;
  const int co = (int)(idx % tpp) * 8;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
  float acc[8];
  {
    const float4 b0 = *reinterpret_cast<const float4*>(bias + co);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + co + 4);
    acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w;
    acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    float zin[ZC], zq[ZC];
#pragma unroll
    for (int c = 0; c < ZC; ++c) zin[c] = z[(((long)b * ZC + c) * H + yy) * W + xx] * in_scale;
#pragma unroll
    for (int o = 0; o < ZC; ++o) {
      if (PQ) {
        float v = pq_b[o];
#pragma unroll
        for (int c = 0; c < ZC; ++c) v += pq_w[o * ZC + c] * zin[c];
        zq[o] = v;
      } else {
        zq[o] = zin[o];
      }
    }
#pragma unroll
    for (int c = 0; c < ZC; ++c) {
      const float* wr = w_t + (long)(tap * ZC + c) * Cout + co;
      const float4 w0 = *reinterpret_cast<const float4*>(wr);
      const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
      acc[0] += w0.x * zq[c]; acc[1] += w0.y * zq[c]; acc[2] += w0.z * zq[c]; acc[3] += w0.w * zq[c];
      acc[4] += w1.x * zq[c]; acc[5] += w1.y * zq[c]; acc[6] += w1.z * zq[c]; acc[7] += w1.w * zq[c];
    }
  }
  uint4 u = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                       pack_bf16x2(acc[6], acc[7]));
  *reinterpret_cast<uint4*>(out + pix * Cout + co) = u;
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// Stage 1: per (image, pixel chunk, group) partial sum / sum of squares, fixed reduction order (deterministic).
__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x, float* __restrict__ partial, int P,
                                                         int C, int chunk_px, int nchunks) {
  __shared__ float sm[256][4];
  const int n = blockIdx.y, chunk = blockIdx.x;
  const int tpp = C / 8;            // threads per pixel
  const int pps = 256 / tpp;        // pixels per block step
  const int cs = threadIdx.x % tpp, ps = threadIdx.x / tpp;
  const int p0 = chunk * chunk_px;
  const int p1 = min(P, p0 + chunk_px);
  pdl_wait();
  pdl_launch();
  float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
  const bf16* xb = x + (long)n * P * C + cs * 8;
  auto accum = [&](const uint4& u) {
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    s_lo += (a.x + a.y) + (b.x + b.y);
    q_lo += (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y);
    s_hi += (c.x + c.y) + (d.x + d.y);
    q_hi += (c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y);
  };
  int p = p0 + ps;
  constexpr int U = 8;  // independent 16 B loads in flight per thread
  for (; p + (U - 1) * pps < p1; p += U * pps) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) u[k] = *reinterpret_cast<const uint4*>(xb + (long)(p + k * pps) * C);
#pragma unroll
    for (int k = 0; k < U; ++k) accum(u[k]);
  }
  for (; p < p1; p += pps) accum(*reinterpret_cast<const uint4*>(xb + (long)p * C));
  sm[threadIdx.x][0] = s_lo;
  sm[threadIdx.x][1] = q_lo;
  sm[threadIdx.x][2] = s_hi;
  sm[threadIdx.x][3] = q_hi;
  __syncthreads();
  if (threadIdx.x < 64) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;  // which: 0 = sum, 1 = sum of squares
    const int cpg = C / 32;                                   // channels per group: 4, 8 or 16
    float acc = 0.f;
    for (int pp = 0; pp < pps; ++pp) {
      if (cpg == 4) {
        const int t = pp * tpp + (g >> 1);
        acc += sm[t][(g & 1) * 2 + which];
      } else {
        const int slots = cpg / 8;
        for (int k = 0; k < slots; ++k) {
          const int t = pp * tpp + g * slots + k;
          acc += sm[t][which] + sm[t][2 + which];
        }
      }
    }
    partial[(((long)n * nchunks + chunk) * 32 + g) * 2 + which] = acc;
  }
}

// Stage 2: (mean, rstd) per (image, group). One block per image, 8 warps x 4 groups; lanes stride over the chunk
// partials and are combined by a fixed shuffle tree in double precision (deterministic).
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ partial, float* __restrict__ stats,
                                                          int nchunks, double inv_count, float eps) {
  const int n = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int g = warp * 4 + k;
    double s = 0.0, q = 0.0;
    for (int c = lane; c < nchunks; c += 32) {
      const float2 pp = *reinterpret_cast<const float2*>(partial + (((long)n * nchunks + c) * 32 + g) * 2);
      s += (double)pp.x;
      q += (double)pp.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      const double mean = s * inv_count;
      double var = q * inv_count - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[((long)n * 32 + g) * 2] = (float)mean;
      stats[((long)n * 32 + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
}

// Finalize for the statistics fused into the GEMM / conv epilogue: partial[img][slot][32][2], slot = 128-row tile (x lane
// quarter for convs). grid = (images, NB): block j sums the slots [j * per, (j + 1) * per) of its image in double precision --
// a warp reads one slot's 32 groups x (sum, sumsq) = 256 contiguous bytes, the 8 warps stride over the slots and are
// combined through shared memory in warp order -- and parks its [32][2] doubles in scratch; the block that draws the last
// ticket of the image adds the NB block results in index order and writes (mean, rstd). Every order is fixed, so the
// statistics are bit-reproducible. (The one-block-per-(image, group) version read the partials with a 256 B stride: 25 us
// per full-resolution layer at 1024^2.) count[img] must be zero on entry and is left zero.
__global__ void __launch_bounds__(256) gn_finalize_fused_kernel(const float* __restrict__ partial, float* __restrict__ stats,
                                                                double* __restrict__ scratch, unsigned* __restrict__ count,
                                                                int nslots, int per, double inv_count, float eps) {
  __shared__ double red[8][64];
  __shared__ bool last;
  const int n = blockIdx.x, j = blockIdx.y, NB = gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch();
  double s = 0.0, q = 0.0;
  const int i0 = j * per, i1 = min(nslots, i0 + per);
  const float2* base = reinterpret_cast<const float2*>(partial) + (long)n * nslots * 32 + lane;
  for (int i = i0 + warp; i < i1; i += 64) {   // eight independent 256 B rows in flight per warp
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = (i + 8 * u < i1) ? __ldcg(base + (long)(i + 8 * u) * 32) : make_float2(0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s += (double)v[u].x;
      q += (double)v[u].y;
    }
  }
  red[warp][2 * lane] = s;
  red[warp][2 * lane + 1] = q;
  __syncthreads();
  if (threadIdx.x < 64) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
    if (NB == 1) red[0][threadIdx.x] = a; else scratch[((long)n * NB + j) * 64 + threadIdx.x] = a;
  }
  if (NB > 1) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(count + n, 1u) == (unsigned)(NB - 1));
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the NB block results, four quarters of the block list in parallel, each in index order, quarters combined in order
    const int v = threadIdx.x & 63, part = threadIdx.x >> 6;
    const int span = (NB + 3) / 4, b0 = part * span, b1 = min(NB, b0 + span);
    const double* sc = scratch + (long)n * NB * 64 + v;
    double a = 0.0;
    for (int b = b0; b < b1; b += 8) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = (b + u < b1) ? __ldcg(sc + (long)(b + u) * 64) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u) a += t[u];
    }
    __syncthreads();   // everyone is done with red[] of the first stage
    red[part][v] = a;
    __syncthreads();
    if (threadIdx.x < 64) red[4][threadIdx.x] = ((red[0][threadIdx.x] + red[1][threadIdx.x]) + red[2][threadIdx.x]) + red[3][threadIdx.x];
    if (threadIdx.x == 0) count[n] = 0u;
    __syncthreads();
    if (threadIdx.x < 64) red[0][threadIdx.x] = red[4][threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const double mean = red[0][2 * threadIdx.x] * inv_count;
    double var = red[0][2 * threadIdx.x + 1] * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[((long)n * 32 + threadIdx.x) * 2] = (float)mean;
    stats[((long)n * 32 + threadIdx.x) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// Stage 3: y = act((x - mean) * rstd * gamma + beta) -> bf16; act = SiLU (x * sigmoid(x), model.py:43-45) or none.
// grid = (blocks per image, images): the image index comes from blockIdx.y and a thread keeps its channel slot (the
// vectors-per-pixel count divides the block size), so the per-vector work has no integer division and the per-thread
// affine map  y = x * a + b  (a = rstd * gamma, b = beta - mean * a) is computed once. The grid is sized to the
// resident block count (3 per SM) and every thread keeps 8 independent 16 B loads in flight.
template <bool SILU>
__global__ void __launch_bounds__(256, 3) gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                       const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int vec_per_img, int C) {
  const int tpp = C / 8;   // 16-byte vectors per pixel; divides the block size, so a thread keeps its channel slot
  const int cpg = C / 32;
  const int c0 = (threadIdx.x % tpp) * 8;
  const int n = blockIdx.y;
  pdl_wait();
  pdl_launch();
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + c0), g1 = *reinterpret_cast<const float4*>(gamma + c0 + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + c0), b1 = *reinterpret_cast<const float4*>(beta + c0 + 4);
  const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const int g_lo = c0 / cpg, g_hi = (c0 + 4) / cpg;
  const float2 st_lo = *reinterpret_cast<const float2*>(stats + ((long)n * 32 + g_lo) * 2);
  const float2 st_hi = *reinterpret_cast<const float2*>(stats + ((long)n * 32 + g_hi) * 2);
  float ka[8], kb[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 st = k < 4 ? st_lo : st_hi;
    ka[k] = st.y * gm[k];
    kb[k] = bt[k] - st.x * ka[k];
  }
  const bf16* xi = x + (long)n * vec_per_img * 8;
  bf16* yi = y + (long)n * vec_per_img * 8;
  const int stride = gridDim.x * blockDim.x;
  auto one = [&](int i, const uint4& u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = unpack_bf16x2(w[k]);
      float t0 = fmaf(a.x, ka[2 * k], kb[2 * k]);
      float t1 = fmaf(a.y, ka[2 * k + 1], kb[2 * k + 1]);
      if (SILU) {
        t0 = silu_fast(t0);
        t1 = silu_fast(t1);
      }
      o[k] = pack_bf16x2(t0, t1);
    }
    *reinterpret_cast<uint4*>(yi + (long)i * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  };
  constexpr int U = 8;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < vec_per_img; i += U * stride) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) u[k] = *reinterpret_cast<const uint4*>(xi + (long)(i + k * stride) * 8);
#pragma unroll
    for (int k = 0; k < U; ++k) one(i + k * stride, u[k]);
  }
  for (; i < vec_per_img; i += stride) one(i, *reinterpret_cast<const uint4*>(xi + (long)i * 8));
}

// ------------------------------------------------------------------------------------------------ upsample
// F.interpolate(scale_factor=2.0, mode="nearest") on NHWC bf16 (Upsample.forward, model.py:63-67).
__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long total_vec,
                                                         int H, int W, int C) {
  // total_vec counts INPUT vectors: every 16 B input vector is read once and written to its 2x2 output pixels
  const int tpp = C / 8;
  const long stride = (long)gridDim.x * blockDim.x;
  constexpr int U = 4;
  auto emit = [&](long i, const uint4& u) {
    const int cv = (int)(i % tpp);
    long pix = i / tpp;
    const int xi = (int)(pix % W);
    pix /= W;
    const int yi = (int)(pix % H);
    const long n = pix / H;
    bf16* o = y + (((n * 2 * H + 2 * yi) * 2 * W + 2 * xi) * (long)C) + cv * 8;
    *reinterpret_cast<uint4*>(o) = u;
    *reinterpret_cast<uint4*>(o + C) = u;
    *reinterpret_cast<uint4*>(o + 2L * W * C) = u;
    *reinterpret_cast<uint4*>(o + 2L * W * C + C) = u;
  };
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < total_vec; i += U * stride) {
    uint4 u[U];
#pragma unroll
    for (int k = 0; k < U; ++k) u[k] = *reinterpret_cast<const uint4*>(x + (i + k * stride) * 8);
#pragma unroll
    for (int k = 0; k < U; ++k) emit(i + k * stride, u[k]);
  }
  for (; i < total_vec; i += stride) emit(i, *reinterpret_cast<const uint4*>(x + i * 8));
}

// ------------------------------------------------------------------------------------------------ attention helpers
// out[c][p] = in[p][col0 + c] (V^T for the P*V GEMM whose B operand must be K-major).
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int P, int C, long ld_in,
                                      int col0) {
  __shared__ bf16 tile[32][34];
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const bf16* src = in + (long)blockIdx.z * P * ld_in;
  bf16* dst = out + (long)blockIdx.z * P * C;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int p = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < P && c < C) ? src[(long)p * ld_in + col0 + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, p = p0 + threadIdx.x;
    if (c < C && p < P) dst[(long)c * P + p] = tile[threadIdx.x][r];
  }
}

// AttnBlock softmax (model.py:195-197) without a materialised fp32 score matrix: the q k^T GEMM runs twice on the
// EPI_ATTN epilogue -- pass 1 keeps only the maximum of every (row, 64-key group), pass 2 writes exp2(alpha s - shift_row)
// as bf16 and the fp32 sum of every (row, group) -- and the P V GEMM scales its rows by 1 / sum. These are the reductions
// between the passes: part[group * n + row] -> per-row value, fixed order (deterministic).
// mode 0: out[row] = mul * max over groups (the pass-2 shift, mul = C^-1/2 * log2 e > 0); mode 1: out[row] = 1 / sum.
__global__ void __launch_bounds__(256) attn_row_reduce_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                              int ngroups, int n, float mul, int mode) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  pdl_launch();
  if (row >= n) return;
  if (mode == 0) {
    float m = -INFINITY;
    for (int g = 0; g < ngroups; ++g) m = fmaxf(m, part[(long)g * n + row]);
    out[row] = mul * m;
  } else {
    float a = 0.f;
    for (int g = 0; g < ngroups; ++g) a += part[(long)g * n + row];
    out[row] = 1.0f / a;
  }
}

// ------------------------------------------------------------------------------------------------ conv_out
// conv_out (3x3, C -> 3, pad 1) on the normalised + SiLU'd NHWC bf16 activation; writes NCHW fp32 with a fused affine
// out = conv * out_scale + out_shift (the caller's `/2 + 0.5`, test_scripts/inference.py:117,142). HBM-bound (N = 3).
template <int C>
__global__ void __launch_bounds__(256) conv_out_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ out, int H,
                                                       int W, float out_scale, float out_shift) {
  constexpr int TW = 32, TH = 8, HW_ = TW + 2, HH_ = TH + 2;
  constexpr int LDP = C + 8;  // pixel stride in elements: (C*2 + 16) B keeps 16 B loads conflict-free
  extern __shared__ __align__(16) uint8_t smem_co[];
  bf16* tile = reinterpret_cast<bf16*>(smem_co);                    // [HH_*HW_][LDP]
  float* ws = reinterpret_cast<float*>(tile + HH_ * HW_ * LDP);     // [3][9][C]
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  for (int i = threadIdx.x; i < 3 * 9 * C; i += 256) ws[i] = w[i];
  constexpr int VPP = C / 8;
  for (int i = threadIdx.x; i < HH_ * HW_ * VPP; i += 256) {
    const int v = i % VPP, pp = i / VPP;
    const int yy = y0 + pp / HW_ - 1, xx = x0 + pp % HW_ - 1;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      u = *reinterpret_cast<const uint4*>(x + (((long)n * H + yy) * W + xx) * C + v * 8);
    *reinterpret_cast<uint4*>(tile + pp * LDP + v * 8) = u;
  }
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int ox = x0 + tx, oy = y0 + ty;
  float acc[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const bf16* px = tile + ((ty + tap / 3) * HW_ + tx + tap % 3) * LDP;
    const float* w0 = ws + (0 * 9 + tap) * C;
    const float* w1 = ws + (1 * 9 + tap) * C;
    const float* w2 = ws + (2 * 9 + tap) * C;
#pragma unroll 4
    for (int v = 0; v < VPP; ++v) {
      const uint4 u = *reinterpret_cast<const uint4*>(px + v * 8);
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      const float4* wv[3] = {reinterpret_cast<const float4*>(w0 + v * 8), reinterpret_cast<const float4*>(w1 + v * 8),
                             reinterpret_cast<const float4*>(w2 + v * 8)};
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float4 wa = wv[o][0], wb = wv[o][1];
        acc[o] += wa.x * a.x + wa.y * a.y + wa.z * b.x + wa.w * b.y + wb.x * c.x + wb.y * c.y + wb.z * d.x + wb.w * d.y;
      }
    }
  }
  if (ox < W && oy < H) {
#pragma unroll
    for (int o = 0; o < 3; ++o) out[(((long)n * 3 + o) * H + oy) * W + ox] = acc[o] * out_scale + out_shift;
  }
}

// conv_out on the tensor cores, without a halo: a plain GEMM over the channels gives every pixel's 27 tap responses
// Y[pixel][tap*3 + o] = sum_c w[o][c][tap] * x[pixel][c] (the activation is read ONCE; the implicit-GEMM conv would pull
// it through L2 three times for 3 useful output columns), and this kernel sums the 9 neighbours' responses:
// out[o](y, x) = bias[o] + sum_tap Y[(y + ky - 1, x + kx - 1)][tap*3 + o], out-of-image neighbours contribute 0 (zero
// padding), then the caller's affine; NCHW fp32. 32 x 8 output pixels per block, (34 x 10) x 27 responses in shared memory.
__global__ void __launch_bounds__(256) conv_out_gather_kernel(const float* __restrict__ Y, const float* __restrict__ bias,
                                                              float* __restrict__ out, int H, int W, float out_scale,
                                                              float out_shift) {
  constexpr int TW = 32, TH = 8, HW_ = TW + 2, HH_ = TH + 2, NR = 27, LDY = 32;
  __shared__ float sm[HH_ * HW_ * NR];
  const int n = blockIdx.z;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  pdl_wait();
  pdl_launch();
  for (int i = threadIdx.x; i < HH_ * HW_ * (LDY / 4); i += 256) {
    const int q = i % (LDY / 4), pp = i / (LDY / 4);
    const int yy = y0 + pp / HW_ - 1, xx = x0 + pp % HW_ - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < H && xx >= 0 && xx < W)
      v = *reinterpret_cast<const float4*>(Y + (((long)n * H + yy) * W + xx) * LDY + q * 4);
    float* d = sm + pp * NR + q * 4;
    if (q * 4 + 0 < NR) d[0] = v.x;
    if (q * 4 + 1 < NR) d[1] = v.y;
    if (q * 4 + 2 < NR) d[2] = v.z;
    if (q * 4 + 3 < NR) d[3] = v.w;
  }
  __syncthreads();
  const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;
  const int ox = x0 + tx, oy = y0 + ty;
  float acc[3] = {bias[0], bias[1], bias[2]};
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float* r = sm + ((ty + tap / 3) * HW_ + tx + tap % 3) * NR + tap * 3;   // stride 27 floats: conflict-free
    acc[0] += r[0];
    acc[1] += r[1];
    acc[2] += r[2];
  }
  if (ox < W && oy < H) {
#pragma unroll
    for (int o = 0; o < 3; ++o) out[(((long)n * 3 + o) * H + oy) * W + ox] = acc[o] * out_scale + out_shift;
  }
}

// Upsample.forward (model.py:63-67) = nearest x2 + 3x3 conv. On the low-resolution input s that is, per output phase
// (a, b) = (Y & 1, X & 1), a 2x2 conv: out(2y+a, 2x+b) = sum_{t,u} W_ab[t][u] . s(y + a - 1 + t, x + b - 1 + u), where
// W_ab[t][u] sums the 3x3 taps that land on the same source pixel: rows {0 | 1,2} for a = 0 and {0,1 | 2} for a = 1 (same
// for columns) -- 4/9 of the FLOPs, no upsampled tensor. Zero padding carries over: the out-of-image taps of the 3x3 conv
// are exactly the out-of-image source pixels of the 2x2 one.
// (Cout, Cin, 3, 3) fp32 -> [phase = 2a + b][Cout][t][u][Cin] bf16 (sums in fp32, one rounding)
__global__ void pack_upconv_phases_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int cin) {
  const long per_phase = (long)cout * 4 * cin;
  const long total = 4 * per_phase;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int u = (int)((i / cin) % 2);
    const int t = (int)((i / (2L * cin)) % 2);
    const int o = (int)((i / (4L * cin)) % cout);
    const int phase = (int)(i / per_phase);
    const int a = phase >> 1, b = phase & 1;
    // taps of the 3x3 kernel that fall on source offset t (rows) / u (columns)
    const int ky0 = a == 0 ? (t == 0 ? 0 : 1) : (t == 0 ? 0 : 2), ky1 = a == 0 ? (t == 0 ? 0 : 2) : (t == 0 ? 1 : 2);
    const int kx0 = b == 0 ? (u == 0 ? 0 : 1) : (u == 0 ? 0 : 2), kx1 = b == 0 ? (u == 0 ? 0 : 2) : (u == 0 ? 1 : 2);
    const float* w = src + ((long)o * cin + c) * 9;
    float acc = 0.f;
    for (int ky = ky0; ky <= ky1; ++ky)
      for (int kx = kx0; kx <= kx1; ++kx) acc += w[ky * 3 + kx];
    dst[i] = __float2bfloat16(acc);
  }
}

// (3, Cin, 3, 3) fp32 -> [tap*3 + o][c] bf16 (conv_out as a tap-response GEMM; rows 27..31 of the buffer stay zero)
__global__ void pack_convout_taps_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int cin) {
  const long total = (long)cout * cin * 9;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int o = (int)((i / cin) % cout);
    const int tap = (int)(i / ((long)cin * cout));
    dst[i] = __float2bfloat16(src[((long)o * cin + c) * 9 + tap]);
  }
}

// ------------------------------------------------------------------------------------------------ encoder helpers
// Encoder.conv_in (model.py:456-460: 3x3, 3 -> ch, pad 1) on the tensor cores: the 27 taps of the 3-channel NCHW fp32 image
// are gathered into the K-major A operand of a plain GEMM, A[(b,y,x)][(ky*3+kx)*3 + c] = x[b][c][y+ky-1][x+kx-1] (zero outside
// the image; columns 27..31 zero), bf16, 64 B per pixel. One thread per pixel.
__global__ void __launch_bounds__(256) im2col_in_kernel(const float* __restrict__ x, bf16* __restrict__ a, int H, int W,
                                                        long total_px) {
  const long pix = blockIdx.x * (long)blockDim.x + threadIdx.x;
  pdl_wait();
  pdl_launch();
  if (pix >= total_px) return;
  const int xx = (int)(pix % W), yy = (int)((pix / W) % H);
  const long b = pix / ((long)W * H);
  const float* img = x + b * 3 * (long)H * W;
  float v[32];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
    const bool in = sy >= 0 && sy < H && sx >= 0 && sx < W;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[tap * 3 + c] = in ? __ldg(img + ((long)c * H + sy) * W + sx) : 0.f;
  }
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(a + pix * 32);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    dst[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                        pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
}

// quant_conv (1x1, 2z -> 2z, autoencoder.py:84) on the fp32 NHWC output of Encoder.conv_out, written as the NCHW
// fp32 `moments` tensor (mean = channels [0, z), logvar = channels [z, 2z); distributions.py:27).
template <int CZ>
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ h, const float* __restrict__ qw,
                                                      const float* __restrict__ qb, float* __restrict__ out, long P_total,
                                                      int P) {
  const long pix = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (pix >= P_total) return;
  float in[CZ];
#pragma unroll
  for (int c = 0; c < CZ; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(h + pix * CZ + c);
    in[c] = v.x; in[c + 1] = v.y; in[c + 2] = v.z; in[c + 3] = v.w;
  }
  const long b = pix / P, p = pix % P;
#pragma unroll
  for (int o = 0; o < CZ; ++o) {
    float v = qb[o];
#pragma unroll
    for (int c = 0; c < CZ; ++c) v += qw[o * CZ + c] * in[c];
    out[(b * CZ + o) * P + p] = v;
  }
}

// ================================================================================================ handle
static long align64(long v) { return (v + 63) / 64 * 64; }

static void vae_add(Vae* v, const std::string& name, int kind, long numel, int cout, int cin, int k) {
  VaeParam p;
  p.name = name;
  p.kind = kind;
  p.numel = numel;
  p.cout = cout;
  p.cin = cin;
  p.k = k;
  if (kind == VP_CONV_BF16) {
    p.offset = v->wb_elems;
    v->wb_elems = align64(v->wb_elems + numel);
  } else {
    p.offset = v->wf_elems;
    v->wf_elems = align64(v->wf_elems + numel);
  }
  v->index[name] = (int)v->params.size();
  v->params.push_back(p);
}

static void vae_add_conv(Vae* v, const std::string& name, int cout, int cin, int k, int kind = VP_CONV_BF16) {
  vae_add(v, name + ".weight", kind, (long)cout * cin * k * k, cout, cin, k);
  vae_add(v, name + ".bias", VP_F32, cout, cout, 1, 1);
}
static void vae_add_norm(Vae* v, const std::string& name, int c) {
  vae_add(v, name + ".weight", VP_F32, c, c, 1, 1);
  vae_add(v, name + ".bias", VP_F32, c, c, 1, 1);
}
static void vae_add_res(Vae* v, const std::string& name, int cin, int cout) {
  vae_add_norm(v, name + ".norm1", cin);
  vae_add_conv(v, name + ".conv1", cout, cin, 3);
  vae_add_norm(v, name + ".norm2", cout);
  vae_add_conv(v, name + ".conv2", cout, cout, 3);
  if (cin != cout) vae_add_conv(v, name + ".nin_shortcut", cout, cin, 1);
}

int vae_create(const VaeConfig& cfg, Vae** out) {
  IR_REQUIRE(cfg.z_channels == 4 && cfg.out_ch == 3, "vae: z_channels 4 and 3 output channels expected");
  IR_REQUIRE(cfg.ch % 64 == 0, "vae: ch must be a multiple of 64");
  Vae* v = new Vae();
  v->cfg = cfg;
  const std::string d = "decoder";
  int block_in = cfg.ch * cfg.ch_mult[3];
  vae_add_conv(v, "post_quant_conv", cfg.z_channels, cfg.z_channels, 1, VP_F32);
  vae_add_conv(v, d + ".conv_in", block_in, cfg.z_channels, 3, VP_CONVIN_F32);
  vae_add_res(v, d + ".mid.block_1", block_in, block_in);
  vae_add_norm(v, d + ".mid.attn_1.norm", block_in);
  // q, k, v weights are registered back to back so that one GEMM with N = 3C produces all three
  vae_add(v, d + ".mid.attn_1.q.weight", VP_CONV_BF16, (long)block_in * block_in, block_in, block_in, 1);
  vae_add(v, d + ".mid.attn_1.k.weight", VP_CONV_BF16, (long)block_in * block_in, block_in, block_in, 1);
  vae_add(v, d + ".mid.attn_1.v.weight", VP_CONV_BF16, (long)block_in * block_in, block_in, block_in, 1);
  vae_add(v, d + ".mid.attn_1.q.bias", VP_F32, block_in, block_in, 1, 1);
  vae_add(v, d + ".mid.attn_1.k.bias", VP_F32, block_in, block_in, 1, 1);
  vae_add(v, d + ".mid.attn_1.v.bias", VP_F32, block_in, block_in, 1, 1);
  vae_add_conv(v, d + ".mid.attn_1.proj_out", block_in, block_in, 1);
  vae_add_res(v, d + ".mid.block_2", block_in, block_in);
  for (int lvl = 3; lvl >= 0; --lvl) {
    const int block_out = cfg.ch * cfg.ch_mult[lvl];
    for (int b = 0; b < cfg.num_res_blocks + 1; ++b) {
      vae_add_res(v, d + ".up." + std::to_string(lvl) + ".block." + std::to_string(b), block_in, block_out);
      block_in = block_out;
    }
    if (lvl != 0) {
      const std::string un = d + ".up." + std::to_string(lvl) + ".upsample.conv";
      vae_add_conv(v, un, block_in, block_in, 3);
      bf16* pw = nullptr;
      if (cudaMalloc(&pw, (size_t)16 * block_in * block_in * sizeof(bf16)) != cudaSuccess) {
        set_last_error("vae_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        vae_destroy(v);
        return IR_ERR_CUDA;
      }
      v->up_w[un + ".weight"] = pw;
    }
  }
  vae_add_norm(v, d + ".norm_out", block_in);
  vae_add_conv(v, d + ".conv_out", cfg.out_ch, block_in, 3, VP_CONVOUT_F32);
  if (cfg.with_encoder) {
    // Encoder (model.py:440-546) + quant_conv (autoencoder.py:84), reference key names
    const std::string e = "encoder";
    v->first_encoder_param = (int)v->params.size();
    vae_add_conv(v, e + ".conv_in", cfg.ch, 3, 3, VP_CONVIN_F32);
    int cin = cfg.ch;
    for (int lvl = 0; lvl < 4; ++lvl) {
      const int cout = cfg.ch * cfg.ch_mult[lvl];
      for (int b = 0; b < cfg.num_res_blocks; ++b) {
        vae_add_res(v, e + ".down." + std::to_string(lvl) + ".block." + std::to_string(b), cin, cout);
        cin = cout;
      }
      if (lvl != 3) vae_add_conv(v, e + ".down." + std::to_string(lvl) + ".downsample.conv", cin, cin, 3);
    }
    vae_add_res(v, e + ".mid.block_1", cin, cin);
    vae_add_norm(v, e + ".mid.attn_1.norm", cin);
    vae_add(v, e + ".mid.attn_1.q.weight", VP_CONV_BF16, (long)cin * cin, cin, cin, 1);
    vae_add(v, e + ".mid.attn_1.k.weight", VP_CONV_BF16, (long)cin * cin, cin, cin, 1);
    vae_add(v, e + ".mid.attn_1.v.weight", VP_CONV_BF16, (long)cin * cin, cin, cin, 1);
    vae_add(v, e + ".mid.attn_1.q.bias", VP_F32, cin, cin, 1, 1);
    vae_add(v, e + ".mid.attn_1.k.bias", VP_F32, cin, cin, 1, 1);
    vae_add(v, e + ".mid.attn_1.v.bias", VP_F32, cin, cin, 1, 1);
    vae_add_conv(v, e + ".mid.attn_1.proj_out", cin, cin, 1);
    vae_add_res(v, e + ".mid.block_2", cin, cin);
    vae_add_norm(v, e + ".norm_out", cin);
    vae_add_conv(v, e + ".conv_out", 2 * cfg.z_channels, cin, 3);
    vae_add_conv(v, "quant_conv", 2 * cfg.z_channels, 2 * cfg.z_channels, 1, VP_F32);
  }
  if (cfg.with_encoder) {
    const size_t ci_w_bytes = (size_t)cfg.ch * 32 * sizeof(bf16);
    if (cudaMalloc(&v->ci_w, ci_w_bytes) != cudaSuccess || cudaMemset(v->ci_w, 0, ci_w_bytes) != cudaSuccess) {
      set_last_error("vae_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
      vae_destroy(v);
      return IR_ERR_CUDA;
    }
  }
  const size_t co_w_bytes = (size_t)32 * (cfg.ch * cfg.ch_mult[0]) * sizeof(bf16);   // 27 tap-response rows + 5 zero rows
  if (cudaMalloc(&v->wb, (size_t)v->wb_elems * sizeof(bf16)) != cudaSuccess ||
      cudaMalloc(&v->wf, (size_t)v->wf_elems * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&v->co_w, co_w_bytes) != cudaSuccess || cudaMalloc(&v->co_b, 4 * sizeof(float)) != cudaSuccess ||
      cudaMemset(v->co_w, 0, co_w_bytes) != cudaSuccess || cudaMemset(v->co_b, 0, 4 * sizeof(float)) != cudaSuccess) {
    set_last_error("vae_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    vae_destroy(v);
    return IR_ERR_CUDA;
  }
  *out = v;
  return IR_OK;
}

void vae_release_graphs(Vae* v);

void vae_destroy(Vae* v) {
  if (!v) return;
  vae_release_graphs(v);
  cudaFree(v->wb);
  cudaFree(v->wf);
  cudaFree(v->co_w);
  cudaFree(v->co_b);
  cudaFree(v->ci_w);
  for (auto& kv : v->up_w) cudaFree(kv.second);
  delete v;
}

// (Cout, Cin, k, k) fp32 -> (Cout, k*k*Cin) bf16, tap-major K (k = (ky*3 + kx)*Cin + c): the layout of the implicit GEMM
__global__ void pack_conv_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int cin, int kk) {
  const long total = (long)cout * cin * kk;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int tap = (int)((i / cin) % kk);
    const int o = (int)(i / ((long)cin * kk));
    dst[i] = __float2bfloat16(src[((long)o * cin + c) * kk + tap]);
  }
}
// (Cout, Cin, 3, 3) fp32 -> [tap*Cin + c][Cout] fp32 (conv_in: coalesced over Cout)
__global__ void pack_convin_kernel(const float* __restrict__ src, float* __restrict__ dst, int cout, int cin) {
  const long total = (long)cout * cin * 9;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int o = (int)(i % cout);
    const int c = (int)((i / cout) % cin);
    const int tap = (int)(i / ((long)cout * cin));
    dst[i] = src[((long)o * cin + c) * 9 + tap];
  }
}
// encoder.conv_in (Cout, 3, 3, 3) fp32 -> (Cout, 32) bf16, column (ky*3+kx)*3 + c (columns 27..31 stay zero): the W operand
// of the im2col GEMM
__global__ void pack_convin_taps_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cout * 27) return;
  const int o = i / 27, r = i - o * 27;
  const int tap = r / 3, c = r - tap * 3;
  dst[o * 32 + r] = __float2bfloat16(src[((long)o * 3 + c) * 9 + tap]);
}
// (3, Cin, 3, 3) fp32 -> [o][tap][c] fp32 (conv_out)
__global__ void pack_convout_kernel(const float* __restrict__ src, float* __restrict__ dst, int cout, int cin) {
  const long total = (long)cout * cin * 9;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int tap = (int)((i / cin) % 9);
    const int o = (int)(i / ((long)cin * 9));
    dst[i] = src[((long)o * cin + c) * 9 + tap];
  }
}

int pack_upconv_phases(const float* w_oihw, bf16* phase_w, int cout, int cin, cudaStream_t s) {
  pack_upconv_phases_kernel<<<(int)((16L * cout * cin + 255) / 256), 256, 0, s>>>(w_oihw, phase_w, cout, cin);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

int upsample_conv_phases_launch(const bf16* x, const bf16* phase_w, const float* bias, bf16* y, int n, int H, int W, int C,
                                float* gn_partial, int force_bn, cudaStream_t s) {
  const int tiles = gemm_conv_tiles_per_image(H, W);
  for (int phase = 0; phase < 4; ++phase) {
    const int a = phase >> 1, b = phase & 1;
    GemmArgs g;
    g.A = x;
    g.W = phase_w + (long)phase * C * 4 * C;
    g.ldw = 4L * C;
    g.M = n * H * W;
    g.N = C;
    g.K = 4 * C;
    g.conv = 1;
    g.conv_taps = 2;
    g.conv_off_y = a - 1;
    g.conv_off_x = b - 1;
    g.o_scale = 2;
    g.o_oy = a;
    g.o_ox = b;
    g.nimg = n;
    g.H = H;
    g.Wd = W;
    g.C = C;
    g.epi = EPI_BF16;
    g.bias = bias;
    g.out_bf16 = y;
    g.ldo_b = C;
    g.force_bn = force_bn;
    if (gn_partial) {
      g.gn_partial = gn_partial;
      g.gn_cpg = C / 32;
      g.gn_slot_off = phase * tiles;
      g.gn_slots_img = 4 * tiles;
    }
    IR_TRY(gemm_launch(g, s));
  }
  return IR_OK;
}

int vae_load_param(Vae* v, const char* name, const float* src, long numel, cudaStream_t s) {
  auto it = v->index.find(name);
  if (it == v->index.end()) {
    set_last_error("vae_load_param: unknown parameter '%s'", name);
    return IR_ERR_INVALID;
  }
  VaeParam& p = v->params[it->second];
  IR_REQUIRE(numel == p.numel, "vae_load_param: '%s' has %ld elements, expected %ld", name, numel, p.numel);
  const int grid = div_up_l(numel, 256);
  switch (p.kind) {
    case VP_CONV_BF16: {
      pack_conv_bf16_kernel<<<grid, 256, 0, s>>>(src, v->wb + p.offset, p.cout, p.cin, p.k * p.k);
      auto up = v->up_w.find(name);
      if (up != v->up_w.end())
        pack_upconv_phases_kernel<<<div_up_l(16L * p.cout * p.cin, 256), 256, 0, s>>>(src, up->second, p.cout, p.cin);
      break;
    }
    case VP_CONVIN_F32:
      pack_convin_kernel<<<grid, 256, 0, s>>>(src, v->wf + p.offset, p.cout, p.cin);
      if (p.name == "encoder.conv_in.weight")
        pack_convin_taps_kernel<<<div_up_l(27L * p.cout, 256), 256, 0, s>>>(src, v->ci_w, p.cout);
      break;
    case VP_CONVOUT_F32:
      pack_convout_kernel<<<grid, 256, 0, s>>>(src, v->wf + p.offset, p.cout, p.cin);
      // tensor-core path: 27 tap-response rows [tap*3 + o][c]
      pack_convout_taps_kernel<<<grid, 256, 0, s>>>(src, v->co_w, p.cout, p.cin);
      break;
    default:
      IR_CUDA_CHECK(cudaMemcpyAsync(v->wf + p.offset, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
      if (p.name == "decoder.conv_out.bias")
        IR_CUDA_CHECK(cudaMemcpyAsync(v->co_b, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  IR_CUDA_CHECK(cudaGetLastError());
  p.loaded = true;
  return IR_OK;
}

// ================================================================================================ decode
struct VaeWs {
  bf16* buf[4];
  bf16 *qkv, *vt, *pm;
  float *att_part, *att_row, *partial, *stats;   // att_part: [ceil(P/64)][P] group maxima / sums; att_row: [2][P]
  long partial_elems;
  double* fin_scratch = nullptr;   // gn_finalize_fused: [B][GN_FIN_BLOCKS][32][2] block sums
  unsigned* fin_count = nullptr;   // [B] tickets (zeroed at the start of every decode / encode call)
  bf16* im2col = nullptr;   // encoder only: A operand of conv_in ([B*H*W][32]: 27 taps of the 3-channel image + 5 zeros)
  float* f32tmp = nullptr;  // encoder only: fp32 NHWC output of conv_out
  float *g_in = nullptr, *g_out = nullptr;   // CUDA-graph replay: the call's input / output staged at fixed addresses
};

static const int GN_MAX_CHUNKS = 2048;
static const int GN_FIN_BLOCKS = 256;

static size_t vae_carve(const Vae* v, VaeWs& w, void* base, int B, int h, int wd) {
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* {
    off = (off + 255) & ~size_t(255);
    void* p = b ? b + off : nullptr;
    off += bytes;
    return p;
  };
  // largest activation over the decoder schedule (pixels x channels)
  long cmax = 0;
  {
    long Hc = h, Wc = wd, C = (long)v->cfg.ch * v->cfg.ch_mult[3];
    cmax = (long)B * Hc * Wc * C;
    for (int lvl = 3; lvl >= 0; --lvl) {
      const long Cout = (long)v->cfg.ch * v->cfg.ch_mult[lvl];
      cmax = std::max(cmax, (long)B * Hc * Wc * std::max(C, Cout));
      C = Cout;
      if (lvl != 0) {
        Hc *= 2;
        Wc *= 2;
        cmax = std::max(cmax, (long)B * Hc * Wc * C);
      }
    }
  }
  for (int i = 0; i < 4; ++i) w.buf[i] = reinterpret_cast<bf16*>(take((size_t)cmax * sizeof(bf16)));
  const long P = (long)h * wd, C = (long)v->cfg.ch * v->cfg.ch_mult[3];
  w.qkv = reinterpret_cast<bf16*>(take((size_t)B * P * 3 * C * sizeof(bf16)));
  w.vt = reinterpret_cast<bf16*>(take((size_t)P * C * sizeof(bf16)));
  w.pm = reinterpret_cast<bf16*>(take((size_t)P * P * sizeof(bf16)));
  w.att_part = reinterpret_cast<float*>(take((size_t)((P + 63) / 64) * P * sizeof(float)));
  w.att_row = reinterpret_cast<float*>(take((size_t)2 * P * sizeof(float)));
  // GroupNorm partials: standalone pass (<= GN_MAX_CHUNKS chunks) or fused (4 warps x 128-pixel tiles at full resolution)
  w.partial_elems = (long)B * std::max<long>(GN_MAX_CHUNKS, (long)gemm_conv_tiles_per_image(8 * h, 8 * wd) * gemm_conv_gn_slots_per_tile()) * 64;
  w.partial = reinterpret_cast<float*>(take((size_t)w.partial_elems * sizeof(float)));
  w.stats = reinterpret_cast<float*>(take((size_t)B * 32 * 2 * sizeof(float)));
  w.fin_scratch = reinterpret_cast<double*>(take((size_t)B * GN_FIN_BLOCKS * 64 * sizeof(double)));
  w.fin_count = reinterpret_cast<unsigned*>(take((size_t)B * sizeof(unsigned)));
  w.g_in = reinterpret_cast<float*>(take((size_t)B * v->cfg.z_channels * h * wd * sizeof(float)));
  w.g_out = reinterpret_cast<float*>(take((size_t)B * v->cfg.out_ch * 64 * h * wd * sizeof(float)));
  return (off + 255) & ~size_t(255);
}

size_t vae_workspace_bytes(const Vae* v, int B, int h, int w) {
  VaeWs ws;
  return vae_carve(v, ws, nullptr, B, h, w);
}

struct VCtx {
  Vae* v;
  VaeWs w;
  int B;
  cudaStream_t s;
  bool stats_ready = false;   // c.w.stats already holds (mean, rstd) of the tensor the next group_norm will see
};

template <typename T>
static const T* vp(const Vae* v, const std::string& name) {
  auto it = v->index.find(name);
  if (it == v->index.end()) return nullptr;
  const VaeParam& p = v->params[it->second];
  if (p.kind == VP_CONV_BF16) return reinterpret_cast<const T*>(v->wb + p.offset);
  return reinterpret_cast<const T*>(v->wf + p.offset);
}

// y = act(GroupNorm(x)); x, y: (B, P, C) NHWC bf16
static int group_norm(VCtx& c, const std::string& name, const bf16* x, bf16* y, int P, int C, bool silu_act) {
  IR_REQUIRE(C % 32 == 0 && (C / 32 == 4 || C / 32 == 8 || C / 32 == 16), "group_norm: C=%d unsupported", C);
  if (c.stats_ready) {
    c.stats_ready = false;   // statistics were produced by the epilogue of the kernel that wrote x
  } else {
    // The chunking must depend on the image size only (never on the batch): the fp32 partial sums are then grouped
    // identically however tiles are batched or sharded across GPUs, which keeps tiled restoration bit-reproducible.
    int chunk_px = 256;
    while (div_up_l(P, chunk_px) > 1024) chunk_px *= 2;
    const int nchunks = div_up_l(P, chunk_px);
    IR_REQUIRE(nchunks <= GN_MAX_CHUNKS, "group_norm: too many chunks");
    IR_CUDA_CHECK(launch_pdl(gn_partial_kernel, dim3(nchunks, c.B), dim3(256), 0, c.s, x, c.w.partial, P, C, chunk_px, nchunks));
    IR_CUDA_CHECK(launch_pdl(gn_finalize_kernel, dim3(c.B), dim3(256), 0, c.s, (const float*)c.w.partial, c.w.stats, nchunks,
                             1.0 / ((double)P * (C / 32)), 1e-6f));
    count_launch(2);
  }
  const long vec_per_img_l = (long)P * C / 8;
  IR_REQUIRE(vec_per_img_l < (1L << 30) && 256 % (C / 8) == 0, "group_norm: image too large / channel count unsupported");
  const int vec_per_img = (int)vec_per_img_l;
  // resident blocks (3 per SM) split over the images of the batch; never more blocks than 8-vector work items
  int gx = (device_num_sms() * 3 + c.B - 1) / c.B;
  const int need = div_up_l(vec_per_img, 256 * 8);
  if (gx > need) gx = need;
  if (gx < 1) gx = 1;
  const float* gamma = vp<float>(c.v, name + ".weight");
  const float* beta = vp<float>(c.v, name + ".bias");
  if (silu_act)
    IR_CUDA_CHECK(launch_pdl(gn_apply_kernel<true>, dim3(gx, c.B), dim3(256), 0, c.s, x, y, (const float*)c.w.stats, gamma, beta,
                             vec_per_img, C));
  else
    IR_CUDA_CHECK(launch_pdl(gn_apply_kernel<false>, dim3(gx, c.B), dim3(256), 0, c.s, x, y, (const float*)c.w.stats, gamma, beta,
                             vec_per_img, C));
  count_launch(3);
  return IR_OK;
}

// After a GEMM / conv launched with fused statistics: reduce the per-(CTA, warp) partials to (mean, rstd).
static int finish_fused_stats(VCtx& c, int P, int C, int nslots) {
  int nb = nslots / 64;   // >= 64 slots (16 KB of partials) per block: one pass of eight 256 B rows per warp
  if (nb > GN_FIN_BLOCKS) nb = GN_FIN_BLOCKS;
  if (nb < 1) nb = 1;
  const int per = (nslots + nb - 1) / nb;
  IR_CUDA_CHECK(launch_pdl(gn_finalize_fused_kernel, dim3(c.B, nb), dim3(256), 0, c.s, (const float*)c.w.partial, c.w.stats,
                           c.w.fin_scratch, c.w.fin_count, nslots, per, 1.0 / ((double)P * (C / 32)), 1e-6f));
  count_launch();
  c.stats_ready = true;
  return IR_OK;
}

static bool fused_stats_ok(int C) { return C % 32 == 0 && (C / 32 == 4 || C / 32 == 8 || C / 32 == 16); }

// want_stats: the output feeds a GroupNorm next, so its statistics are accumulated in the epilogue
static int conv3x3(VCtx& c, const std::string& name, const bf16* x, bf16* y, const bf16* resid, int H, int W, int Cin,
                   int Cout, bool want_stats) {
  GemmArgs g;
  g.A = x;
  g.W = vp<bf16>(c.v, name + ".weight");
  g.ldw = 9L * Cin;
  g.M = c.B * H * W;
  g.N = Cout;
  g.K = 9 * Cin;
  g.conv = 1;
  g.nimg = c.B;
  g.H = H;
  g.Wd = W;
  g.C = Cin;
  g.epi = EPI_BF16;
  g.bias = vp<float>(c.v, name + ".bias");
  g.out_bf16 = y;
  g.resid_bf16 = resid;
  g.ldo_b = Cout;
  const int nslots = gemm_conv_tiles_per_image(H, W) * gemm_conv_gn_slots_per_tile();   // one partial per (tile, lane quarter)
  const bool fuse = want_stats && fused_stats_ok(Cout) && (long)c.B * nslots * 64 <= c.w.partial_elems;
  if (fuse) {
    g.gn_partial = c.w.partial;
    g.gn_cpg = Cout / 32;
  }
  IR_TRY(gemm_launch(g, c.s));
  if (fuse) IR_TRY(finish_fused_stats(c, H * W, Cout, nslots));
  return IR_OK;
}

// Upsample.forward (model.py:63-67): nearest x2 + 3x3 conv as four phase 2x2 convs on the low-resolution input
// (pack_upconv_phases_kernel). x: (B, H, W, C) -> y: (B, 2H, 2W, C); the output feeds a GroupNorm, whose statistics are
// accumulated in the four epilogues (one partial slot per (phase, tile)).
static int upsample_conv(VCtx& c, const std::string& name, const bf16* x, bf16* y, int H, int W, int C) {
  auto up = c.v->up_w.find(name + ".weight");
  IR_REQUIRE(up != c.v->up_w.end(), "upsample_conv: no phase weights for '%s'", name.c_str());
  const int tiles = gemm_conv_tiles_per_image(H, W);
  const int per_tile = gemm_conv_gn_slots_per_tile();
  const bool fuse = fused_stats_ok(C) && (long)c.B * 4 * tiles * per_tile * 64 <= c.w.partial_elems;
  IR_TRY(upsample_conv_phases_launch(x, up->second, vp<float>(c.v, name + ".bias"), y, c.B, H, W, C,
                                     fuse ? c.w.partial : nullptr, 0, c.s));
  if (fuse) IR_TRY(finish_fused_stats(c, 4 * H * W, C, 4 * tiles * per_tile));
  return IR_OK;
}

// 1x1 conv of an NHWC activation (nin_shortcut, attention proj_out, the gathered taps of encoder.conv_in). With a residual
// or GroupNorm statistics to fuse it runs as a single-tap implicit GEMM on the conv epilogue (pixel-owner threads,
// TMA-fetched residual, TMA-store boxes, per-tile GroupNorm partials) instead of the plain GEMM's transposing one:
// conv_in 312 -> 130 us at 1024^2, proj_out 56 -> 36 us. want_stats: the output feeds a GroupNorm.
static int conv1x1(VCtx& c, const bf16* wgt, const float* bias, const bf16* x, bf16* y, const bf16* resid, int H, int W,
                   int Cin, int Cout, bool want_stats) {
  GemmArgs g;
  g.A = x;
  g.W = wgt;
  g.ldw = Cin;
  g.M = c.B * H * W;
  g.N = Cout;
  g.K = Cin;
  if (resid || want_stats) {
    g.conv = 1;
    g.conv_taps = 1;
    g.nimg = c.B;
    g.H = H;
    g.Wd = W;
    g.C = Cin;
  } else {
    g.lda = Cin;   // nothing fused: the plain GEMM's 8-warp TMA-store epilogue measured faster (43 vs 57 us at 512^2, 128 -> 256)
  }
  g.epi = EPI_BF16;
  g.bias = bias;
  g.out_bf16 = y;
  g.resid_bf16 = resid;
  g.ldo_b = Cout;
  const int nslots = gemm_conv_tiles_per_image(H, W) * gemm_conv_gn_slots_per_tile();
  const bool fuse = want_stats && fused_stats_ok(Cout) && (long)c.B * nslots * 64 <= c.w.partial_elems;
  if (fuse) {
    g.gn_partial = c.w.partial;
    g.gn_cpg = Cout / 32;
  }
  IR_TRY(gemm_launch(g, c.s));
  if (fuse) IR_TRY(finish_fused_stats(c, H * W, Cout, nslots));
  return IR_OK;
}

// ResnetBlock.forward (model.py:131-151); h lives in buf[cur]; returns the index of the buffer holding the result
static int res_block(VCtx& c, const std::string& name, int& cur, int H, int W, int Cin, int Cout, bool out_feeds_norm) {
  const int P = H * W;
  bf16* x = c.w.buf[cur];
  bf16* t1 = c.w.buf[(cur + 1) & 3];
  bf16* t2 = c.w.buf[(cur + 2) & 3];
  bf16* t3 = c.w.buf[(cur + 3) & 3];
  IR_TRY(group_norm(c, name + ".norm1", x, t1, P, Cin, true));
  IR_TRY(conv3x3(c, name + ".conv1", t1, t2, nullptr, H, W, Cin, Cout, true));
  IR_TRY(group_norm(c, name + ".norm2", t2, t1, P, Cout, true));
  const bf16* skip = x;
  if (Cin != Cout) {
    IR_TRY(conv1x1(c, vp<bf16>(c.v, name + ".nin_shortcut.weight"), vp<float>(c.v, name + ".nin_shortcut.bias"), x, t2,
                   nullptr, H, W, Cin, Cout, false));
    skip = t2;
  }
  IR_TRY(conv3x3(c, name + ".conv2", t1, t3, skip, H, W, Cout, Cout, out_feeds_norm));
  cur = (cur + 3) & 3;
  return IR_OK;
}

// AttnBlock.forward (model.py:181-205)
static int attn_block(VCtx& c, const std::string& name, int& cur, int H, int W, int C) {
  const int P = H * W;
  bf16* x = c.w.buf[cur];
  bf16* hn = c.w.buf[(cur + 1) & 3];
  bf16* ao = c.w.buf[(cur + 2) & 3];
  bf16* y = c.w.buf[(cur + 3) & 3];
  IR_TRY(group_norm(c, name + ".norm", x, hn, P, C, false));
  {  // q, k, v in one GEMM: (B*P, C) x (3C, C)^T
    GemmArgs g;
    g.A = hn; g.lda = C; g.W = vp<bf16>(c.v, name + ".q.weight"); g.ldw = C;
    g.M = c.B * P; g.N = 3 * C; g.K = C; g.epi = EPI_BF16; g.bias = vp<float>(c.v, name + ".q.bias");
    g.out_bf16 = c.w.qkv; g.ldo_b = 3 * C;
    IR_TRY(gemm_launch(g, c.s));
  }
  const float scale = 1.0f / sqrtf((float)C);
  for (int b = 0; b < c.B; ++b) {  // one image at a time bounds the P x P score workspace
    const bf16* qkv = c.w.qkv + (long)b * P * 3 * C;
    transpose_bf16_kernel<<<dim3(div_up_l(P, 32), div_up_l(C, 32), 1), dim3(32, 8), 0, c.s>>>(qkv, c.w.vt, P, C, 3L * C,
                                                                                                2 * C);
    IR_CUDA_CHECK(cudaGetLastError());
    // softmax(q k^T C^-1/2) v in three GEMM passes (see attn_row_reduce_kernel); no P x P fp32 matrix
    const int ngroups = (P + 63) / 64;
    const float alpha2 = scale * 1.4426950408889634f;
    float* shift = c.w.att_row;
    float* inv_l = c.w.att_row + P;
    GemmArgs gs;
    gs.A = qkv; gs.lda = 3L * C; gs.W = qkv + C; gs.ldw = 3L * C;
    gs.M = P; gs.N = P; gs.K = C; gs.epi = EPI_ATTN;
    gs.att_mode = 1; gs.att_out = c.w.att_part;
    IR_TRY(gemm_launch(gs, c.s));   // pass 1: group maxima of q k^T
    IR_CUDA_CHECK(launch_pdl(attn_row_reduce_kernel, dim3(div_up_l(P, 256)), dim3(256), 0, c.s, (const float*)c.w.att_part, shift,
                             ngroups, P, alpha2, 0));
    gs.att_mode = 2; gs.alpha = alpha2; gs.att_row = shift; gs.out_bf16 = c.w.pm; gs.ldo_b = P;
    IR_TRY(gemm_launch(gs, c.s));   // pass 2: exp2(alpha2 s - shift) -> bf16, group sums
    IR_CUDA_CHECK(launch_pdl(attn_row_reduce_kernel, dim3(div_up_l(P, 256)), dim3(256), 0, c.s, (const float*)c.w.att_part, inv_l,
                             ngroups, P, 1.0f, 1));
    {  // h = (exp / sum) * v
      GemmArgs g;
      g.A = c.w.pm; g.lda = P; g.W = c.w.vt; g.ldw = P;
      g.M = P; g.N = C; g.K = P; g.epi = EPI_ATTN; g.att_mode = 3; g.att_row = inv_l;
      g.out_bf16 = ao + (long)b * P * C; g.ldo_b = C;
      IR_TRY(gemm_launch(g, c.s));
    }
    count_launch(3);
  }
  IR_TRY(conv1x1(c, vp<bf16>(c.v, name + ".proj_out.weight"), vp<float>(c.v, name + ".proj_out.bias"), ao, y, x,
                 H, W, C, C, true));   // feeds mid.block_2.norm1
  cur = (cur + 3) & 3;
  return IR_OK;
}

static int vae_decode_body(Vae* v, const float* z, float* out, int B, int h, int w, float in_scale, float out_scale,
                           float out_shift, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  IR_REQUIRE(z && out && B > 0 && h > 0 && w > 0, "vae_decode: bad arguments");
  IR_REQUIRE(h % 2 == 0 && w % 2 == 0, "vae_decode: latent size must be even");
  for (int i = 0; i < (int)v->params.size() && (v->first_encoder_param < 0 || i < v->first_encoder_param); ++i)
    IR_REQUIRE(v->params[i].loaded, "vae_decode: parameter '%s' was never loaded", v->params[i].name.c_str());
  const size_t need = vae_workspace_bytes(v, B, h, w);
  if (!workspace || workspace_bytes < need) {
    set_last_error("vae_decode: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return IR_ERR_WORKSPACE;
  }
  VCtx c;
  c.v = v;
  c.B = B;
  c.s = s;
  vae_carve(v, c.w, workspace, B, h, w);
  IR_CUDA_CHECK(cudaMemsetAsync(c.w.fin_count, 0, (size_t)B * sizeof(unsigned), s));
  const std::string d = "decoder";
  const VaeConfig& cfg = v->cfg;
  int C = cfg.ch * cfg.ch_mult[3];
  int H = h, W = w;
  int cur = 0;
  {
    const long threads = (long)B * H * W * (C / 8);
    conv_in_kernel<4><<<div_up_l(threads, 256), 256, 0, s>>>(z, vp<float>(v, "post_quant_conv.weight"),
                                                             vp<float>(v, "post_quant_conv.bias"),
                                                             vp<float>(v, d + ".conv_in.weight"),
                                                             vp<float>(v, d + ".conv_in.bias"), c.w.buf[cur], B, H, W, C,
                                                             in_scale);
    IR_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  IR_TRY(res_block(c, d + ".mid.block_1", cur, H, W, C, C, true));
  IR_TRY(attn_block(c, d + ".mid.attn_1", cur, H, W, C));
  IR_TRY(res_block(c, d + ".mid.block_2", cur, H, W, C, C, true));
  for (int lvl = 3; lvl >= 0; --lvl) {
    const int Cout = cfg.ch * cfg.ch_mult[lvl];
    for (int b = 0; b < cfg.num_res_blocks + 1; ++b) {
      // the last block of a level is followed by the (norm-free) upsample, except at level 0 (norm_out)
      const bool feeds_norm = (b < cfg.num_res_blocks) || lvl == 0;
      IR_TRY(res_block(c, d + ".up." + std::to_string(lvl) + ".block." + std::to_string(b), cur, H, W, C, Cout, feeds_norm));
      C = Cout;
    }
    static const bool legacy_upconv = [] {
      const char* e = debug_env("IR_VAE_UPCONV_LEGACY");   // debugging aid: A/B against upsample2x + 3x3 conv
      return e && e[0] == '1';
    }();
    if (lvl != 0 && !legacy_upconv) {
      IR_TRY(upsample_conv(c, d + ".up." + std::to_string(lvl) + ".upsample.conv", c.w.buf[cur], c.w.buf[(cur + 2) & 3], H, W, C));
      H *= 2;
      W *= 2;
      cur = (cur + 2) & 3;
    } else if (lvl != 0) {
      bf16* up = c.w.buf[(cur + 1) & 3];
      const long total_vec = (long)B * H * W * C / 8;   // input vectors
      int grid = div_up_l(total_vec, 256);
      if (grid > 148 * 16) grid = 148 * 16;
      upsample2x_kernel<<<grid, 256, 0, s>>>(c.w.buf[cur], up, total_vec, H, W, C);
      IR_CUDA_CHECK(cudaGetLastError());
      count_launch();
      H *= 2;
      W *= 2;
      IR_TRY(conv3x3(c, d + ".up." + std::to_string(lvl) + ".upsample.conv", up, c.w.buf[(cur + 2) & 3], nullptr, H, W,
                     C, C, true));
      cur = (cur + 2) & 3;
    }
  }
  bf16* hn = c.w.buf[(cur + 1) & 3];
  IR_TRY(group_norm(c, d + ".norm_out", c.w.buf[cur], hn, H * W, C, true));
  static const bool legacy_conv_out = [] {
    const char* e = debug_env("IR_VAE_CONVOUT_LEGACY");   // debugging aid: A/B against the CUDA-core kernel
    return e && e[0] == '1';
  }();
  if (!legacy_conv_out) {
    // conv_out (3x3, C -> 3) as a tap-response GEMM (M = pixels, N = 27 -> 32, K = C; fp32 rows into a free activation
    // buffer) + a 9-neighbour gather-sum that writes the NCHW image with the caller's affine
    IR_REQUIRE(v->cfg.out_ch == 3, "vae_decode: 3 output channels expected");
    float* Y = reinterpret_cast<float*>(c.w.buf[(cur + 2) & 3]);
    GemmArgs g;
    g.A = hn;
    g.lda = C;
    g.W = v->co_w;
    g.ldw = C;
    g.M = B * H * W;
    g.N = 32;
    g.K = C;
    g.epi = EPI_F32;
    g.out_f32 = Y;
    g.ldo_f = 32;
    IR_TRY(gemm_launch(g, s));
    IR_CUDA_CHECK(launch_pdl(conv_out_gather_kernel, dim3(div_up_l(W, 32), div_up_l(H, 8), B), dim3(256), 0, s,
                             (const float*)Y, (const float*)v->co_b, out, H, W, out_scale, out_shift));
    IR_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return IR_OK;
  }
  IR_REQUIRE(C == 128, "vae_decode: conv_out kernel is specialised for 128 input channels (got %d)", C);
  {
    constexpr int smem = (10 * 34) * (128 + 8) * 2 + 3 * 9 * 128 * 4;
    IR_TRY(ensure_smem_optin((const void*)conv_out_kernel<128>, smem));
    conv_out_kernel<128><<<dim3(div_up_l(W, 32), div_up_l(H, 8), B), 256, smem, s>>>(
        hn, vp<float>(v, d + ".conv_out.weight"), vp<float>(v, d + ".conv_out.bias"), out, H, W, out_scale, out_shift);
    IR_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return IR_OK;
}

// ================================================================================================ encode
static size_t vae_enc_carve(const Vae* v, VaeWs& w, void* base, int B, int H, int W) {
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* {
    off = (off + 255) & ~size_t(255);
    void* p = b ? b + off : nullptr;
    off += bytes;
    return p;
  };
  // largest activation: level 0 (ch channels at full resolution); deeper levels halve the pixels per channel doubling
  long cmax = 0;
  const long imax = (long)B * H * W * 32;   // conv_in's im2col operand
  {
    long Hc = H, Wc = W, C = v->cfg.ch;
    for (int lvl = 0; lvl < 4; ++lvl) {
      const long Cout = (long)v->cfg.ch * v->cfg.ch_mult[lvl];
      cmax = std::max(cmax, (long)B * Hc * Wc * std::max(C, Cout));
      C = Cout;
      if (lvl != 3) {
        Hc /= 2;
        Wc /= 2;
      }
    }
  }
  for (int i = 0; i < 4; ++i) w.buf[i] = reinterpret_cast<bf16*>(take((size_t)cmax * sizeof(bf16)));
  w.im2col = reinterpret_cast<bf16*>(take((size_t)imax * sizeof(bf16)));
  const long h = H / 8, wd = W / 8;
  const long P = h * wd, C = (long)v->cfg.ch * v->cfg.ch_mult[3];
  w.qkv = reinterpret_cast<bf16*>(take((size_t)B * P * 3 * C * sizeof(bf16)));
  w.vt = reinterpret_cast<bf16*>(take((size_t)P * C * sizeof(bf16)));
  w.pm = reinterpret_cast<bf16*>(take((size_t)P * P * sizeof(bf16)));
  w.att_part = reinterpret_cast<float*>(take((size_t)((P + 63) / 64) * P * sizeof(float)));
  w.att_row = reinterpret_cast<float*>(take((size_t)2 * P * sizeof(float)));
  w.partial_elems = (long)B * std::max<long>(GN_MAX_CHUNKS, (long)gemm_conv_tiles_per_image(H, W) * gemm_conv_gn_slots_per_tile()) * 64;
  w.partial = reinterpret_cast<float*>(take((size_t)w.partial_elems * sizeof(float)));
  w.stats = reinterpret_cast<float*>(take((size_t)B * 32 * 2 * sizeof(float)));
  w.fin_scratch = reinterpret_cast<double*>(take((size_t)B * GN_FIN_BLOCKS * 64 * sizeof(double)));
  w.fin_count = reinterpret_cast<unsigned*>(take((size_t)B * sizeof(unsigned)));
  w.f32tmp = reinterpret_cast<float*>(take((size_t)B * P * 2 * v->cfg.z_channels * sizeof(float)));
  w.g_in = reinterpret_cast<float*>(take((size_t)B * 3 * H * W * sizeof(float)));
  w.g_out = reinterpret_cast<float*>(take((size_t)B * P * 2 * v->cfg.z_channels * sizeof(float)));
  return (off + 255) & ~size_t(255);
}

size_t vae_encode_workspace_bytes(const Vae* v, int B, int H, int W) {
  VaeWs ws;
  return vae_enc_carve(v, ws, nullptr, B, H, W);
}

// Downsample.forward (model.py:92-101: F.pad(x, (0,1,0,1)) + 3x3 stride-2 conv) as an implicit GEMM: the tensor map of the
// A operand gathers the [8][16]-pixel tile of tap (ky, kx) at pixel stride 2 (elementStrides) and zero-fills the pad, so
// nothing is materialised. The output feeds the next level's first norm1: its GroupNorm statistics come out of the epilogue.
static int downsample(VCtx& c, const std::string& name, const bf16* x, bf16* y, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2;
  GemmArgs g;
  g.A = x;
  g.W = vp<bf16>(c.v, name + ".weight");
  g.ldw = 9L * C;
  g.M = c.B * Ho * Wo;
  g.N = C;
  g.K = 9 * C;
  g.conv = 1;
  g.conv_stride = 2;
  g.nimg = c.B;
  g.H = Ho;
  g.Wd = Wo;
  g.C = C;
  g.epi = EPI_BF16;
  g.bias = vp<float>(c.v, name + ".bias");
  g.out_bf16 = y;
  g.ldo_b = C;
  const int nslots = gemm_conv_tiles_per_image(Ho, Wo) * gemm_conv_gn_slots_per_tile();
  const bool fuse = fused_stats_ok(C) && (long)c.B * nslots * 64 <= c.w.partial_elems;
  if (fuse) {
    g.gn_partial = c.w.partial;
    g.gn_cpg = C / 32;
  }
  IR_TRY(gemm_launch(g, c.s));
  if (fuse) IR_TRY(finish_fused_stats(c, Ho * Wo, C, nslots));
  return IR_OK;
}

// AutoencoderKL.encode up to the moments (autoencoder.py:82-86): Encoder.forward (model.py:521-546) + quant_conv.
// x: (B,3,H,W) fp32 in [-1,1]; moments: (B, 2z, H/8, W/8) fp32 (mean | logvar).
static int vae_encode_body(Vae* v, const float* x, float* moments, int B, int H, int W, void* workspace, size_t workspace_bytes,
                           cudaStream_t s) {
  IR_REQUIRE(v->cfg.with_encoder && v->first_encoder_param >= 0, "vae_encode: handle was created without the encoder");
  IR_REQUIRE(x && moments && B > 0 && H > 0 && W > 0, "vae_encode: bad arguments");
  IR_REQUIRE(H % 16 == 0 && W % 16 == 0, "vae_encode: image size must be a multiple of 16 (got %dx%d)", H, W);
  for (int i = v->first_encoder_param; i < (int)v->params.size(); ++i)
    IR_REQUIRE(v->params[i].loaded, "vae_encode: parameter '%s' was never loaded", v->params[i].name.c_str());
  const size_t need = vae_encode_workspace_bytes(v, B, H, W);
  if (!workspace || workspace_bytes < need) {
    set_last_error("vae_encode: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return IR_ERR_WORKSPACE;
  }
  VCtx c;
  c.v = v;
  c.B = B;
  c.s = s;
  vae_enc_carve(v, c.w, workspace, B, H, W);
  IR_CUDA_CHECK(cudaMemsetAsync(c.w.fin_count, 0, (size_t)B * sizeof(unsigned), s));
  const std::string e = "encoder";
  const VaeConfig& cfg = v->cfg;
  int C = cfg.ch;
  int cur = 0;
  {
    // conv_in: the 27 taps of the 3-channel image gathered per pixel (K = 27 -> 32) + one GEMM; the first norm1's statistics
    // come out of its epilogue
    const long total_px = (long)B * H * W;
    IR_CUDA_CHECK(launch_pdl(im2col_in_kernel, dim3(div_up_l(total_px, 256)), dim3(256), 0, s, x, c.w.im2col, H, W, total_px));
    count_launch();
    // the gathered taps are an NHWC activation of 32 channels: a 1x1 conv
    IR_TRY(conv1x1(c, v->ci_w, vp<float>(v, e + ".conv_in.bias"), c.w.im2col, c.w.buf[cur], nullptr, H, W, 32, C, true));
  }
  for (int lvl = 0; lvl < 4; ++lvl) {
    const int Cout = cfg.ch * cfg.ch_mult[lvl];
    for (int b = 0; b < cfg.num_res_blocks; ++b) {
      // the last block of levels 0..2 feeds the (norm-free) downsample conv; everything else feeds a GroupNorm
      const bool feeds_norm = (b + 1 < cfg.num_res_blocks) || lvl == 3;
      IR_TRY(res_block(c, e + ".down." + std::to_string(lvl) + ".block." + std::to_string(b), cur, H, W, C, Cout, feeds_norm));
      C = Cout;
    }
    if (lvl != 3) {
      IR_TRY(downsample(c, e + ".down." + std::to_string(lvl) + ".downsample.conv", c.w.buf[cur], c.w.buf[(cur + 1) & 3], H, W, C));
      cur = (cur + 1) & 3;
      H /= 2;
      W /= 2;
    }
  }
  IR_TRY(res_block(c, e + ".mid.block_1", cur, H, W, C, C, true));
  IR_TRY(attn_block(c, e + ".mid.attn_1", cur, H, W, C));
  IR_TRY(res_block(c, e + ".mid.block_2", cur, H, W, C, C, true));
  bf16* hn = c.w.buf[(cur + 1) & 3];
  IR_TRY(group_norm(c, e + ".norm_out", c.w.buf[cur], hn, H * W, C, true));
  {
    // conv_out: 3x3, C -> 2z, fp32 output (the latents are the product of the path: no bf16 rounding here)
    GemmArgs g;
    g.A = hn;
    g.W = vp<bf16>(v, e + ".conv_out.weight");
    g.ldw = 9L * C;
    g.M = B * H * W;
    g.N = 2 * cfg.z_channels;
    g.K = 9 * C;
    g.conv = 1;
    g.nimg = B;
    g.H = H;
    g.Wd = W;
    g.C = C;
    g.epi = EPI_F32;
    g.bias = vp<float>(v, e + ".conv_out.bias");
    g.out_f32 = c.w.f32tmp;
    g.ldo_f = 2 * cfg.z_channels;
    IR_TRY(gemm_launch(g, s));
  }
  IR_REQUIRE(cfg.z_channels == 4, "vae_encode: z_channels 4 expected");
  const long P_total = (long)B * H * W;
  moments_kernel<8><<<div_up_l(P_total, 256), 256, 0, s>>>(c.w.f32tmp, vp<float>(v, "quant_conv.weight"),
                                                           vp<float>(v, "quant_conv.bias"), moments, P_total, H * W);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

// ------------------------------------------------------------------------------------------------ CUDA-graph replay
// A decode / encode call is 100-120 launches whose sizes, order and device pointers are fixed for a given (shape, workspace,
// output affine): the second call with a key captures the body into a graph (on a stream of the library's own: the
// caller's may be the legacy default stream), later calls copy the input to a fixed staging buffer of the workspace, launch
// the graph and copy the result back. Same kernels in the same order: bit-identical to the eager path (the first call with
// a key, the per-launch profile pass and calls made while the caller's stream is itself being captured stay eager).
static bool vae_stream_is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return st != cudaStreamCaptureStatusNone;
}

static void vae_destroy_graphs(Vae* v) {
  for (VaeGraph& g : v->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  v->graphs.clear();
}

void vae_set_graphs(Vae* v, bool on) {
  v->graphs_enabled = on;
  if (!on) vae_destroy_graphs(v);
}

void vae_release_graphs(Vae* v) {
  vae_destroy_graphs(v);
  if (v->cap_stream) cudaStreamDestroy(v->cap_stream);
  v->cap_stream = nullptr;
}

// body(in, out, stream) runs the eager path on the given pointers
template <typename Body>
static int vae_run_graphed(Vae* v, const VaeGraph& key, const float* in, size_t in_bytes, float* out, size_t out_bytes,
                           float* g_in, float* g_out, cudaStream_t s, Body body) {
  const bool use_graph = v->graphs_enabled && !prof_enabled() && !vae_stream_is_capturing(s);
  if (!use_graph) return body(in, out, s);
  VaeGraph* hit = nullptr;
  for (VaeGraph& g : v->graphs)
    if (g.same_key(key)) hit = &g;
  if (!hit) {
    // first call with this key: eager (one-time initialisation -- shared-memory opt-ins, the driver entry point -- stays
    // outside any capture)
    if (v->graphs.size() >= 16) vae_destroy_graphs(v);
    v->graphs.push_back(key);
    return body(in, out, s);
  }
  IR_CUDA_CHECK(cudaMemcpyAsync(g_in, in, in_bytes, cudaMemcpyDeviceToDevice, s));
  if (!hit->exec) {
    const long long before = launch_count_value();
    if (!v->cap_stream) IR_CUDA_CHECK(cudaStreamCreateWithFlags(&v->cap_stream, cudaStreamNonBlocking));
    if (cudaStreamBeginCapture(v->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      vae_set_graphs(v, false);
      return body(in, out, s);
    }
    const int st = body(g_in, g_out, v->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(v->cap_stream, &graph);
    if (st != IR_OK || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      if (st == IR_OK) set_last_error("vae: graph capture failed: %s", cudaGetErrorString(ce));
      vae_set_graphs(v, false);   // not again on this handle; the eager path serves the call
      return st != IR_OK ? st : body(in, out, s);
    }
    const int launches = (int)(launch_count_value() - before);
    count_launch(-launches);   // nothing ran during capture; replays are counted below
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      cudaGetLastError();
      vae_set_graphs(v, false);
      return body(in, out, s);
    }
    // `hit` may have been invalidated by nothing: the vector is not touched between the lookup and here
    hit->exec = exec;
    hit->launches = launches;
  }
  IR_CUDA_CHECK(cudaGraphLaunch(hit->exec, s));
  count_launch(hit->launches);
  IR_CUDA_CHECK(cudaMemcpyAsync(out, g_out, out_bytes, cudaMemcpyDeviceToDevice, s));
  return IR_OK;
}

int vae_decode(Vae* v, const float* z, float* out, int B, int h, int w, float in_scale, float out_scale,
               float out_shift, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  IR_REQUIRE(z && out && B > 0 && h > 0 && w > 0, "vae_decode: bad arguments");
  const size_t need = vae_workspace_bytes(v, B, h, w);
  if (!workspace || workspace_bytes < need) {
    set_last_error("vae_decode: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return IR_ERR_WORKSPACE;
  }
  VaeWs ws;
  vae_carve(v, ws, workspace, B, h, w);
  VaeGraph key;
  key.kind = 0; key.B = B; key.H = h; key.W = w; key.ws = workspace;
  key.f[0] = in_scale; key.f[1] = out_scale; key.f[2] = out_shift;
  return vae_run_graphed(v, key, z, (size_t)B * v->cfg.z_channels * h * w * sizeof(float), out,
                         (size_t)B * v->cfg.out_ch * 64 * h * w * sizeof(float), ws.g_in, ws.g_out, s,
                         [&](const float* in, float* o, cudaStream_t st) {
                           return vae_decode_body(v, in, o, B, h, w, in_scale, out_scale, out_shift, workspace, workspace_bytes, st);
                         });
}

int vae_encode(Vae* v, const float* x, float* moments, int B, int H, int W, void* workspace, size_t workspace_bytes,
               cudaStream_t s) {
  IR_REQUIRE(v->cfg.with_encoder && v->first_encoder_param >= 0, "vae_encode: handle was created without the encoder");
  IR_REQUIRE(x && moments && B > 0 && H > 0 && W > 0, "vae_encode: bad arguments");
  IR_REQUIRE(H % 16 == 0 && W % 16 == 0, "vae_encode: image size must be a multiple of 16 (got %dx%d)", H, W);
  const size_t need = vae_encode_workspace_bytes(v, B, H, W);
  if (!workspace || workspace_bytes < need) {
    set_last_error("vae_encode: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return IR_ERR_WORKSPACE;
  }
  VaeWs ws;
  vae_enc_carve(v, ws, workspace, B, H, W);
  VaeGraph key;
  key.kind = 1; key.B = B; key.H = H; key.W = W; key.ws = workspace;
  return vae_run_graphed(v, key, x, (size_t)B * 3 * H * W * sizeof(float), moments,
                         (size_t)B * 2 * v->cfg.z_channels * (H / 8) * (W / 8) * sizeof(float), ws.g_in, ws.g_out, s,
                         [&](const float* in, float* o, cudaStream_t st) {
                           return vae_encode_body(v, in, o, B, H, W, workspace, workspace_bytes, st);
                         });
}

}  // namespace ir
