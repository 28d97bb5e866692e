// Stage-1 SwinIR restoration module (the `preprocess_model` of test_scripts/inference.py:92-103) on the sm_100a kernels:
// diffusion/model/swinir.py -- SwinIR.forward :867-905, forward_features :852-865, RSTB :430-493, SwinTransformerBlock
// :175-290, WindowAttention :76-156, Mlp :25-41 -- with the parameters of configs/swinir.yaml (embed 180, 8 x 6 blocks,
// 6 heads x 30, window 8, mlp_ratio 2, PixelUnshuffle(8), 'nearest+conv' upsampler, '1conv').
//
// Layout: the 180-wide token features live in rows of 192 elements (the pad columns are kept at zero), so every linear
// layer is a tcgen05 GEMM with K = 192 / 384 and every 3x3 conv is the implicit-GEMM kernel with 192 (or 64) input
// channels; the residual stream is fp32, GEMM operands bf16. Window partition, cyclic shift, relative-position bias and
// the shift mask are index arithmetic inside the window-attention kernel; nothing is materialised.
#include "swinir.cuh"

namespace ir {

namespace {

constexpr int CP = 192;   // padded token width (embed_dim 180)
constexpr int HP = 384;   // padded MLP hidden width (360)

inline int div_up_l(long a, long b) { return (int)((a + b - 1) / b); }
inline long align64(long v) { return (v + 63) / 64 * 64; }

__constant__ float c_rgb_mean[3] = {0.4488f, 0.4371f, 0.4040f};   // swinir.py:692-693

// (x - mean) * img_range, PixelUnshuffle(r) (swinir.py:871-872, 712-715): NCHW fp32 -> NHWC bf16 with C = 3*r*r,
// channel c*r*r + dy*r + dx of output pixel (y, x) = input channel c at (y*r + dy, x*r + dx).
__global__ void __launch_bounds__(256) swin_unshuffle_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int H,
                                                             int W, int r) {
  const int h = H / r, w = W / r, C = 3 * r * r;
  const long total = (long)B * h * w * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % C);
    long p = i / C;
    const int ox = (int)(p % w);
    p /= w;
    const int oy = (int)(p % h);
    const long b = p / h;
    const int c = ch / (r * r), dy = (ch / r) % r, dx = ch % r;
    const float v = x[((b * 3 + c) * H + (oy * r + dy)) * (long)W + ox * r + dx] - c_rgb_mean[c];
    out[i] = __float2bfloat16(v);
  }
}

// LayerNorm(180, eps 1e-5, affine) over rows of 192 (pad columns written as zero). One warp per row, 6 values per lane.
template <bool OUT_BF16>
__global__ void __launch_bounds__(256) swin_ln_kernel(const float* __restrict__ x, void* __restrict__ out,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta, long rows,
                                                      int C) {
  const long row = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_wait();
  pdl_launch();
  if (row >= rows) return;
  const float* xr = x + row * CP;
  float v[6];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = i * 32 + lane;
    v[i] = c < C ? xr[c] : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = i * 32 + lane;
    const float d = c < C ? v[i] - mean : 0.f;
    v[i] = d;
    sq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(sq) / C + 1e-5f);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int c = i * 32 + lane;
    const float y = c < C ? v[i] * rstd * gamma[c] + beta[c] : 0.f;
    if (OUT_BF16)
      reinterpret_cast<bf16*>(out)[row * CP + c] = __float2bfloat16(y);
    else
      reinterpret_cast<float*>(out)[row * CP + c] = y;
  }
}

// ---- warp-level tensor-core helpers (mma.sync m16n8k16, ldmatrix) for the 64-token windows
IR_DEVINL void sw_ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
IR_DEVINL void sw_ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
IR_DEVINL void sw_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Window attention (WindowAttention.forward, swinir.py:125-156) for window 8 and head_dim 30, fused with window_partition /
// window_reverse (:44-73), the cyclic shift (:261-265, 282-286), the relative-position bias and the shift mask (:227-248).
// grid = (windows * B, heads), block = 4 warps: warp w owns query tokens [16w, 16w+16) of the window; S = q k^T and P v run
// on mma.sync (head_dim padded to 32 with zeros), softmax in registers (a row lives in one lane quad).
__global__ void __launch_bounds__(128) swin_window_attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                               const float* __restrict__ bias /* [heads][64][64] */, int H,
                                                               int W, int C, int heads, int shift, float scale) {
  constexpr int WS = 8, N = 64, HDIM = 30, LDS_ = 40;   // smem row stride in bf16 (80 B: conflict-free ldmatrix)
  __shared__ __align__(16) bf16 sq[N * LDS_], sk[N * LDS_], sv[N * LDS_];
  __shared__ int sreg[N];
  __shared__ long stok[N];
  const int head = blockIdx.y;
  const int nwx = W / WS, nw = (H / WS) * nwx;
  const int b = blockIdx.x / nw, wi = blockIdx.x % nw;
  const int wy = wi / nwx, wx = wi % nwx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_wait();
  pdl_launch();
  if (tid < N) {
    const int sy = wy * WS + tid / WS, sx = wx * WS + tid % WS;   // coordinates in the shifted image
    const int y = (sy + shift) % H, x = (sx + shift) % W;         // torch.roll(x, -shift): shifted[sy] = x[(sy + shift) % H]
    stok[tid] = ((long)b * H + y) * W + x;
    int reg = 0;
    if (shift > 0) {   // calculate_mask: region id of the token in the shifted image
      const int ry = sy < H - WS ? 0 : (sy < H - shift ? 1 : 2);
      const int rx = sx < W - WS ? 0 : (sx < W - shift ? 1 : 2);
      reg = ry * 3 + rx;
    }
    sreg[tid] = reg;
  }
  __syncthreads();
  // q, k, v rows of this head (30 bf16 = 15 words each) -> shared memory rows of 20 words, words 15..19 zero
  for (int i = tid; i < N * 3 * 20; i += 128) {
    const int wd = i % 20, which = (i / 20) % 3, r = i / 60;
    uint32_t v = 0u;
    if (wd < HDIM / 2) v = *reinterpret_cast<const uint32_t*>(qkv + stok[r] * (3L * C) + which * C + head * HDIM + wd * 2);
    bf16* dst = which == 0 ? sq : (which == 1 ? sk : sv);
    *reinterpret_cast<uint32_t*>(dst + r * LDS_ + wd * 2) = v;
  }
  __syncthreads();
  // ---- S = Q K^T (16 query rows of this warp x 64 keys)
  uint32_t qf[2][4];
  {
    const int r = warp * 16 + (lane & 15), cbase = (lane >> 4) * 8;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) sw_ldsm_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], smem_u32(sq + r * LDS_ + kk * 16 + cbase));
  }
  float sc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sc[i][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      const int kr = np * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int kc = kk * 16 + ((lane >> 3) & 1) * 8;
      uint32_t b0, b1, b2, b3;
      sw_ldsm_x4(b0, b1, b2, b3, smem_u32(sk + kr * LDS_ + kc));
      sw_mma(sc[2 * np], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
      sw_mma(sc[2 * np + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
    }
  }
  // ---- scale, relative-position bias, shift mask, softmax (rows r0 = lane/4 and r0 + 8 of the warp's 16)
  const int r0 = warp * 16 + (lane >> 2);
  const float* bias0 = bias + ((long)head * N + r0) * N;
  const float* bias1 = bias0 + 8 * N;
  const int reg0 = sreg[r0], reg1 = sreg[r0 + 8];
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = i * 8 + (lane & 3) * 2;
    const float2 bb0 = *reinterpret_cast<const float2*>(bias0 + col), bb1 = *reinterpret_cast<const float2*>(bias1 + col);
    const int rc0 = sreg[col], rc1 = sreg[col + 1];
    sc[i][0] = sc[i][0] * scale + bb0.x + (rc0 != reg0 ? -100.0f : 0.f);
    sc[i][1] = sc[i][1] * scale + bb0.y + (rc1 != reg0 ? -100.0f : 0.f);
    sc[i][2] = sc[i][2] * scale + bb1.x + (rc0 != reg1 ? -100.0f : 0.f);
    sc[i][3] = sc[i][3] * scale + bb1.y + (rc1 != reg1 ? -100.0f : 0.f);
    mx0 = fmaxf(mx0, fmaxf(sc[i][0], sc[i][1]));
    mx1 = fmaxf(mx1, fmaxf(sc[i][2], sc[i][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i][0] = __expf(sc[i][0] - mx0);
    sc[i][1] = __expf(sc[i][1] - mx0);
    sc[i][2] = __expf(sc[i][2] - mx1);
    sc[i][3] = __expf(sc[i][3] - mx1);
    sum0 += sc[i][0] + sc[i][1];
    sum1 += sc[i][2] + sc[i][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  // ---- O = P V (P stays in registers as the A fragments)
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t a0 = pack_bf16x2(sc[2 * kk][0], sc[2 * kk][1]);
    const uint32_t a1 = pack_bf16x2(sc[2 * kk][2], sc[2 * kk][3]);
    const uint32_t a2 = pack_bf16x2(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
    const uint32_t a3 = pack_bf16x2(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
    const int vr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      const int vc = np * 16 + (lane >> 4) * 8;
      uint32_t b0, b1, b2, b3;
      sw_ldsm_x4_t(b0, b1, b2, b3, smem_u32(sv + vr * LDS_ + vc));
      sw_mma(o[2 * np], a0, a1, a2, a3, b0, b1);
      sw_mma(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- window_reverse + roll(+shift): every token goes back to its own position; d 30, 31 are padding
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const long tok = stok[r0 + h2 * 8];
    const float inv = h2 == 0 ? inv0 : inv1;
    bf16* orow = out + tok * CP + head * HDIM + (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i * 8 + (lane & 3) * 2 < HDIM)
        *reinterpret_cast<uint32_t*>(orow + i * 8) = pack_bf16x2(o[i][2 * h2] * inv, o[i][2 * h2 + 1] * inv);
    }
  }
  if (head == 0) {   // keep the pad columns of the 192-wide rows at zero
    for (int i = tid; i < N * (CP - 180); i += 128) out[stok[i / (CP - 180)] * CP + 180 + i % (CP - 180)] = __float2bfloat16(0.f);
  }
}

// exact GELU (nn.GELU default, erf) in place on rows of HP elements; the pad columns are (re)zeroed
__global__ void __launch_bounds__(256) swin_gelu_kernel(bf16* __restrict__ h, long rows, int hidden) {
  const long total = rows * HP;
  pdl_wait();
  pdl_launch();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % HP);
    float v = 0.f;
    if (c < hidden) {
      const float x = __bfloat162float(h[i]);
      v = 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    }
    h[i] = __float2bfloat16(v);
  }
}

// LeakyReLU in place on a bf16 tensor
__global__ void __launch_bounds__(256) swin_lrelu_kernel(bf16* __restrict__ x, long n8, float slope) {
  pdl_wait();
  pdl_launch();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
    uint4 u = *reinterpret_cast<const uint4*>(x + i * 8);
    uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = unpack_bf16x2(w[k]);
      w[k] = pack_bf16x2(a.x >= 0.f ? a.x : a.x * slope, a.y >= 0.f ? a.y : a.y * slope);
    }
    *reinterpret_cast<uint4*>(x + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// F.interpolate(scale_factor=2, mode='nearest') on NHWC bf16, 16-byte vectors
__global__ void __launch_bounds__(256) swin_upsample2x_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long total_vec,
                                                              int H, int W, int C) {
  const int tpp = C / 8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total_vec; i += (long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % tpp);
    long pix = i / tpp;
    const int ox = (int)(pix % (2 * W));
    pix /= 2 * W;
    const int oy = (int)(pix % (2 * H));
    const long n = pix / (2 * H);
    const uint4 u = *reinterpret_cast<const uint4*>(x + (((n * H + oy / 2) * W + ox / 2) * (long)C) + cv * 8);
    *reinterpret_cast<uint4*>(y + i * 8) = u;
  }
}

// conv_last output (NHWC fp32, 4 columns of which 3 are real) -> x / img_range + mean, NCHW fp32 (swinir.py:899-901)
__global__ void __launch_bounds__(256) swin_output_kernel(const float* __restrict__ h, float* __restrict__ out, long P_total,
                                                          long P) {
  const long pix = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (pix >= P_total) return;
  const float4 v = *reinterpret_cast<const float4*>(h + pix * 4);
  const long b = pix / P, p = pix % P;
  out[(b * 3 + 0) * P + p] = v.x + c_rgb_mean[0];
  out[(b * 3 + 1) * P + p] = v.y + c_rgb_mean[1];
  out[(b * 3 + 2) * P + p] = v.z + c_rgb_mean[2];
}

// ---- weight packing
// Linear (cout, cin) fp32 -> bf16 [cout][ld] with zero pad columns
__global__ void swin_pack_linear_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int cin, int ld) {
  const long total = (long)cout * ld;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % ld);
    const long o = i / ld;
    dst[i] = __float2bfloat16(c < cin ? src[o * cin + c] : 0.f);
  }
}
// Conv (cout, cin, 3, 3) fp32 -> bf16 [cout][tap][cpad] (tap-major K of the implicit GEMM), zero pad channels
__global__ void swin_pack_conv_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int cout, int cin, int cpad) {
  const long total = (long)cout * 9 * cpad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cpad);
    const int tap = (int)((i / cpad) % 9);
    const long o = i / (9L * cpad);
    dst[i] = __float2bfloat16(c < cin ? src[(o * cin + c) * 9 + tap] : 0.f);
  }
}
// relative_position_bias_table (225, heads) -> bias[head][64][64] through relative_position_index (swinir.py:103-114,138-141)
__global__ void swin_pack_relbias_kernel(const float* __restrict__ table, float* __restrict__ dst, int heads) {
  const int total = heads * 64 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i % 64, q = (i / 64) % 64, h = i / 4096;
    const int dy = q / 8 - j / 8 + 7, dx = q % 8 - j % 8 + 7;
    dst[i] = table[(dy * 15 + dx) * heads + h];
  }
}

int cpad_of(int c) { return (c + 63) / 64 * 64; }

void swin_add(Swin* s, const std::string& name, int kind, long numel, int cout, int cin, int k) {
  SwinParam p;
  p.name = name;
  p.kind = kind;
  p.numel = numel;
  p.cout = cout;
  p.cin = cin;
  p.k = k;
  if (kind == SP_LINEAR) {
    p.offset = s->wb_elems;
    s->wb_elems = align64(s->wb_elems + (long)cout * cpad_of(cin));
  } else if (kind == SP_CONV) {
    // rows padded to a multiple of 4 (GEMM N % 4 == 0; conv_last has 3 output channels): the extra row stays zero
    p.offset = s->wb_elems;
    s->wb_elems = align64(s->wb_elems + (long)((cout + 3) / 4 * 4) * 9 * cpad_of(cin));
  } else if (kind == SP_RELBIAS) {
    p.offset = s->wf_elems;
    s->wf_elems = align64(s->wf_elems + (long)s->cfg.heads * 64 * 64);
  } else {
    p.offset = s->wf_elems;
    s->wf_elems = align64(s->wf_elems + numel + 4);   // slack stays zero: a 3-entry bias is read as 4
  }
  s->index[name] = (int)s->params.size();
  s->params.push_back(p);
}
void swin_add_linear(Swin* s, const std::string& n, int cout, int cin) {
  swin_add(s, n + ".weight", SP_LINEAR, (long)cout * cin, cout, cin, 1);
  swin_add(s, n + ".bias", SP_F32, cout, cout, 1, 1);
}
void swin_add_conv(Swin* s, const std::string& n, int cout, int cin) {
  swin_add(s, n + ".weight", SP_CONV, (long)cout * cin * 9, cout, cin, 3);
  swin_add(s, n + ".bias", SP_F32, cout, cout, 1, 1);
}
void swin_add_norm(Swin* s, const std::string& n, int c) {
  swin_add(s, n + ".weight", SP_F32, c, c, 1, 1);
  swin_add(s, n + ".bias", SP_F32, c, c, 1, 1);
}

template <typename T>
const T* sp(const Swin* s, const std::string& name) {
  auto it = s->index.find(name);
  if (it == s->index.end()) return nullptr;
  const SwinParam& p = s->params[it->second];
  if (p.kind == SP_LINEAR || p.kind == SP_CONV) return reinterpret_cast<const T*>(s->wb + p.offset);
  return reinterpret_cast<const T*>(s->wf + p.offset);
}

struct SwinWs {
  float *x, *first, *rin, *f32out;
  bf16 *img, *xn, *qkv, *att, *hid, *xb, *u0, *u1;
};

size_t swin_carve(const Swin* s, SwinWs& w, void* base, int B, int H, int W) {
  uint8_t* b = reinterpret_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) -> void* {
    off = (off + 255) & ~size_t(255);
    void* p = b ? b + off : nullptr;
    off += bytes;
    return p;
  };
  const int r = s->cfg.sf;
  const long L = (long)B * (H / r) * (W / r);
  const long P = (long)B * H * W;
  w.x = (float*)take(L * CP * 4);
  w.first = (float*)take(L * CP * 4);
  w.rin = (float*)take(L * CP * 4);
  w.img = (bf16*)take(L * 3 * r * r * 2);
  w.xn = (bf16*)take(L * CP * 2);
  w.qkv = (bf16*)take(L * 3 * s->cfg.embed_dim * 2);
  w.att = (bf16*)take(L * CP * 2);
  w.hid = (bf16*)take(L * HP * 2);
  w.xb = (bf16*)take(L * CP * 2);
  w.u0 = (bf16*)take(P * s->cfg.num_feat * 2);
  w.u1 = (bf16*)take(P * s->cfg.num_feat * 2);
  w.f32out = (float*)take(P * 4 * 4);
  return (off + 255) & ~size_t(255);
}

struct SCtx {
  Swin* s;
  SwinWs w;
  cudaStream_t st;
};

int linear(SCtx& c, const std::string& name, const bf16* A, long lda, int M, int N, int K, int epi, bf16* out_b, long ldo_b,
           float* out_f, const float* resid_f) {
  GemmArgs g;
  g.A = A;
  g.lda = lda;
  g.W = sp<bf16>(c.s, name + ".weight");
  g.ldw = K;
  g.M = M;
  g.N = N;
  g.K = K;
  g.epi = epi;
  g.gelu_erf = 1;   // SwinIR's Mlp uses nn.GELU (erf)
  g.bias = sp<float>(c.s, name + ".bias");
  g.out_bf16 = out_b;
  g.ldo_b = ldo_b;
  g.out_f32 = out_f;
  g.resid_f32 = resid_f;
  g.ldo_f = CP;
  return gemm_launch(g, c.st);
}

// 3x3 conv, pad 1, NHWC bf16 input with Cin (multiple of 64) channels; EPI_F32 (fp32 out, optional fp32 residual and bf16
// copy) when out_f is set, else EPI_BF16
int conv3(SCtx& c, const std::string& name, const bf16* x, int B, int H, int W, int Cin, int N, bf16* out_b, long ldo_b,
          float* out_f, long ldo_f, const float* resid_f) {
  GemmArgs g;
  g.A = x;
  g.W = sp<bf16>(c.s, name + ".weight");
  g.ldw = 9L * Cin;
  g.M = B * H * W;
  g.N = N;
  g.K = 9 * Cin;
  g.conv = 1;
  g.nimg = B;
  g.H = H;
  g.Wd = W;
  g.C = Cin;
  g.bias = sp<float>(c.s, name + ".bias");
  g.out_bf16 = out_b;
  g.ldo_b = ldo_b;
  if (out_f) {
    g.epi = EPI_F32;
    g.out_f32 = out_f;
    g.resid_f32 = resid_f;
    g.ldo_f = ldo_f;
  } else {
    g.epi = EPI_BF16;
  }
  return gemm_launch(g, c.st);
}

int lrelu(SCtx& c, bf16* x, long n, float slope) {
  IR_REQUIRE(n % 8 == 0, "lrelu: element count must be a multiple of 8");
  int grid = div_up_l(n / 8, 256);
  if (grid > 148 * 16) grid = 148 * 16;
  IR_CUDA_CHECK(launch_pdl(swin_lrelu_kernel, dim3(grid), dim3(256), 0, c.st, x, n / 8, slope));
  count_launch();
  return IR_OK;
}

}  // namespace

int swin_create(const SwinConfig& cfg, Swin** out) {
  IR_REQUIRE(cfg.embed_dim == 180 && cfg.heads == 6 && cfg.window == 8 && cfg.mlp_ratio == 2 && cfg.sf == 8 && cfg.num_feat == 64,
             "swinir: kernels are specialised for configs/swinir.yaml (embed 180, 6 heads, window 8, mlp_ratio 2, sf 8)");
  Swin* s = new Swin();
  s->cfg = cfg;
  const int C = cfg.embed_dim, hid = C * cfg.mlp_ratio;
  swin_add_conv(s, "conv_first.1", C, 3 * cfg.sf * cfg.sf);
  swin_add_norm(s, "patch_embed.norm", C);
  for (int l = 0; l < cfg.num_layers; ++l) {
    for (int b = 0; b < cfg.depth; ++b) {
      const std::string p = "layers." + std::to_string(l) + ".residual_group.blocks." + std::to_string(b);
      swin_add_norm(s, p + ".norm1", C);
      swin_add(s, p + ".attn.relative_position_bias_table", SP_RELBIAS, 225L * cfg.heads, cfg.heads, 225, 1);
      swin_add_linear(s, p + ".attn.qkv", 3 * C, C);
      swin_add_linear(s, p + ".attn.proj", C, C);
      swin_add_norm(s, p + ".norm2", C);
      swin_add_linear(s, p + ".mlp.fc1", hid, C);
      swin_add_linear(s, p + ".mlp.fc2", C, hid);
    }
    swin_add_conv(s, "layers." + std::to_string(l) + ".conv", C, C);
  }
  swin_add_norm(s, "norm", C);
  swin_add_conv(s, "conv_after_body", C, C);
  swin_add_conv(s, "conv_before_upsample.0", cfg.num_feat, C);
  swin_add_conv(s, "conv_up1", cfg.num_feat, cfg.num_feat);
  swin_add_conv(s, "conv_up2", cfg.num_feat, cfg.num_feat);
  swin_add_conv(s, "conv_up3", cfg.num_feat, cfg.num_feat);
  swin_add_conv(s, "conv_hr", cfg.num_feat, cfg.num_feat);
  swin_add_conv(s, "conv_last", 3, cfg.num_feat);
  if (cudaMalloc(&s->wb, (size_t)s->wb_elems * sizeof(bf16)) != cudaSuccess ||
      cudaMalloc(&s->wf, (size_t)s->wf_elems * sizeof(float)) != cudaSuccess) {
    set_last_error("swin_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    swin_destroy(s);
    return IR_ERR_CUDA;
  }
  cudaMemset(s->wb, 0, (size_t)s->wb_elems * sizeof(bf16));
  cudaMemset(s->wf, 0, (size_t)s->wf_elems * sizeof(float));
  *out = s;
  return IR_OK;
}

void swin_destroy(Swin* s) {
  if (!s) return;
  cudaFree(s->wb);
  cudaFree(s->wf);
  delete s;
}

int swin_load_param(Swin* s, const char* name, const float* src, long numel, cudaStream_t st) {
  auto it = s->index.find(name);
  if (it == s->index.end()) {
    set_last_error("swin_load_param: unknown parameter '%s'", name);
    return IR_ERR_INVALID;
  }
  SwinParam& p = s->params[it->second];
  IR_REQUIRE(numel == p.numel, "swin_load_param: '%s' has %ld elements, expected %ld", name, numel, p.numel);
  switch (p.kind) {
    case SP_LINEAR: {
      const int ld = cpad_of(p.cin);
      swin_pack_linear_kernel<<<div_up_l((long)p.cout * ld, 256), 256, 0, st>>>(src, s->wb + p.offset, p.cout, p.cin, ld);
      break;
    }
    case SP_CONV: {
      const int cp = cpad_of(p.cin);
      swin_pack_conv_kernel<<<div_up_l((long)p.cout * 9 * cp, 256), 256, 0, st>>>(src, s->wb + p.offset, p.cout, p.cin, cp);
      break;
    }
    case SP_RELBIAS:
      swin_pack_relbias_kernel<<<div_up_l(s->cfg.heads * 4096, 256), 256, 0, st>>>(src, s->wf + p.offset, s->cfg.heads);
      break;
    default:
      IR_CUDA_CHECK(cudaMemcpyAsync(s->wf + p.offset, src, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  IR_CUDA_CHECK(cudaGetLastError());
  p.loaded = true;
  return IR_OK;
}

size_t swin_workspace_bytes(const Swin* s, int B, int H, int W) {
  SwinWs w;
  return swin_carve(s, w, nullptr, B, H, W);
}

int swin_forward(Swin* s, const float* x, float* out, int B, int H, int W, void* workspace, size_t workspace_bytes,
                 cudaStream_t st) {
  const SwinConfig& cfg = s->cfg;
  IR_REQUIRE(x && out && B > 0, "swin_forward: bad arguments");
  IR_REQUIRE(H % (cfg.sf * cfg.window) == 0 && W % (cfg.sf * cfg.window) == 0 && H > 0 && W > 0,
             "swin_forward: image size %dx%d must be a multiple of %d (unshuffle %d x window %d)", H, W, cfg.sf * cfg.window,
             cfg.sf, cfg.window);
  for (const SwinParam& p : s->params) IR_REQUIRE(p.loaded, "swin_forward: parameter '%s' was never loaded", p.name.c_str());
  const size_t need = swin_workspace_bytes(s, B, H, W);
  if (!workspace || workspace_bytes < need) {
    set_last_error("swin_forward: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return IR_ERR_WORKSPACE;
  }
  SCtx c;
  c.s = s;
  c.st = st;
  swin_carve(s, c.w, workspace, B, H, W);
  const int C = cfg.embed_dim, hid = C * cfg.mlp_ratio, r = cfg.sf, F = cfg.num_feat;
  const int h = H / r, w = W / r;
  const long L = (long)B * h * w;
  const int Cin0 = 3 * r * r;   // 192
  auto grid_for = [](long n) {
    int g = div_up_l(n, 256);
    return g > 148 * 16 ? 148 * 16 : g;
  };
  // fp32 streams and the bf16 conv operand: pad columns must be zero (the kernels below write the real columns only)
  IR_CUDA_CHECK(cudaMemsetAsync(c.w.x, 0, L * CP * 4, st));
  IR_CUDA_CHECK(cudaMemsetAsync(c.w.first, 0, L * CP * 4, st));
  IR_CUDA_CHECK(cudaMemsetAsync(c.w.xb, 0, L * CP * 2, st));
  IR_CUDA_CHECK(cudaMemsetAsync(c.w.hid, 0, L * HP * 2, st));

  swin_unshuffle_kernel<<<grid_for(L * Cin0), 256, 0, st>>>(x, c.w.img, B, H, W, r);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  // conv_first (swinir.py:883) -> x_first (fp32, kept for the long skip)
  IR_TRY(conv3(c, "conv_first.1", c.w.img, B, h, w, Cin0, C, nullptr, 0, c.w.first, CP, nullptr));
  // patch_embed (flatten + LayerNorm, :535-539) -> residual stream x
  IR_CUDA_CHECK(launch_pdl(swin_ln_kernel<false>, dim3(div_up_l(L, 8)), dim3(256), 0, st, (const float*)c.w.first, (void*)c.w.x,
                           sp<float>(s, "patch_embed.norm.weight"), sp<float>(s, "patch_embed.norm.bias"), L, C));
  count_launch();
  const float scale = 1.0f / sqrtf((float)(C / cfg.heads));
  const int nwin = (h / cfg.window) * (w / cfg.window);
  for (int l = 0; l < cfg.num_layers; ++l) {
    IR_CUDA_CHECK(cudaMemcpyAsync(c.w.rin, c.w.x, L * CP * 4, cudaMemcpyDeviceToDevice, st));   // RSTB skip (:493)
    for (int bi = 0; bi < cfg.depth; ++bi) {
      const std::string p = "layers." + std::to_string(l) + ".residual_group.blocks." + std::to_string(bi);
      const int shift = (bi % 2 == 0) ? 0 : cfg.window / 2;
      IR_CUDA_CHECK(launch_pdl(swin_ln_kernel<true>, dim3(div_up_l(L, 8)), dim3(256), 0, st, (const float*)c.w.x, (void*)c.w.xn,
                               sp<float>(s, p + ".norm1.weight"), sp<float>(s, p + ".norm1.bias"), L, C));
      IR_TRY(linear(c, p + ".attn.qkv", c.w.xn, CP, (int)L, 3 * C, CP, EPI_BF16, c.w.qkv, 3L * C, nullptr, nullptr));
      IR_CUDA_CHECK(launch_pdl(swin_window_attn_kernel, dim3(nwin * B, cfg.heads), dim3(128), 0, st, (const bf16*)c.w.qkv, c.w.att,
                               sp<float>(s, p + ".attn.relative_position_bias_table"), h, w, C, cfg.heads, shift, scale));
      IR_TRY(linear(c, p + ".attn.proj", c.w.att, CP, (int)L, C, CP, EPI_F32, nullptr, 0, c.w.x, c.w.x));   // x += proj(attn)
      IR_CUDA_CHECK(launch_pdl(swin_ln_kernel<true>, dim3(div_up_l(L, 8)), dim3(256), 0, st, (const float*)c.w.x, (void*)c.w.xn,
                               sp<float>(s, p + ".norm2.weight"), sp<float>(s, p + ".norm2.bias"), L, C));
      // fc1 + exact GELU in the GEMM epilogue; the pad columns [360, 384) of hid stay at their memset zero
      IR_TRY(linear(c, p + ".mlp.fc1", c.w.xn, CP, (int)L, hid, CP, EPI_BF16_GELU, c.w.hid, HP, nullptr, nullptr));
      // x += fc2(gelu(fc1)); the last block of the group also leaves the bf16 copy the RSTB conv reads
      const bool last = bi + 1 == cfg.depth;
      IR_TRY(linear(c, p + ".mlp.fc2", c.w.hid, HP, (int)L, C, HP, EPI_F32, last ? c.w.xb : nullptr, CP, c.w.x, c.w.x));
      count_launch(4);
    }
    // x = conv(residual_group(x)) + x_in (:492-493): tokens are already the NHWC image of 192-channel pixels
    IR_TRY(conv3(c, "layers." + std::to_string(l) + ".conv", c.w.xb, B, h, w, CP, C, nullptr, 0, c.w.x, CP, c.w.rin));
  }
  // norm (:864) -> conv_after_body + x_first (:884)
  swin_ln_kernel<true><<<div_up_l(L, 8), 256, 0, st>>>(c.w.x, c.w.xn, sp<float>(s, "norm.weight"), sp<float>(s, "norm.bias"), L, C);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  IR_TRY(conv3(c, "conv_after_body", c.w.xn, B, h, w, CP, C, c.w.xb, CP, c.w.x, CP, c.w.first));
  // upsampler 'nearest+conv' with upscale 8 (:885-892)
  IR_TRY(conv3(c, "conv_before_upsample.0", c.w.xb, B, h, w, CP, F, c.w.u0, F, nullptr, 0, nullptr));
  IR_TRY(lrelu(c, c.w.u0, L * F, 0.01f));   // nn.LeakyReLU(inplace=True): default slope
  bf16 *cur = c.w.u0, *nxt = c.w.u1;
  int ch = h, cw = w;
  const char* ups[3] = {"conv_up1", "conv_up2", "conv_up3"};
  for (int u = 0; u < 3; ++u) {
    const long total_vec = (long)B * (2 * ch) * (2 * cw) * F / 8;
    swin_upsample2x_kernel<<<grid_for(total_vec), 256, 0, st>>>(cur, nxt, total_vec, ch, cw, F);
    IR_CUDA_CHECK(cudaGetLastError());
    count_launch();
    ch *= 2;
    cw *= 2;
    IR_TRY(conv3(c, ups[u], nxt, B, ch, cw, F, F, cur, F, nullptr, 0, nullptr));
    IR_TRY(lrelu(c, cur, (long)B * ch * cw * F, 0.2f));
  }
  IR_TRY(conv3(c, "conv_hr", cur, B, ch, cw, F, F, nxt, F, nullptr, 0, nullptr));
  IR_TRY(lrelu(c, nxt, (long)B * ch * cw * F, 0.2f));
  // conv_last (3 channels padded to 4) in fp32, then + mean and NCHW
  IR_TRY(conv3(c, "conv_last", nxt, B, ch, cw, F, 4, nullptr, 0, c.w.f32out, 4, nullptr));
  const long P_total = (long)B * H * W;
  swin_output_kernel<<<div_up_l(P_total, 256), 256, 0, st>>>(c.w.f32out, out, P_total, (long)H * W);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

}  // namespace ir
