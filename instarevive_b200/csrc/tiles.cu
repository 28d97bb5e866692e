// Tiled-latent scheduler kernels and pixel post-processing (integer-indexed, HBM-bound):
//   tile gather / ordered overlap blend  -- the two tile loops of process(), test_scripts/inference.py:119-153
//   wavelet colour fix / AdaIN            -- utils/image/align_color.py:44-119
//   clamp * 255 -> uint8 NHWC             -- test_scripts/inference.py:159-160
// Bit-exactness: window coordinates are integers computed on the host exactly as _sliding_windows does
// (inference.py:40-53); the blend adds the covering tiles of each pixel in tile-list order, which is the order of the
// reference's `buffer[tile] += out` loop, then divides by the (integer-valued) cover count once -- the same fp32
// operation sequence as the reference, independent of how tiles were sharded across GPUs.
#include "tiles.cuh"

namespace ir {

static inline int div_up_t(long a, long b) { return (int)((a + b - 1) / b); }
static inline int capped_grid(long total, int threads) {
  long g = (total + threads - 1) / threads;
  return (int)(g > 148L * 32 ? 148L * 32 : (g < 1 ? 1 : g));
}

// dst[t][n][c][y][x] = src[n][c][hi_t*scale + y][wi_t*scale + x]; coords: int32 [ntiles][2] = (hi, wi) in tile units
__global__ void tile_gather_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                   const int* __restrict__ coords, int ntiles, int N, int C, int H, int W, int th,
                                   int tw, int scale) {
  const long per_tile = (long)N * C * th * tw;
  const long total = per_tile * ntiles;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / per_tile);
    long r = i - (long)t * per_tile;
    const int x = (int)(r % tw);
    r /= tw;
    const int y = (int)(r % th);
    r /= th;  // r = n*C + c
    const int hi = coords[2 * t] * scale, wi = coords[2 * t + 1] * scale;
    dst[i] = src[(r * H + hi + y) * W + wi + x];
  }
}

// out[n][c][y][x] = (sum over tiles t (in index order) covering (y, x) of tiles[t][n][c][y-hi][x-wi]) / cover count
__global__ void tile_blend_kernel(const float* __restrict__ tiles, const int* __restrict__ coords, int ntiles,
                                  float* __restrict__ out, int N, int C, int H, int W, int th, int tw, int scale) {
  extern __shared__ int sc[];  // [ntiles][2]
  for (int i = threadIdx.x; i < 2 * ntiles; i += blockDim.x) sc[i] = coords[i] * scale;
  __syncthreads();
  const long per_tile = (long)N * C * th * tw;
  const long total = (long)N * C * H * W;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long nc = i / ((long)W * H);
    float acc = 0.f;
    float cnt = 0.f;
    for (int t = 0; t < ntiles; ++t) {
      const int ly = y - sc[2 * t], lx = x - sc[2 * t + 1];
      if (ly >= 0 && ly < th && lx >= 0 && lx < tw) {
        acc += tiles[(long)t * per_tile + (nc * th + ly) * tw + lx];
        cnt += 1.0f;
      }
    }
    out[i] = acc / cnt;
  }
}

int tile_gather_launch(const float* src, float* dst, const int* coords, int ntiles, int N, int C, int H, int W, int th,
                       int tw, int scale, cudaStream_t s) {
  IR_REQUIRE(src && dst && coords && ntiles > 0 && th <= H && tw <= W, "tile_gather: bad arguments");
  const long total = (long)ntiles * N * C * th * tw;
  tile_gather_kernel<<<capped_grid(total, 256), 256, 0, s>>>(src, dst, coords, ntiles, N, C, H, W, th, tw, scale);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

int tile_blend_launch(const float* tiles, const int* coords, int ntiles, float* out, int N, int C, int H, int W, int th,
                      int tw, int scale, cudaStream_t s) {
  IR_REQUIRE(tiles && out && coords && ntiles > 0 && ntiles <= 4096, "tile_blend: bad arguments");
  const long total = (long)N * C * H * W;
  tile_blend_kernel<<<capped_grid(total, 256), 256, 2 * ntiles * sizeof(int), s>>>(tiles, coords, ntiles, out, N, C, H,
                                                                                    W, th, tw, scale);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

// ------------------------------------------------------------------------------------------------ wavelet colour fix
// One a-trous level for a stack of images (img index = blockIdx.y): low = blur(cur, radius) with replicate padding and
// the separable-looking but explicitly 3x3 kernel [1 2 1]^2/16 (wavelet_blur, align_color.py:73-92);
// high += cur - low (wavelet_decomposition, align_color.py:94-106).
__global__ void wavelet_level_kernel(const float* __restrict__ cur, float* __restrict__ low, float* __restrict__ high,
                                     int H, int W, int radius, int planes_with_high) {
  const int plane = blockIdx.y;
  const float* src = cur + (long)plane * H * W;
  float* dl = low + (long)plane * H * W;
  const float kw[3] = {0.0625f, 0.125f, 0.0625f};
  const float km[3] = {0.125f, 0.25f, 0.125f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
    const int x = i % W, y = i / W;
    float acc = 0.f;
    // same accumulation order as a direct 3x3 correlation: rows top to bottom, columns left to right
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = min(max(y + dy * radius, 0), H - 1);
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = min(max(x + dx * radius, 0), W - 1);
        const float wgt = (dy == 0) ? km[dx + 1] : kw[dx + 1];
        acc += wgt * src[yy * W + xx];
      }
    }
    dl[i] = acc;
    if (plane < planes_with_high) {
      float* dh = high + (long)plane * H * W;
      dh[i] += src[i] - acc;
    }
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long n) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

size_t wavelet_workspace_bytes(int N, int C, int H, int W) {
  // two ping-pong stacks of (content, style) planes + the content high-frequency accumulator
  return (size_t)(2 * 2 + 1) * N * C * H * W * sizeof(float) + 1024;
}

int wavelet_reconstruction_launch(const float* content, const float* style, float* out, int N, int C, int H, int W,
                                  void* workspace, size_t workspace_bytes, cudaStream_t s) {
  IR_REQUIRE(content && style && out, "wavelet: null pointer");
  const long plane_elems = (long)N * C * H * W;
  if (!workspace || workspace_bytes < wavelet_workspace_bytes(N, C, H, W)) {
    set_last_error("wavelet: workspace too small");
    return IR_ERR_WORKSPACE;
  }
  float* ping = reinterpret_cast<float*>(workspace);  // [2][N*C*H*W]: content then style
  float* pong = ping + 2 * plane_elems;
  float* high = pong + 2 * plane_elems;
  IR_CUDA_CHECK(cudaMemcpyAsync(ping, content, plane_elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
  IR_CUDA_CHECK(cudaMemcpyAsync(ping + plane_elems, style, plane_elems * sizeof(float), cudaMemcpyDeviceToDevice, s));
  IR_CUDA_CHECK(cudaMemsetAsync(high, 0, plane_elems * sizeof(float), s));
  const int planes = 2 * N * C;
  float* cur = ping;
  float* nxt = pong;
  for (int lvl = 0; lvl < 5; ++lvl) {
    wavelet_level_kernel<<<dim3(div_up_t((long)H * W, 256), planes), 256, 0, s>>>(cur, nxt, high, H, W, 1 << lvl, N * C);
    IR_CUDA_CHECK(cudaGetLastError());
    float* t = cur;
    cur = nxt;
    nxt = t;
  }
  // content_high_freq + style_low_freq (align_color.py:119)
  add_kernel<<<capped_grid(plane_elems, 256), 256, 0, s>>>(high, cur + plane_elems, out, plane_elems);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch(6);
  return IR_OK;
}

// ------------------------------------------------------------------------------------------------ AdaIN
// adaptive_instance_normalization, align_color.py:44-71: per (n, c) mean and unbiased variance (+1e-5).
__global__ void __launch_bounds__(256) adain_kernel(const float* __restrict__ content, const float* __restrict__ style,
                                                    float* __restrict__ out, int HW) {
  __shared__ double red[4][8];
  const long base = (long)blockIdx.x * HW;
  double cs = 0, cq = 0, ss = 0, sq = 0;
  for (int i = threadIdx.x; i < HW; i += 256) {
    const double a = content[base + i], b = style[base + i];
    cs += a;
    cq += a * a;
    ss += b;
    sq += b * b;
  }
  double v[4] = {cs, cq, ss, sq};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double t = 0;
    for (int w = 0; w < 8; ++w) t += red[k][w];
    v[k] = t;
  }
  const double n = HW;
  const double cmean = v[0] / n, smean = v[2] / n;
  const double cvar = (v[1] - n * cmean * cmean) / (n - 1.0) + 1e-5;
  const double svar = (v[3] - n * smean * smean) / (n - 1.0) + 1e-5;
  const float cm = (float)cmean, sm_ = (float)smean, cstd = (float)sqrt(cvar), sstd = (float)sqrt(svar);
  for (int i = threadIdx.x; i < HW; i += 256) out[base + i] = (content[base + i] - cm) / cstd * sstd + sm_;
}

int adain_launch(const float* content, const float* style, float* out, int N, int C, int HW, cudaStream_t s) {
  IR_REQUIRE(content && style && out && HW > 1, "adain: bad arguments");
  adain_kernel<<<N * C, 256, 0, s>>>(content, style, out, HW);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

// ------------------------------------------------------------------------------------------------ uint8
// (N,3,H,W) fp32 -> (N,H,W,3) uint8: clamp(0,1) * 255, truncation (numpy astype(np.uint8) after clip).
__global__ void to_uint8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int N, int C, int H, int W) {
  const long total = (long)N * H * W * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long r = i / C;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const long n = r / H;
    float v = img[((n * C + c) * H + y) * W + x];
    v = fminf(fmaxf(v, 0.f), 1.f) * 255.0f;
    out[i] = (uint8_t)fminf(fmaxf(v, 0.f), 255.f);
  }
}

int to_uint8_launch(const float* img, uint8_t* out, int N, int C, int H, int W, cudaStream_t s) {
  const long total = (long)N * C * H * W;
  to_uint8_kernel<<<capped_grid(total, 256), 256, 0, s>>>(img, out, N, C, H, W);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

}  // namespace ir
