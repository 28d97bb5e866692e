// Micro-probe (B200): a tiled TMA load through a tensor map with elementStrides = 2 on the two pixel dimensions of an
// NHWC bf16 tensor -- the A operand of a stride-2 3x3 convolution (Downsample.forward, model.py:92-101: pad (0,1,0,1),
// stride 2) as an implicit GEMM. Questions: (1) how many bytes does the load complete on the mbarrier (the dense
// [8][16][64] box = 16384, or the traversed 16 x 32 pixel extent)? (2) is the shared-memory image dense, row r = (i, j)
// holding pixel (y0 + 2 i, x0 + 2 j), 128B-swizzled like an unstrided box? (3) are out-of-bounds pixels zero-filled?
// The wait is bounded, so a wrong byte count cannot hang the GPU.
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I instarevive_b200/csrc -o /tmp/tma_probe
//              tools/probes/tma_elem_stride_probe.cu && /tmp/tma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
typedef __nv_bfloat16 bf16;
namespace ir { void set_last_error(const char*, ...) {} }
#include "common.cuh"
using namespace ir;

constexpr int H = 20, W = 40, C = 64;
constexpr int BOX_BYTES = 8 * 16 * 64 * 2;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, uint16_t* out, int* status, int x0, int y0,
                                             uint32_t expect) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 2 * BOX_BYTES);
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * BOX_BYTES / 2; i += 128) reinterpret_cast<uint16_t*>(sm)[i] = 0xFFFFu;   // sentinel
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, expect);
    tma_load_3d(sm, &tm, bar, 0, x0, y0);
  }
  bool done = false;
  const long long t0 = clock64();
  while (!done && clock64() - t0 < 4000000LL) done = mbar_try_wait(bar, 0);   // ~2 ms bound
  if (tid == 0) status[0] = done ? 1 : 0;
  __syncthreads();
  for (int i = tid; i < 2 * BOX_BYTES / 2; i += 128) out[i] = reinterpret_cast<uint16_t*>(sm)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  // element value encodes (y, x, c): y * 1000 + x * 10 + (c & 7) is exact in bf16 only for small numbers, so encode
  // y * 64 + x in the integer range bf16 holds exactly (< 256) per channel parity instead: channels 0..31 hold y, 32..63 x
  std::vector<bf16> h((size_t)H * W * C);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      for (int c = 0; c < C; ++c) h[((size_t)y * W + x) * C + c] = __float2bfloat16(c < 32 ? (float)(y + 1) : (float)(x + 1));
  bf16* d;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  uint16_t* dout;
  int* dstat;
  cudaMalloc(&dout, 2 * BOX_BYTES);
  cudaMalloc(&dstat, 4);
  const int smem = 1024 + 2 * BOX_BYTES + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  // variants: box given as the traversed extent (32 x 16 elements at stride 2 -> 16 x 8 pixels) or as the loaded count (16 x 8)
  for (int variant = 0; variant < 2; ++variant) {
    cuuint64_t gd[3] = {C, W, H};
    cuuint64_t gs[2] = {C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t bx[3] = {64, variant == 0 ? 32u : 16u, variant == 0 ? 16u : 8u};
    cuuint32_t es[3] = {1, 2, 2};
    CUtensorMap tm;
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d (box %u x %u, elementStrides 2,2): encode CUresult %d\n", variant, bx[1], bx[2], (int)r);
    fflush(stdout);
    if (r != CUDA_SUCCESS) continue;
    const int starts[3][2] = {{0, 0}, {1, 2}, {W - 30, H - 13}};   // the last one runs past the right / bottom edge
    for (auto& st : starts) {
      for (uint32_t expect : {(uint32_t)BOX_BYTES, (uint32_t)(variant == 0 ? BOX_BYTES : BOX_BYTES / 4), (uint32_t)(4 * BOX_BYTES)}) {
        cudaMemset(dstat, 0, 4);
        probe<<<1, 128, smem>>>(tm, dout, dstat, st[0], st[1], expect);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("  CUDA error %s\n", cudaGetErrorString(e));
          return 1;
        }
        int stat;
        cudaMemcpy(&stat, dstat, 4, cudaMemcpyDeviceToHost);
        std::vector<uint16_t> o(BOX_BYTES);
        cudaMemcpy(o.data(), dout, 2 * BOX_BYTES, cudaMemcpyDeviceToHost);
        // decode rows: row r (128 B) chunk c16 is stored at chunk (c16 ^ (r & 7)) under SWIZZLE_128B
        int written_rows = 0, ok_rows = 0;
        for (int r = 0; r < 256; ++r) {
          auto at = [&](int c) { return o[r * 64 + (((c >> 3) ^ (r & 7)) << 3) + (c & 7)]; };
          if (at(0) == 0xFFFFu && at(40) == 0xFFFFu) continue;
          ++written_rows;
          const int i = r / 16, j = r % 16;
          const int y = st[1] + 2 * i, x = st[0] + 2 * j;
          bf16 by, bxv;
          uint16_t uy = at(3), ux = at(40);
          memcpy(&by, &uy, 2);
          memcpy(&bxv, &ux, 2);
          const float gy = __bfloat162float(by), gx = __bfloat162float(bxv);
          const bool inb = y < H && x < W;
          if ((inb && gy == (float)(y + 1) && gx == (float)(x + 1)) || (!inb && gy == 0.f && gx == 0.f)) ++ok_rows;
        }
        printf("  start (x %2d, y %2d) expect_tx %6u: barrier %s, rows written %3d, rows matching (y0+2i, x0+2j | zero OOB) %3d\n",
               st[0], st[1], expect, stat ? "COMPLETED" : "not completed", written_rows, ok_rows);
        fflush(stdout);
      }
    }
  }
  printf("TMA STRIDE PROBE DONE\n");
  return 0;
}
