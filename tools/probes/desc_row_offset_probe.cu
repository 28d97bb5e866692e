// Micro-probe (B200): may the A-operand shared-memory descriptor of tcgen05.mma (K-major, SWIZZLE_128B) start at a row
// that is NOT a multiple of 8 rows (1024 B) inside a tile written with the TMA 128B-swizzle pattern? The implicit-GEMM
// conv reuses one halo box for the three ky taps through +16-row (+2048 B, swizzle-phase preserving) descriptor offsets;
// if +1-row (+128 B) offsets work too, the three kx taps can share one (18 x 10) box and the conv's L2 -> shared-memory
// A traffic drops 3x. Variants per row offset: descriptor base-offset field [49,52) = 0, or
// (start_address >> 7) & 7 (the rule the PTX ISA gives for a pattern that does not start on a 1024 B boundary).
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I instarevive_b200/csrc -o /tmp/desc_probe
//              tools/probes/desc_row_offset_probe.cu && /tmp/desc_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
typedef __nv_bfloat16 bf16;
namespace ir { void set_last_error(const char*, ...) {} }
#include "common.cuh"
using namespace ir;

constexpr int ROWS_EXT = 160;   // rows of the extended A tile in shared memory (128 + room for offsets)
constexpr int N = 32;           // UMMA N
constexpr int K = 64;           // one 128 B swizzle row

// a_ext: [ROWS_EXT][64] bf16 row-major; b: [N][64] bf16 row-major; out: [128][N] fp32
__global__ void __launch_bounds__(128) probe(const bf16* a_ext, const bf16* b, float* out, int row_off, int use_base_off) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = sm;                              // ROWS_EXT * 128 B, 1024-aligned: the pattern TMA SWIZZLE_128B writes
  uint8_t* sB = sm + ROWS_EXT * 128;             // N * 128 B (ROWS_EXT * 128 is a multiple of 1024)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < ROWS_EXT * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a_ext + r * 64 + c * 8);
  }
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * 64 + c * 8);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 32);
    tmem_relinquish();
  }
  fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)row_off * 128u;
    uint64_t da = make_smem_desc_sw128(a_addr);
    if (use_base_off) da |= (uint64_t)((a_addr >> 7) & 7u) << 49;
    const uint64_t db = make_smem_desc_sw128(smem_u32(sB));
    constexpr uint32_t idesc = make_idesc_bf16(128, N);
#pragma unroll
    for (int k = 0; k < K / 16; ++k) umma_bf16(tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < N; ++j) out[(warp * 32 + (tid & 31)) * N + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<bf16> ha(ROWS_EXT * 64), hb(N * 64);
  std::vector<float> fa(ROWS_EXT * 64), fb(N * 64);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { ha[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { hb[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f); fb[i] = __bfloat162float(hb[i]); }
  bf16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 1024 + ROWS_EXT * 128 + N * 128 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(128 * N);
  const int offs[] = {0, 8, 16, 1, 2, 3, 5, 9, 17, 18};
  for (int off : offs) {
    for (int ub = 0; ub < 2; ++ub) {
      cudaMemset(dout, 0, 128 * N * 4);
      probe<<<1, 128, smem>>>(da, db, dout, off, ub);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("row_off %2d base_off %d: CUDA error %s\n", off, ub, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(ho.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      int bad_rows = 0;
      for (int r = 0; r < 128; ++r) {
        double rowerr = 0;
        for (int j = 0; j < N; ++j) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)fa[(r + off) * 64 + k] * fb[j * 64 + k];
          rowerr = fmax(rowerr, fabs(ref - ho[r * N + j]));
        }
        if (rowerr > 1e-3) ++bad_rows;
        maxerr = fmax(maxerr, rowerr);
      }
      printf("row_off %2d base_offset_field %s: max err %.4g, wrong rows %d/128 -> %s\n", off, ub ? "(addr>>7)&7" : "0          ",
             maxerr, bad_rows, bad_rows == 0 ? "OK" : "MISMATCH");
    }
  }
  printf("DESC PROBE DONE\n");
  return 0;
}
