// Micro-probe (B200): cycles of the attention kernel's exponential section per warp, with its ingredients switched on
// and off, at 1 and 2 warps per SM sub-partition. Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// /tmp/exp_probe tools/probes/exp_section_probe.cu && /tmp/exp_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// MODE bit 0: MUFU ex2; bit 1: F2FP pack; bit 2: FADD2 row sum; bit 3: FFMA2 scale
template <int MODE>
__global__ void probe(const float* in, uint32_t* out, long long* cyc, int reps) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 131 + i) & 1023];
  uint64_t sc = pack2(in[0], in[0]), nm = pack2(in[1], in[1]);
  uint32_t acc = 0;
  uint64_t sa = pack2(0.f, 0.f), sb = sa;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 128; i += 2) {
      uint64_t x = pack2(s[i], s[i + 1]);
      if (MODE & 8) x = fma2(x, sc, nm);
      float e0, e1;
      unpack2(x, e0, e1);
      if (MODE & 1) { e0 = ex2(e0); e1 = ex2(e1); }
      if (MODE & 4) { if (i & 2) sb = add2(sb, pack2(e0, e1)); else sa = add2(sa, pack2(e0, e1)); }
      if (MODE & 2) acc ^= packbf(e0, e1); else acc ^= __float_as_uint(e0) + __float_as_uint(e1);
      s[i] = e0 * 0.999f; s[i + 1] = e1 * 0.999f;   // keep the loop-carried values alive (2 FMUL per pair in every mode)
    }
  }
  long long t1 = clock64();
  float a, b2; unpack2(add2(sa, sb), a, b2);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __float_as_uint(a + b2);
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) / reps;
}

template <int MODE>
static void run(const char* name, const float* in, uint32_t* out, long long* cyc) {
  for (int warps : {4, 8, 16}) {   // 1, 2, 4 warps per SM sub-partition (one CTA per SM)
    probe<MODE><<<148, warps * 32>>>(in, out, cyc, 200);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps/SMSP %d: %6lld cycles per 128 elements per warp (%.2f cycles / element / SMSP)\n", name, warps / 4, h,
           (double)h / 128.0 / (warps / 4));
  }
}

int main() {
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
  float h[1024]; for (int i = 0; i < 1024; ++i) h[i] = -0.5f - (i % 7) * 0.1f;
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  run<0>("baseline (2 FMUL / pair)", in, out, cyc);
  run<1>("+ MUFU ex2", in, out, cyc);
  run<2>("+ F2FP pack", in, out, cyc);
  run<3>("+ MUFU + F2FP", in, out, cyc);
  run<15>("+ MUFU + F2FP + FADD2 + FFMA2", in, out, cyc);
  run<14>("+ F2FP + FADD2 + FFMA2 (no MUFU)", in, out, cyc);
  printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
