// Host-side interface of the tcgen05 GEMM / implicit-GEMM convolution kernel (gemm.cu).
#pragma once
#include "common.cuh"

namespace ir {

enum GemmEpilogue {
  EPI_BF16 = 0,       // out_bf16 = alpha*acc + bias (+ resid_bf16)
  EPI_BF16_GELU = 1,  // out_bf16 = gelu(alpha*acc + bias): tanh approximation, or exact erf with GemmArgs::gelu_erf
  EPI_F32 = 2,        // out_f32 = resid_f32 + gate * (alpha*acc + bias); optional bf16 copy of out_f32
  EPI_QKV = 3,        // attn.qkv projection scattered head-major for the tcgen05 attention kernel:
                      //   q, k -> [B][H][T][hd] bf16, v -> transposed [B][H][hd][Tp] bf16 (acc + bias)
  EPI_ATTN = 4,       // the passes of a materialised single-head attention (VAE AttnBlock, model.py:181-205), see att_mode
  EPI_BF16_GELU_ERF = 5,  // internal: the instantiation EPI_BF16_GELU + GemmArgs::gelu_erf is launched as (callers pass EPI_BF16_GELU)
};

// C[b][m][n] = sum_k A[b][m][k] * W[b][n][k]   (both operands K-major bf16, fp32 accumulation in TMEM).
// conv != 0: A is an NHWC activation [nimg][H][Wd][C] and the kernel computes a 3x3, stride 1, pad 1
// convolution as an implicit GEMM with K = 9*C (k = tap*C + c, tap = ky*3 + kx) and M = nimg*H*Wd;
// W is [N][9*C].
struct GemmArgs {
  const bf16* A = nullptr;
  long lda = 0;      // row stride of A in elements (plain GEMM)
  long strideA = 0;  // batch stride of A in elements (0 with batch > 1: every batch entry reads the same A)
  const bf16* W = nullptr;
  long ldw = 0;
  long strideW = 0;
  int M = 0, N = 0, K = 0, batch = 1;

  int conv = 0;
  int nimg = 0, H = 0, Wd = 0, C = 0;

  int epi = EPI_BF16;
  int gelu_erf = 0;   // EPI_BF16_GELU only
  float alpha = 1.0f;
  const float* bias = nullptr;  // [N]
  long stride_bias = 0;         // batch stride of bias in elements

  bf16* out_bf16 = nullptr;
  const bf16* resid_bf16 = nullptr;  // EPI_BF16 only; same layout as out_bf16
  long ldo_b = 0;
  long stride_ob = 0;

  float* out_f32 = nullptr;
  const float* resid_f32 = nullptr;  // same layout as out_f32; may alias out_f32
  long ldo_f = 0;
  long stride_of = 0;

  const float* gate = nullptr;  // gate[(row / rows_per_gate) * gate_ld + n]
  long gate_ld = 0;
  int rows_per_gate = 1;

  // Optional fused GroupNorm statistics (EPI_BF16 only): per (image, group) sum / sum of squares of the STORED bf16
  // output, written as per-(tile, warp) partials gn_partial[img][tile][32][2] (the 4 epilogue warps are combined in fixed order; every entry is written exactly
  // once: no zeroing; tile = 128-row tile index inside the image). Requires N == 32 * gn_cpg; plain GEMM:
  // gn_rows_per_img % 128 == 0.
  float* gn_partial = nullptr;
  int gn_cpg = 0;
  int gn_rows_per_img = 0;

  // EPI_ATTN only. att_mode 1: no matrix output, att_out[pair * M + row] = max of acc over the 64-column pair `pair` of the
  // row. att_mode 2: out_bf16 = exp2(alpha * acc - att_row[row]) and att_out[pair * M + row] = the fp32 sum of those
  // exponentials over the pair. att_mode 3: out_bf16 = acc * att_row[row]. att_out holds ceil(N / 64) * M floats.
  int att_mode = 0;
  const float* att_row = nullptr;
  float* att_out = nullptr;

  // EPI_QKV only: N = 3*H*hd, rows are (b, t) with T tokens per sample
  bf16* q_heads = nullptr;
  bf16* k_heads = nullptr;
  bf16* vt_heads = nullptr;
  int qkv_T = 0, qkv_Tp = 0, qkv_H = 0, qkv_hd = 0;

  // conv only. conv_taps: taps per axis (3 = 3x3, pad 1; 2 = one phase of "nearest x2 upsample + 3x3 conv" folded into a
  // 2x2 conv on the low-resolution input, K = 4*C). conv_off_*: halo origin (source pixel = output pixel + off + tap).
  // Output pixel (y, x) is stored at (y * o_scale + o_oy, x * o_scale + o_ox) of an (H * o_scale) x (Wd * o_scale) image.
  int conv_taps = 3;
  // conv_taps 1: a 1x1 conv on the NHWC activation (K = C; C < 64 allowed when C % 8 == 0) -- same arithmetic as a plain
  // GEMM, but on the conv epilogue (pixel-owner threads, TMA-store boxes, per-tile GroupNorm partials).
  // conv_stride 2 (3x3 only): Downsample's stride-2 conv with zero padding on the right / bottom (model.py:92-101).
  // (H, Wd) name the OUTPUT grid, the input activation is (2H, 2Wd); source pixel = (2y + ky, 2x + kx).
  int conv_stride = 1;
  int conv_off_y = -1, conv_off_x = -1;
  int o_scale = 1, o_oy = 0, o_ox = 0;
  // fused GroupNorm statistics of a conv: partial slot = img * gn_slots_img + gn_slot_off + tile (0 = tiles per image)
  int gn_slot_off = 0, gn_slots_img = 0;

  int sm_limit = 0;  // 0 = all SMs; otherwise the persistent grid uses at most this many SMs (two half-width GEMMs of independent
                     // streams then run side by side instead of one after the other)
  int force_bn = 0;  // 0 = heuristic; 64 / 128 / 256 = single-CTA tile width; cg*1000 + bn forces (cta_group, width)
};

// bf16 tiled tensor map (innermost dimension first; strides in bytes for dims 1..rank-1; zero OOB fill).
// swizzle_bytes: 128 or 32 (the inner box extent must equal it), or 0 = no swizzle (inner box extent a multiple of 16 B).
int make_tensor_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes, int elem_bytes = 2,   // elem_bytes: 2 = bf16, 4 = fp32
                    const uint32_t* elem_strides = nullptr);   // traversal stride per dimension (nullptr = 1): the box then
                                                               // names the traversed extent, ceil(box / stride) elements land
int device_num_sms();
int gemm_conv_tiles_per_image(int H, int W);  // 128-pixel tiles per image of the implicit-GEMM conv
// GroupNorm partial slots a conv (bf16 output) writes per tile: one per TMEM lane quarter (no barrier between the epilogue
// warps); partial[(((img * slots_img + slot_off + tile) * 4 + quarter) * 32 + group) * 2]
int gemm_conv_gn_slots_per_tile();

// Launch with the programmatic-dependent-launch attribute (the kernel must call pdl_wait()); IR_NO_PDL=1 disables it.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream);
// diagnostics (library built with -DIR_DEBUG only): CTA 0 of every GEMM launch writes %globaltimer stamps of its roles
// into this device buffer of slots x 16 int64, launch i into record i % slots (tools/gpu_gemm_trace.py); nullptr = off
void gemm_set_trace(long long* device_buf, int slots);

// number of kernel launches issued by this library since load (bench.py's gpu_launches counter)
void count_launch(int n = 1);
long long launch_count_value();

// Optional per-launch timing (bench.py roofline pass): CUDA events around the launches of one kernel class.
enum ProfClass { PROF_GEMM = 0, PROF_CONV = 1, PROF_ATTN = 2, PROF_XATTN = 3, PROF_CAL = 4, PROF_NUM = 8 };   // ATTN: tcgen05 self-attention; XATTN: var-len cross-attention; CAL: empty kernels of ir_profile_calibrate
bool prof_enabled();
void prof_before(cudaStream_t s);
void prof_after(cudaStream_t s, int klass, double flops, int M = 0, int N = 0, int K = 0);

}  // namespace ir
