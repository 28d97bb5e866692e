// Host-side launchers of the fused elementwise / reduction kernels of the DiT + ControlNet path (elementwise.cu).
#pragma once
#include "gemm.cuh"

namespace ir {

// 2-D sin-cos position table, (gh*gw, D) fp32. PixArt.py:258-307 (float64 math, w-coordinate in the first half).
int pos_embed_launch(float* table, int gh, int gw, int D, int base_size, float pe_interpolation, cudaStream_t s);

// PatchEmbed (2x2 stride-2 conv) + bias + pos_embed. PixArtMS.py:22-46, pixart_controlnet.py:78-87,215.
// x: (B,C,H,W) fp32; wt: transposed conv weight [C*4][D] fp32; out_f32/out_bf16: (B*T, D), either may be null.
int patch_embed_launch(const float* x, const float* wt, const float* bias, const float* pos, float* out_f32,
                       bf16* out_bf16, int B, int C, int H, int W, int D, cudaStream_t s);

// Sinusoidal embedding of scalars: out[r][0:128] = cos(v*f_i), out[r][128:256] = sin(v*f_i), f_i = exp(-ln(1e4) i/128).
// PixArt_blocks.py:336-353.
int sinusoid_launch(const float* vals, float* out, int rows, cudaStream_t s);

// y[g(r)] (+)= act_out(bias + W * act_in(x[r])), fp32 weights W [N][K]; one warp per output feature.
// Output row r is stored at y + (r / rows_per_group) * group_stride + (r % rows_per_group) * ldy.
enum { ACT_NONE = 0, ACT_SILU = 1 };
int small_linear_launch(const float* x, long ldx, const float* W, const float* bias, float* y, long ldy,
                        int rows_per_group, long group_stride, int rows, int N, int K, int act_in, int act_out,
                        int accumulate, cudaStream_t s);

// mod[blk][b][j][:] = table[blk][j][:] + t0[b][j][:], j < 6. PixArtMS.py:74.
int adaln_table_launch(const float* tables, const float* t0, float* mod, int nblk, int B, int D, cudaStream_t s);

// out_bf16[row] = LN(x[row]) * (1 + scale[b]) + shift[b], eps 1e-6, no affine. PixArtMS.py:58,64,75,77.
// shift/scale: per-sample vectors, sample b = row / T at shift + b*mod_stride.
int ln_modulate_launch(const float* x, bf16* out, const float* shift, const float* scale, long mod_stride, int rows,
                       int T, int D, cudaStream_t s);

// T2IFinalLayer + unpatchify: LN, modulate with (table + t), Linear D->p*p*Cout, scatter to (B,Cout,H,W).
// PixArt_blocks.py:271-275, pixart_controlnet.py:165-177.
int final_layer_launch(const float* x, const float* table, const float* t, const float* W, const float* bias,
                       float* out, int B, int gh, int gw, int D, int cout, cudaStream_t s);

// Gather caption rows: out_bf16[i][:] = y[idx[i]][:]. pixart_controlnet.py:222-228 (masked_select packing).
int gather_rows_launch(const float* y, const int* idx, bf16* out, int rows, int K, cudaStream_t s);

// x0 = (x - sqrt(1-abar) * eps) / sqrt(abar), eps = channels [0,C) of the (B,2C,H,W) model output.
// scripts/DMD/transformer_train/generate.py:44-51,84-85.
int lincomb3_launch(const float* x, const float* m0, const float* m1, float* out, long n, float ca, float c0, float c1,
                    cudaStream_t s);
int eps_to_x0_launch(const float* x, const float* model_out, float* x0, int B, int C, int HW, float sqrt_abar,
                     float sqrt_one_minus_abar, cudaStream_t s);

int f32_to_bf16_launch(const float* x, bf16* out, long n, cudaStream_t s);

}  // namespace ir
