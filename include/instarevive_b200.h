/*
 * instarevive_b200 -- C ABI of the B200-native (sm_100a) one-step restoration hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes only (no torch / C++ types). Every device pointer is
 * owned by the caller (PyTorch); the library owns only its packed weight copies and small per-handle caches.
 * Calls are asynchronous on the CUDA stream passed as `stream` (a cudaStream_t cast to void*); they never
 * synchronise the device. All functions return 0 on success and a non-zero ir_status otherwise, in which case
 * ir_last_error() describes the failure (the Python host raises RuntimeError, mirroring the reference's
 * exception-only error convention: assert / raise in diffusion/model/nets/pixart_controlnet.py:172,
 * scripts/DMD/transformer_train/generate.py:17).
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   ir_dit_*          ControlPixArtMSHalf.forward           diffusion/model/nets/pixart_controlnet.py:191-251
 *                     (PixArtMSBlock.forward                diffusion/model/nets/PixArtMS.py:71-79,
 *                      AttentionKVCompress / MultiHeadCrossAttention / T2IFinalLayer / embedders
 *                                                            diffusion/model/nets/PixArt_blocks.py:28-58,123-158,259-463)
 *   ir_eps_to_x0      eps_to_mu + chunk(2)[0]               scripts/DMD/transformer_train/generate.py:44-51,84-85
 *   ir_vae_*          AutoencoderKL.decode -> Decoder.forward ldm/models/autoencoder.py:88-91,
 *                                                            ldm/modules/diffusionmodules/model.py:622-655
 *   ir_tile_*         tile loops of process()               test_scripts/inference.py:119-153
 *   ir_wavelet_*      wavelet_reconstruction                utils/image/align_color.py:73-119
 *   ir_to_uint8       clamp / *255 / uint8 NHWC             test_scripts/inference.py:159-160
 */
#ifndef INSTAREVIVE_B200_H_
#define INSTAREVIVE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ir_status {
  IR_STATUS_OK = 0,
  IR_STATUS_INVALID = 1,     /* bad argument / shape */
  IR_STATUS_CUDA = 2,        /* a CUDA runtime call failed */
  IR_STATUS_UNSUPPORTED = 3, /* configuration outside what the kernels are specialised for */
  IR_STATUS_WORKSPACE = 4,   /* caller-provided workspace too small */
  IR_STATUS_DRIVER = 5       /* driver entry point (tensor-map encode) unavailable */
} ir_status;

/* Thread-local description of the last failure on this thread; never NULL. */
const char* ir_last_error(void);
/* Library version string. */
const char* ir_version(void);
/* Kernels launched by this library since load (all threads); bench.py reports the delta as gpu_launches. */
long long ir_launch_count(void);

/* Per-launch timing of the tensor-core kernels (bench.py roofline pass). Between begin and end every GEMM / conv /
 * attention launch is bracketed by CUDA events on its stream; end synchronises the device and returns, per class
 * (0 = tcgen05 GEMM, 1 = tcgen05 implicit-GEMM conv, 2 = tcgen05 self-attention, 3 = var-len cross-attention; arrays of
 * 8), the summed launch durations in ms,
 * the summed algorithmic FLOPs and the launch counts. Not thread-safe; not for use inside CUDA-graph capture. */
void ir_profile_begin(void);
int ir_profile_end(double* ms_by_class, double* flops_by_class, long long* launches_by_class);
/* Per-launch records of the running profile pass, in launch order: class, problem shape (GEMM: M x N x K with the batch
 * folded into M; conv: pixels x Cout x taps*Cin; self-attention: B*heads, T, head_dim) and duration in ms. Call after the
 * work was enqueued and before ir_profile_end; synchronises the device. Returns the record count (only `cap` written). */
long long ir_profile_records(int* klass, int* M, int* N, int* K, float* ms, long long cap);
/* Floor of the event-pair measurement: enqueues n empty kernels, each bracketed like a profiled launch (class 4 of
 * ir_profile_end: ms / launches = what an event pair adds to a launch's duration on this device). Between begin and end. */
int ir_profile_calibrate(int n, void* stream);

/* ------------------------------------------------------------------ DiT + ControlNet-Half ---- */
typedef struct ir_dit ir_dit; /* opaque: packed weights of one (device, model) */

typedef struct ir_dit_config {
  int depth;              /* 28 for PixArtMS_XL_2 */
  int copy_blocks;        /* 13: ControlPixArtMSHalf(copy_blocks_num) */
  int hidden;             /* 1152 */
  int heads;              /* 16 */
  int patch;              /* 2 */
  int in_channels;        /* 4 */
  int out_channels;       /* 8 (learned sigma) */
  int caption_channels;   /* 4096 */
  int mlp_ratio;          /* 4 */
  int base_size;          /* input_size // patch_size (PixArt.py:100) */
  float pe_interpolation; /* PixArt.py:76 */
} ir_dit_config;

int ir_dit_create(const ir_dit_config* cfg, ir_dit** out);
void ir_dit_destroy(ir_dit* h);
/* Parameter table: the names are the reference state_dict keys (SURVEY 8b weight contract). */
int ir_dit_num_params(const ir_dit* h);
int ir_dit_param_info(const ir_dit* h, int i, char* name, int name_cap, long long* numel, int* rows, int* cols);
/* Copy one fp32 parameter (device pointer, reference layout) into the packed device representation. */
int ir_dit_load_param(ir_dit* h, const char* name, const float* src_dev, long long numel, void* stream);
size_t ir_dit_workspace_bytes(const ir_dit* h, int B, int H, int W, int sum_l);
/* Pre-size the caches the handle owns -- the 2-D sincos position table (PixArt.py:258-307; up to max_tokens tokens) and the
 * caption K/V of all blocks (PixArt_blocks.py:47-50; up to max_sum_l packed caption tokens) -- so that no forward calls
 * cudaMalloc. Optional: forwards grow the buffers on demand (grow-only, never freed before ir_dit_destroy). */
int ir_dit_reserve(ir_dit* h, int max_tokens, int max_sum_l);
/* CUDA-graph replay of ir_dit_forward (default on): the second call with a given (B, H, W, caption layout, workspace)
 * captures the forward, later calls replay it (inputs / output are staged through fixed buffers of the workspace).
 * enable = 0 switches back to plain stream-ordered launches and drops the cached graphs. */
int ir_dit_set_graphs(ir_dit* h, int enable);
/* Dual-chain schedule (default on): the ControlNet chain (pixart_controlnet.py:238-240: controlnet[i] depends on the base
 * chain only through block 0) runs on a second stream ahead of the base chain, one event per control block. enable = 0
 * launches everything on the caller's stream in program order (profilers that join launches by order want that). */
int ir_dit_set_dual_chain(ir_dit* h, int enable);
/*
 * x, c: (B,4,H,W) fp32 latents (c may be NULL: plain 28-block path); timestep: (B) fp32;
 * y: (rows,4096) fp32 caption embeddings; y_index: device int32 (sum_l) valid rows of y, sample-major;
 * kv_off / kv_len: device int32 (B) offsets / lengths into the packed caption rows; max_l = max over samples of
 * (kv_off % 8 + kv_len) (<= 384; the cross-attention key window, whose TMA box starts at the packed row rounded down to a
 * multiple of 8), kv_total = sum over samples of kv_len (FLOP accounting) -- host copies;
 * img_hw: device fp32 (B,2); aspect: device fp32 (B); out: (B,8,H,W) fp32.
 * reuse_caption != 0 skips the caption embedding and the per-block K/V projections and reuses the ones cached by
 * the previous call on this handle (same captions, same sum_l).
 */
int ir_dit_forward(ir_dit* h, const float* x, const float* c, const float* timestep, const float* y,
                   const int32_t* y_index, const int32_t* kv_off, const int32_t* kv_len, const float* img_hw,
                   const float* aspect, float* out, int B, int H, int W, int sum_l, int max_l, long long kv_total,
                   int reuse_caption, void* workspace, size_t workspace_bytes, void* stream);

/* x0 = (x - sqrt(1-abar) * eps) / sqrt(abar); eps = channels [0,C) of model_out (B,2C,H,W). */
int ir_eps_to_x0(const float* x, const float* model_out, float* x0, int B, int C, int HW, float sqrt_abar,
                 float sqrt_one_minus_abar, void* stream);

/* out = ca*x + c0*m0 + c1*m1 (m1 may be NULL), fp32, in place allowed: the state update of the multistep DPM-Solver++
 * (diffusion/model/dpm_solver.py:551-597 first-order, :805-863 second-order). */
int ir_lincomb3(const float* x, const float* m0, const float* m1, float* out, long long n, float ca, float c0, float c1,
                void* stream);

/* ------------------------------------------------------------------ VAE decoder ---- */
typedef struct ir_vae ir_vae; /* opaque: packed decoder weights */

typedef struct ir_vae_config {
  int ch;             /* 128 (configs/cldm.yaml:69-84) */
  int z_channels;     /* 4 */
  int out_ch;         /* 3 */
  int num_res_blocks; /* 2 */
  int ch_mult[4];     /* 1,2,4,4 */
  int with_encoder;   /* nonzero: the handle also holds encoder.* and quant_conv.* and can run ir_vae_encode */
} ir_vae_config;

int ir_vae_create(const ir_vae_config* cfg, ir_vae** out);
void ir_vae_destroy(ir_vae* h);
/* Parameter names are the reference keys: post_quant_conv.{weight,bias}, decoder.* (SURVEY 8b "VAE surface"). */
int ir_vae_num_params(const ir_vae* h);
int ir_vae_param_info(const ir_vae* h, int i, char* name, int name_cap, long long* numel);
int ir_vae_load_param(ir_vae* h, const char* name, const float* src_dev, long long numel, void* stream);
size_t ir_vae_workspace_bytes(const ir_vae* h, int B, int h_lat, int w_lat);
/* CUDA-graph replay of ir_vae_decode / ir_vae_encode (default on): the second call with a given (batch, size, workspace,
 * output affine) captures the ~110 launches of the call, later calls replay them (input / output staged through fixed
 * buffers of the workspace). enable = 0: plain stream-ordered launches, cached graphs dropped. */
int ir_vae_set_graphs(ir_vae* h, int enable);
/* out (B,3,8h,8w) fp32 = decode(z * in_scale) * out_scale + out_shift; z: (B,4,h,w) fp32 latents.
 * in_scale = 1/scaling_factor and out = x/2 + 0.5 reproduce test_scripts/inference.py:116-117,140-142. */
int ir_vae_decode(ir_vae* h, const float* z, float* out, int B, int h_lat, int w_lat, float in_scale, float out_scale,
                  float out_shift, void* workspace, size_t workspace_bytes, void* stream);

/* VAE encoder: moments (B, 8, H/8, W/8) fp32 = quant_conv(Encoder(x)), x: (B,3,H,W) fp32 in [-1,1]
 * (AutoencoderKL.encode, ldm/models/autoencoder.py:82-86; Encoder.forward, ldm/modules/diffusionmodules/model.py:521-546;
 * Downsample, model.py:70-89). Channels [0,4) are the mean = DiagonalGaussianDistribution.mode()
 * (ldm/modules/distributions/distributions.py:24-62), channels [4,8) the log-variance. Parameter names:
 * encoder.*, quant_conv.{weight,bias}. H and W must be multiples of 16. */
size_t ir_vae_encode_workspace_bytes(const ir_vae* h, int B, int H, int W);
int ir_vae_encode(ir_vae* h, const float* x, float* moments, int B, int H, int W, void* workspace, size_t workspace_bytes,
                  void* stream);

/* ------------------------------------------------------------------ SwinIR stage 1 ---- */
/* The `preprocess_model` of test_scripts/inference.py:92-103,245-248: diffusion/model/swinir.py SwinIR.forward (:867-905)
 * with the parameters of configs/swinir.yaml. Parameter names are the reference's (conv_first.1.*, patch_embed.norm.*,
 * layers.L.residual_group.blocks.B.{norm1,attn.{relative_position_bias_table,qkv,proj},norm2,mlp.{fc1,fc2}}.*,
 * layers.L.conv.*, norm.*, conv_after_body.*, conv_before_upsample.0.*, conv_up{1,2,3}.*, conv_hr.*, conv_last.*); the
 * buffers relative_position_index / attn_mask are functions of the window size and are not loaded.
 * x, out: (B,3,H,W) fp32, x in [0,1]; H and W multiples of 64 (PixelUnshuffle 8 x window 8). */
typedef struct ir_swinir ir_swinir;
int ir_swinir_create(ir_swinir** out);
void ir_swinir_destroy(ir_swinir* h);
int ir_swinir_num_params(const ir_swinir* h);
int ir_swinir_param_info(const ir_swinir* h, int i, char* name, int name_cap, long long* numel);
int ir_swinir_load_param(ir_swinir* h, const char* name, const float* src_dev, long long numel, void* stream);
size_t ir_swinir_workspace_bytes(const ir_swinir* h, int B, int H, int W);
int ir_swinir_forward(ir_swinir* h, const float* x, float* out, int B, int H, int W, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ------------------------------------------------------------------ tile scheduler / pixel post-processing ---- */
/* coords: device int32 [ntiles][2] = (hi, wi) window origins from _sliding_windows (test_scripts/inference.py:40-53),
 * multiplied by `scale` inside the kernels (1 for latents, 8 for pixels).
 * gather: dst[t][n][c][y][x] = src[n][c][hi_t*scale + y][wi_t*scale + x]  (inference.py:130,140,144). */
int ir_tile_gather(const float* src, float* dst, const int32_t* coords, int ntiles, int N, int C, int H, int W, int th,
                   int tw, int scale, void* stream);
/* blend: out = (sum of covering tiles, in tile-list order) / cover count  (inference.py:133-136,151-153). */
int ir_tile_blend(const float* tiles, const int32_t* coords, int ntiles, float* out, int N, int C, int H, int W, int th,
                  int tw, int scale, void* stream);
size_t ir_wavelet_workspace_bytes(int N, int C, int H, int W);
/* out = high_freq(content) + low_freq(style), 5-level a-trous decomposition (utils/image/align_color.py:73-119). */
int ir_wavelet_reconstruction(const float* content, const float* style, float* out, int N, int C, int H, int W,
                              void* workspace, size_t workspace_bytes, void* stream);
/* adaptive_instance_normalization (utils/image/align_color.py:44-71). */
int ir_adain(const float* content, const float* style, float* out, int N, int C, int HW, void* stream);
/* (N,C,H,W) fp32 in [0,1] -> (N,H,W,C) uint8 with clamp and truncation (test_scripts/inference.py:159-160). */
int ir_to_uint8(const float* img, uint8_t* out, int N, int C, int H, int W, void* stream);

/* ------------------------------------------------------------------ unit entry points (parity tests) ---- */
/* out = epilogue(alpha * A[M,K] * W[N,K]^T + bias). epilogue: 0 bf16, 1 bf16+GELU(tanh), 2 fp32 (+gate, +resid). */
int ir_gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int batch, long long strideA,
                 long long strideW, long long strideO, int epilogue, float alpha, void* out_bf16, float* out_f32,
                 const float* resid_f32, const float* gate, long long gate_ld, int rows_per_gate, int force_bn,
                 void* stream);
/* 3x3 stride-1 pad-1 convolution on NHWC bf16: act (n,H,W,C), weight (Cout, 9*C) tap-major, out (n,H,W,Cout). */
int ir_conv3x3_bf16(const void* act, const void* weight, const float* bias, int n, int H, int W, int C, int Cout,
                    void* out_bf16, float* out_f32, const void* resid_bf16, const float* resid_f32, int force_bn,
                    void* stream);
/* Downsample.forward (ldm/modules/diffusionmodules/model.py:92-101): F.pad(x, (0,1,0,1)) + 3x3 stride-2 conv as an implicit
 * GEMM (the A tiles are gathered at pixel stride 2 by the TMA unit). act (n,2*Ho,2*Wo,C) NHWC bf16, weight (Cout, 9*C)
 * tap-major, out (n,Ho,Wo,Cout) NHWC bf16. */
int ir_conv3x3_s2_bf16(const void* act, const void* weight, const float* bias, int n, int Ho, int Wo, int C, int Cout,
                       void* out_bf16, int force_bn, void* stream);
/* 1x1 conv on the conv epilogue (nin_shortcut / proj_out / the gathered taps of Encoder.conv_in, model.py:120-128,
 * 178-179,456-460): act (n,H,W,C) NHWC bf16 (C a multiple of 64, or a multiple of 8 below 64), weight (Cout, C),
 * resid_bf16 (optional) and out (n,H,W,Cout) NHWC bf16. */
int ir_conv1x1_bf16(const void* act, const void* weight, const float* bias, int n, int H, int W, int C, int Cout,
                    void* out_bf16, const void* resid_bf16, int force_bn, void* stream);
/* One pass of the materialised single-head attention of AttnBlock (model.py:181-205) on the GEMM kernel, acc = A[M,K] W[N,K]^T:
 * mode 1: att_out[g * M + m] = max of acc over the 64-column group g of row m (no matrix output);
 * mode 2: out_bf16 = exp2(alpha * acc - att_row[m]), att_out[g * M + m] = fp32 sum of those exponentials over group g;
 * mode 3: out_bf16 = acc * att_row[m]. att_out holds ceil(N / 64) * M floats; ldo % 8 == 0. */
int ir_gemm_attn_pass(const void* A, const void* W, int M, int N, int K, long long lda, long long ldw, int mode, float alpha,
                      const float* att_row, float* att_out, void* out_bf16, long long ldo, int force_bn, void* stream);
/* Upsample.forward (ldm/modules/diffusionmodules/model.py:63-67): nearest x2 + 3x3 conv (C -> C) as four 2x2 phase convs
 * on the low-resolution input. act (n,H,W,C) NHWC bf16, weight_oihw (C,C,3,3) fp32 in the reference layout, phase_w_ws:
 * 16*C*C bf16 scratch for the pre-summed phase weights, out (n,2H,2W,C) NHWC bf16. */
int ir_upsample_conv3x3_bf16(const void* act, const float* weight_oihw, const float* bias, int n, int H, int W, int C,
                             void* phase_w_ws, void* out_bf16, int force_bn, void* stream);
int ir_attention_bf16(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldk,
                      long long ldv, long long ldo, int B, int heads, int head_dim, int Tq, int Tk,
                      const int32_t* kv_off, const int32_t* kv_len, float scale, void* stream);
/* qkv projection with the head-major scatter epilogue: A (M,K) bf16, W (3*H*hd, K) bf16 ->
 * q, k: [M/T][H][T][hd] bf16, vt: [M/T][H][hd][Tp] bf16. */
int ir_gemm_qkv_heads(const void* A, const void* W, const float* bias, int M, int K, int T, int Tp, int H, int hd,
                      void* q_heads, void* k_heads, void* vt_heads, int force_bn, void* stream);
/* tcgen05 self-attention on head-major operands (see ir_gemm_qkv_heads); out: (B*T, ldo) bf16. */
int ir_attention_tc_bf16(const void* q_heads, const void* k_heads, const void* vt_heads, void* out, long long ldo, int B,
                         int H, int head_dim, int T, int Tp, float scale, void* stream);
/* tcgen05 cross-attention to the packed caption (MultiHeadCrossAttention.forward, PixArt_blocks.py:43-58 with
 * BlockDiagonalMask.from_seqlens([T]*B, y_lens)): q (B*T, ldq) bf16 head-interleaved; kv (sum_l, ldkv) bf16 = kv_linear
 * output (K at columns [0, heads*hd), V at [heads*hd, 2*heads*hd)); sample b attends to rows [kv_off[b], +kv_len[b]);
 * max_l >= every (kv_off[b] % 8 + kv_len[b]) (host copy, <= 384); vt_ws: device scratch of ir_cross_attention_vt_bytes(heads, sum_l) bytes
 * that receives the transposed V half (inside ir_dit_forward this copy is made once per caption); out (B*T, ldo) bf16. */
size_t ir_cross_attention_vt_bytes(int heads, int sum_l);
int ir_cross_attention_tc_bf16(const void* q, const void* kv, void* vt_ws, void* out, long long ldq, long long ldkv,
                               long long ldo, int B, int heads, int head_dim, int T, int sum_l, const int32_t* kv_off,
                               const int32_t* kv_len, int max_l, float scale, void* stream);
/* Diagnostics: with a device buffer of the returned length (int64 elements) set, ir_attention_tc_bf16 runs an
 * instrumented instantiation that records SM-clock stamps of the warp roles of CTA (0,0,0); NULL switches it off. */
int ir_debug_attention_trace(long long* device_buf);
/* Diagnostics: in a library built with -DIR_DEBUG, CTA 0 of every tcgen05 GEMM / conv launch writes %globaltimer stamps of
 * its warp roles into record (launch index % slots) of device_buf (slots x 16 int64). Returns the record length in int64
 * elements (16), or 0 in a release build, which carries no trace code. NULL switches the trace off. */
int ir_debug_gemm_trace(long long* device_buf, int slots);
int ir_ln_modulate(const float* x, void* out_bf16, const float* shift, const float* scale, long long mod_stride,
                   int rows, int T, int D, void* stream);
int ir_pos_embed(float* table, int gh, int gw, int D, int base_size, float pe_interpolation, void* stream);
/* tokens (B*T, D) fp32 = PatchEmbed(x) + pos (PixArtMS.py:22-46, pixart_controlnet.py:78-87); used by forward_c. */
int ir_dit_patch_embed(ir_dit* h, const float* x, float* tokens, int B, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INSTAREVIVE_B200_H_ */
