// VAE (decoder, optional encoder) handle and launchers (vae.cu).
#pragma once
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "gemm.cuh"

namespace ir {

struct VaeConfig {
  int ch = 128;                   // configs/cldm.yaml:69-84
  int z_channels = 4;
  int out_ch = 3;
  int num_res_blocks = 2;
  int ch_mult[4] = {1, 2, 4, 4};
  int with_encoder = 0;           // also register encoder.* / quant_conv.* (Encoder, model.py:440-546)
};

enum VaeParamKind { VP_F32 = 0, VP_CONV_BF16 = 1, VP_CONVIN_F32 = 2, VP_CONVOUT_F32 = 3 };

struct VaeParam {
  std::string name;  // reference state_dict key ("post_quant_conv.*", "decoder.*")
  int kind;
  long numel;
  int cout, cin, k;
  long offset;
  bool loaded = false;
};

// one cached CUDA graph of a decode (kind 0) / encode (kind 1) call; exec == nullptr: the key was seen once (eager run)
struct VaeGraph {
  int kind = 0, B = 0, H = 0, W = 0;
  const void* ws = nullptr;
  float f[3] = {0.f, 0.f, 0.f};   // decode: in_scale, out_scale, out_shift (baked into the captured kernels' arguments)
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
  bool same_key(const VaeGraph& o) const {
    return kind == o.kind && B == o.B && H == o.H && W == o.W && ws == o.ws && f[0] == o.f[0] && f[1] == o.f[1] && f[2] == o.f[2];
  }
};

struct Vae {
  VaeConfig cfg;
  std::vector<VaeGraph> graphs;
  cudaStream_t cap_stream = nullptr;
  bool graphs_enabled = true;
  std::vector<VaeParam> params;
  std::unordered_map<std::string, int> index;
  bf16* wb = nullptr;
  float* wf = nullptr;
  long wb_elems = 0, wf_elems = 0;
  int first_encoder_param = -1;   // params[first_encoder_param..] belong to the encoder (decode does not need them)
  // decoder.conv_out as a tap-response GEMM: (32, C) bf16 weight rows [tap*3 + o][c] (rows 27..31 zero) and the bias
  bf16* co_w = nullptr;
  float* co_b = nullptr;
  // encoder.conv_in as an im2col GEMM: (ch, 32) bf16 weight rows, column (ky*3+kx)*3 + c (columns 27..31 zero)
  bf16* ci_w = nullptr;
  // decoder.up.N.upsample.conv folded with the nearest x2 upsample: per conv four phase matrices (Cout, 2, 2, Cin) bf16
  std::unordered_map<std::string, bf16*> up_w;
};

int vae_create(const VaeConfig& cfg, Vae** out);
void vae_destroy(Vae* v);
int vae_load_param(Vae* v, const char* name, const float* src_dev, long numel, cudaStream_t s);
void vae_set_graphs(Vae* v, bool on);   // CUDA-graph replay of decode / encode calls (default on)
size_t vae_workspace_bytes(const Vae* v, int B, int h, int w);
// z: (B,4,h,w) fp32 -> out: (B,3,8h,8w) fp32 = decode(z * in_scale) * out_scale + out_shift
int vae_decode(Vae* v, const float* z, float* out, int B, int h, int w, float in_scale, float out_scale,
               float out_shift, void* workspace, size_t workspace_bytes, cudaStream_t s);

// x: (B,3,H,W) fp32 in [-1,1] -> moments: (B, 2z, H/8, W/8) fp32 = quant_conv(Encoder(x)) (mean | logvar)
size_t vae_encode_workspace_bytes(const Vae* v, int B, int H, int W);
int vae_encode(Vae* v, const float* x, float* moments, int B, int H, int W, void* workspace, size_t workspace_bytes,
               cudaStream_t s);

// Upsample.forward (model.py:63-67), nearest x2 + 3x3 conv, as four 2x2 phase convs on the low-resolution input:
// pack_upconv_phases sums the (Cout, Cin, 3, 3) fp32 taps into [phase][Cout][2][2][Cin] bf16 (16*Cout*Cin elements);
// the launcher maps x (n,H,W,C) NHWC bf16 to y (n,2H,2W,C). gn_partial (optional): fused GroupNorm partials,
// 4 * gemm_conv_tiles_per_image(H, W) slots of 64 floats per image.
int pack_upconv_phases(const float* w_oihw, bf16* phase_w, int cout, int cin, cudaStream_t s);
int upsample_conv_phases_launch(const bf16* x, const bf16* phase_w, const float* bias, bf16* y, int n, int H, int W, int C,
                                float* gn_partial, int force_bn, cudaStream_t s);

}  // namespace ir
