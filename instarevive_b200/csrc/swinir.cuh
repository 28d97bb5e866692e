// Stage-1 SwinIR (diffusion/model/swinir.py, configs/swinir.yaml) handle and launcher (swinir.cu) -- SURVEY 8f row 2.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "gemm.cuh"

namespace ir {

struct SwinConfig {
  int embed_dim = 180;   // configs/swinir.yaml
  int num_layers = 8;    // RSTBs
  int depth = 6;         // SwinTransformerBlocks per RSTB
  int heads = 6;
  int window = 8;
  int mlp_ratio = 2;
  int sf = 8;            // PixelUnshuffle factor == upscale of the 'nearest+conv' upsampler
  int num_feat = 64;
};

enum SwinParamKind { SP_F32 = 0, SP_LINEAR = 1, SP_CONV = 2, SP_RELBIAS = 3 };

struct SwinParam {
  std::string name;
  int kind;
  long numel;          // elements of the reference tensor
  int cout, cin, k;    // linear: (cout, cin, 1); conv: (cout, cin, 3)
  long offset;         // into wb (SP_LINEAR / SP_CONV) or wf
  bool loaded = false;
};

struct Swin {
  SwinConfig cfg;
  std::vector<SwinParam> params;
  std::unordered_map<std::string, int> index;
  bf16* wb = nullptr;
  float* wf = nullptr;
  long wb_elems = 0, wf_elems = 0;
};

int swin_create(const SwinConfig& cfg, Swin** out);
void swin_destroy(Swin* s);
int swin_load_param(Swin* s, const char* name, const float* src_dev, long numel, cudaStream_t st);
size_t swin_workspace_bytes(const Swin* s, int B, int H, int W);
// x: (B,3,H,W) fp32 in [0,1], H and W multiples of sf*window -> out: (B,3,H,W) fp32 (SwinIR.forward, swinir.py:867-905)
int swin_forward(Swin* s, const float* x, float* out, int B, int H, int W, void* workspace, size_t workspace_bytes,
                 cudaStream_t st);

}  // namespace ir
