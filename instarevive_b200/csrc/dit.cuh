// DiT + ControlNet-Half model handle: packed weights on the device and the forward schedule (dit.cu).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "attention.cuh"
#include "elementwise.cuh"

namespace ir {

struct DitConfig {
  int depth = 28;           // PixArtMS_XL_2: PixArtMS.py:291-293
  int copy_blocks = 13;     // ControlPixArtMSHalf default, pixart_controlnet.py:188
  int hidden = 1152;
  int heads = 16;
  int patch = 2;
  int in_ch = 4;
  int out_ch = 8;           // learn_sigma / pred_sigma: 2 * in_ch
  int caption_ch = 4096;
  int mlp_ratio = 4;
  int base_size = 32;       // input_size // patch_size, PixArt.py:100
  float pe_interpolation = 1.0f;
};

enum ParamKind { PK_BF16 = 0, PK_F32 = 1, PK_F32_TRANSPOSED = 2 };

struct ParamEntry {
  std::string name;   // reference state_dict key
  int kind;
  long numel;
  int rows, cols;     // logical (out, in) shape for matrices; rows = numel, cols = 1 for vectors
  long offset;        // element offset into the bf16 or fp32 blob
  bool loaded = false;
};

struct BlockW {
  const bf16 *qkv, *proj, *q_lin, *cproj, *fc1, *fc2;
  const float *b_qkv, *b_proj, *b_q, *b_kv, *b_cproj, *b_fc1, *b_fc2;
};

struct DitGraphKey {
  int B = 0, H = 0, W = 0, sumL = 0, max_len = 0, has_c = 0;
  const void *ws = nullptr, *pos = nullptr, *ykv = nullptr, *ykv_t = nullptr;
  bool operator==(const DitGraphKey& o) const {
    return B == o.B && H == o.H && W == o.W && sumL == o.sumL && max_len == o.max_len && has_c == o.has_c && ws == o.ws &&
           pos == o.pos && ykv == o.ykv && ykv_t == o.ykv_t;
  }
};
struct DitGraph {
  DitGraphKey key;
  cudaGraphExec_t exec = nullptr;   // null: the key was seen once (eager run), capture on the next call
  int launches = 0;                 // kernel launches one replay stands for (ir_launch_count accounting)
};

struct Dit {
  DitConfig cfg;
  int nblk = 0;  // depth + copy_blocks
  std::vector<ParamEntry> params;
  std::unordered_map<std::string, int> index;
  bf16* wb = nullptr;   // bf16 blob
  float* wf = nullptr;  // fp32 blob
  long wb_elems = 0, wf_elems = 0;

  // resolved pointers
  std::vector<BlockW> blocks;      // [depth] base, then [copy_blocks] control
  const bf16* kv_all = nullptr;    // [nblk][2D][D]
  const float* b_kv_all = nullptr; // [nblk][2D]
  const float* tables = nullptr;   // [nblk][6][D]
  const bf16* before_proj = nullptr;
  const float* b_before = nullptr;
  std::vector<const bf16*> after_proj;
  std::vector<const float*> b_after;
  const float *xw_t, *xb;                      // x_embedder (transposed [C*4][D]) and bias
  const float *t_w0, *t_b0, *t_w2, *t_b2;      // t_embedder.mlp
  const float *cs_w0, *cs_b0, *cs_w2, *cs_b2;  // csize_embedder.mlp
  const float *ar_w0, *ar_b0, *ar_w2, *ar_b2;  // ar_embedder.mlp
  const float *tb_w, *tb_b;                    // t_block.1
  const bf16 *y_fc1, *y_fc2;
  const float *y_b1, *y_b2;
  const float *fin_table, *fin_w, *fin_b;

  // caches owned by the handle
  float* pos = nullptr;  // (gh*gw, D) table for the last (gh, gw); grow-only buffer
  long pos_cap = 0;      // capacity in elements
  int pos_gh = 0, pos_gw = 0;
  cudaStream_t side = nullptr;       // control-chain stream of the dual-chain schedule (created on first use)
  cudaEvent_t ev_fork = nullptr;     // base block 0 done -> the control chain may start
  std::vector<cudaEvent_t> ev_c;     // [copy_blocks] control block i done -> the base chain may inject c_i
  bool dual_chain = true;            // control chain on a second stream (dit_forward); off = one stream, program order
  bool graphs_enabled = true;        // replay the forward as a CUDA graph from the third call with a key on
  std::vector<DitGraph> graphs;
  cudaStream_t cap_stream = nullptr; // capture happens on this stream (the caller's may be the uncapturable legacy stream)
  bf16* ykv = nullptr;   // [nblk][sumL][2D] caption K/V of the last caption
  long ykv_cap = 0;      // capacity in elements
  bf16* ykv_t = nullptr; // [nblk][H][72][roundup8(sumL)] V halves transposed (keys contiguous) for the cross-attention TMA
  long ykv_t_cap = 0;
  int ykv_sumL = -1;
};

int dit_create(const DitConfig& cfg, Dit** out);
void dit_destroy(Dit* d);
int dit_load_param(Dit* d, const char* name, const float* src_dev, long numel, cudaStream_t s);
size_t dit_workspace_bytes(const Dit* d, int B, int H, int W, int sumL);

struct DitForwardArgs {
  const float* x = nullptr;         // (B, in_ch, H, W)
  const float* c = nullptr;         // (B, in_ch, H, W) or null -> plain base path
  const float* timestep = nullptr;  // (B)
  const float* y = nullptr;         // (rows, caption_ch) caption embeddings (all tokens, padded)
  const int* y_index = nullptr;     // device (sumL): rows of y that are valid, sample-major
  const int* kv_off = nullptr;      // device (B): first packed row of each sample
  const int* kv_len = nullptr;      // device (B): valid tokens of each sample
  const float* img_hw = nullptr;    // device (B, 2)
  const float* aspect = nullptr;    // device (B)
  float* out = nullptr;             // (B, out_ch, H, W)
  int B = 0, H = 0, W = 0, sumL = 0;
  int max_len = 0;                  // host copy of max_b (kv_off[b] % 8 + kv_len[b]) (<= 384): the cross-attention key window
  long kv_total = 0;                // host copy of sum_b kv_len[b] (FLOP accounting only)
  int reuse_caption = 0;            // 1: caption K/V of the previous call are still valid
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
};

int dit_forward(Dit* d, const DitForwardArgs& a, cudaStream_t s);
// pre-size the handle-owned caches (position table for up to max_tokens tokens, caption K/V for up to max_sum_l packed
// caption tokens) so that no forward allocates
int dit_reserve(Dit* d, int max_tokens, int max_sum_l);
void dit_set_graphs(Dit* d, bool on);
void dit_set_dual_chain(Dit* d, bool on);
int dit_patch_embed(Dit* d, const float* x, float* tokens, int B, int H, int W, cudaStream_t s);

}  // namespace ir
