// Host-side interface of the fused attention kernel (attention.cu).
#pragma once
#include "gemm.cuh"

namespace ir {

struct AttnArgs {
  const bf16* q = nullptr;  // [B*Tq][ldq], head h occupies columns [h*head_dim, (h+1)*head_dim)
  const bf16* k = nullptr;  // [rows][ldk]
  const bf16* v = nullptr;  // [rows][ldv]
  bf16* out = nullptr;      // [B*Tq][ldo]
  long ldq = 0, ldk = 0, ldv = 0, ldo = 0;
  int B = 0, heads = 0, head_dim = 0;
  int Tq = 0;
  int Tk = 0;                   // uniform keys per sample (used when kv_len == nullptr)
  const int* kv_off = nullptr;  // device [B], optional: first kv row of each sample
  const int* kv_len = nullptr;  // device [B], optional: kv rows of each sample
  float scale = 1.0f;
};

int attention_launch(const AttnArgs& a, cudaStream_t stream);

// tcgen05 self-attention on head-major operands produced by the qkv GEMM (EPI_QKV):
// q, k: [B][H][T][head_dim] bf16; vt: [B][H][head_dim][Tp] bf16 (V transposed); out: [B*T][ldo] bf16, head-interleaved.
struct AttnTcArgs {
  const bf16* q = nullptr;
  const bf16* k = nullptr;
  const bf16* vt = nullptr;
  bf16* out = nullptr;
  long ldo = 0;
  int B = 0, H = 0, head_dim = 0, T = 0, Tp = 0;
  float scale = 1.0f;
};
int attention_tc_launch(const AttnTcArgs& a, cudaStream_t stream);
// tcgen05 cross-attention to the packed caption (xattention_tc.cu): q [B*T][ldq] head-interleaved; kv [sumL][ldkv] the
// kv_linear output of one block (K at columns [0, H*hd), V at [H*hd, 2*H*hd)); sample b attends to rows
// [kv_off[b], kv_off[b] + kv_len[b]). max_len = max_b (kv_off[b] % 8 + kv_len[b]) (the key window: TMA starts at the
// packed row rounded down to a multiple of 8) and kv_total = sum_b kv_len[b] are host copies (box size, FLOP accounting); vt = the V half transposed by xattention_transpose_v ([H][72][roundup8(sumL)], keys contiguous);
// out [B*T][ldo].
struct XAttnTcArgs {
  const bf16* q = nullptr;
  const bf16* kv = nullptr;
  const bf16* vt = nullptr;
  bf16* out = nullptr;
  long ldq = 0, ldkv = 0, ldo = 0;
  int B = 0, H = 0, head_dim = 0, T = 0, sumL = 0;
  const int* kv_off = nullptr;
  const int* kv_len = nullptr;
  int max_len = 0;
  long kv_total = 0;
  float scale = 1.0f;
};
int xattention_tc_launch(const XAttnTcArgs& a, cudaStream_t stream);
// transposed copy of the V half of nblk consecutive kv_linear outputs [nblk][sumL][ldkv] -> [nblk][H][72][roundup8(sumL)]
// (once per caption); xattention_vt_elems = elements per block
long xattention_vt_elems(int H, int sumL);
int xattention_transpose_v(const bf16* kv, bf16* vt, int nblk, int H, int sumL, long ldkv, cudaStream_t stream);

// diagnostics: when a device buffer of attention_tc_trace_len() int64 is set, attention_tc_launch runs the instrumented
// instantiation, which records SM-clock stamps of the roles of CTA (0,0,0) (tools/gpu_attn_probe.py --trace)
void attention_tc_set_trace(long long* device_buf);
int attention_tc_trace_len();

}  // namespace ir
