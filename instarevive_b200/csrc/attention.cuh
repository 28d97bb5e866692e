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

}  // namespace ir
