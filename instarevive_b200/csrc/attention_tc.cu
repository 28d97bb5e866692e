// tcgen05 / TMEM flash attention for the DiT self-attention (AttentionKVCompress.forward with sr_ratio 1,
// diffusion/model/nets/PixArt_blocks.py:123-158: softmax(q k^T / sqrt(72)) v, 16 heads, no mask).
//
// One CTA owns 256 query rows of one (sample, head) as two 128-row tiles that ping-pong on the tensor core:
//   warp 0      TMA producer: Q tiles once, then a ring of K / V^T tiles (128 keys each)
//   warp 1      MMA issuer (one thread): S_q = Q_q K^T into TMEM, O_q += P_q V into TMEM
//   warps 2-5   softmax of query tile 0, warps 6-9 softmax of query tile 1 (one thread per query row, no shuffles):
//               tcgen05.ld the S row, online max with lazy rescaling of the O accumulator (tcgen05.ld/st, only when the
//               running max grew by more than 2^8), exp2, bf16 P written to shared memory in the 128B-swizzled K-major
//               layout the P*V MMA consumes.
// head_dim 72 is handled without padding the data in HBM: q / k are stored head-major [B][H][T][72] and v transposed
// [B][H][72][Tp] by the qkv GEMM epilogue (EPI_QKV); TMA boxes read 64 + 16 columns and the tensor-map bounds make the
// hardware zero-fill columns 72..79 (and keys / rows beyond T), so QK^T runs 5 k-steps of 16 and P*V is an N = 80 MMA.
#include "attention.cuh"

namespace ir {

namespace {

constexpr int HD = 72;
constexpr int BQ = 128;          // rows per query tile (UMMA M)
constexpr int BKV = 128;         // keys per tile (UMMA N of S, K of P*V)
constexpr int NV = 80;           // head dim padded to a multiple of 16 (UMMA N of P*V)
constexpr int STAGES = 3;
constexpr int Q64_BYTES = BQ * 64 * 2;    // 16384
constexpr int Q16_BYTES = BQ * 16 * 2;    // 4096
constexpr int QTILE_BYTES = Q64_BYTES + Q16_BYTES;
constexpr int K64_BYTES = BKV * 64 * 2;   // 16384
constexpr int K16_BYTES = BKV * 16 * 2;   // 4096
constexpr int VT_ATOM_BYTES = NV * 64 * 2;  // 10240: [80 rows (d)][64 keys]
constexpr int STAGE_BYTES = K64_BYTES + K16_BYTES + 2 * VT_ATOM_BYTES;  // 40960
constexpr int P_BYTES = BQ * BKV * 2;     // 32768: two [128][64] atoms
constexpr int SMEM_BYTES = 1024 + 2 * QTILE_BYTES + STAGES * STAGE_BYTES + 2 * P_BYTES + 256;
constexpr int NTHREADS = 320;   // warp 0 TMA (+TMEM alloc), warp 1 MMA, warps 2-9 softmax (2 query tiles x 4 lane quarters)
constexpr int TMEM_S0 = 0, TMEM_S1 = 128, TMEM_O0 = 256, TMEM_O1 = 384;

struct AttnTcDev {
  bf16* out;
  long ldo;
  int T;        // tokens per sample (queries = keys)
  int H;
  float scale_log2e;
};

IR_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace

__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ64, const __grid_constant__ CUtensorMap tmQ16,
               const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
               const __grid_constant__ CUtensorMap tmVT, const AttnTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                  // [2][Q64 | Q16]
  uint8_t* sKV = sQ + 2 * QTILE_BYTES;                 // [STAGES][K64 | K16 | VT0 | VT1]
  uint8_t* sP = sKV + STAGES * STAGE_BYTES;            // [2][P atom0 | P atom1]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * P_BYTES);
  uint64_t* q_full = bars;                  // [1]
  uint64_t* kv_full = bars + 1;             // [STAGES]
  uint64_t* kv_empty = bars + 1 + STAGES;   // [STAGES]
  uint64_t* s_full = bars + 1 + 2 * STAGES;   // [2]
  uint64_t* p_full = s_full + 2;              // [2]
  uint64_t* o_full = p_full + 2;              // [2]
  uint64_t* s_free = o_full + 2;              // [2] softmax has copied S into registers: S buffer may be overwritten
  uint64_t* pv_done = s_free + 2;             // [2] P*V of the previous tile finished: P smem / O TMEM may be touched
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.T + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ64);
    tma_prefetch_desc(&tmQ16);
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmK16);
    tma_prefetch_desc(&tmVT);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], BQ);
      mbar_init(&o_full[i], 1);
      mbar_init(&s_free[i], BQ);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * QTILE_BYTES);
      for (int qt = 0; qt < 2; ++qt) {
        tma_load_4d(sQ + qt * QTILE_BYTES, &tmQ64, q_full, 0, q0 + qt * BQ, head, b);
        tma_load_4d(sQ + qt * QTILE_BYTES + Q64_BYTES, &tmQ16, q_full, 64, q0 + qt * BQ, head, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        uint8_t* st = sKV + stage * STAGE_BYTES;
        mbar_arrive_expect_tx(&kv_full[stage], STAGE_BYTES);
        tma_load_4d(st, &tmK64, &kv_full[stage], 0, j * BKV, head, b);
        tma_load_4d(st + K64_BYTES, &tmK16, &kv_full[stage], 64, j * BKV, head, b);
        tma_load_4d(st + K64_BYTES + K16_BYTES, &tmVT, &kv_full[stage], j * BKV, 0, head, b);
        tma_load_4d(st + K64_BYTES + K16_BYTES + VT_ATOM_BYTES, &tmVT, &kv_full[stage], j * BKV + 64, 0, head, b);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, NV);
      auto issue_s = [&](int qt, int stage) {
        const uint32_t qa = smem_u32(sQ + qt * QTILE_BYTES);
        const uint32_t ka = smem_u32(sKV + stage * STAGE_BYTES);
        const uint64_t dq = make_smem_desc_sw128(qa), dk = make_smem_desc_sw128(ka);
        const uint32_t td = tmem_base + (qt == 0 ? TMEM_S0 : TMEM_S1);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(td, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_bf16(td, make_smem_desc_sw32(qa + Q64_BYTES), make_smem_desc_sw32(ka + K64_BYTES), idesc_s, 1);
        umma_commit(&s_full[qt]);
      };
      auto issue_pv = [&](int qt, int stage, bool accumulate) {
        const uint32_t pa = smem_u32(sP + qt * P_BYTES);
        const uint32_t va = smem_u32(sKV + stage * STAGE_BYTES + K64_BYTES + K16_BYTES);
        const uint32_t td = tmem_base + (qt == 0 ? TMEM_O0 : TMEM_O1);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          const uint64_t dp = make_smem_desc_sw128(pa + (kk >> 2) * (BQ * 128)) + (uint64_t)(2 * (kk & 3));
          const uint64_t dv = make_smem_desc_sw128(va + (kk >> 2) * VT_ATOM_BYTES) + (uint64_t)(2 * (kk & 3));
          umma_bf16(td, dp, dv, idesc_o, (accumulate || kk != 0) ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      // Event-driven issue: per query tile, S(t) may go as soon as the softmax warps have copied S(t-1) out of TMEM
      // (s_free) and K(t) has landed; P*V(t) as soon as P(t) is in shared memory (p_full). Polling both query tiles
      // keeps the tensor core fed whichever softmax group finishes first.
      int s_next[2] = {1, 1}, pv_next[2] = {0, 0};
      int kv_seen = 0;       // tiles [0, kv_seen] have landed
      int kv_released = 0;   // tiles [0, kv_released) have been handed back to the producer
      while (pv_next[0] < n_tiles || pv_next[1] < n_tiles) {
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          const int ts = s_next[qt];
          if (ts < n_tiles) {
            if (ts > kv_seen && mbar_try_wait(&kv_full[ts % STAGES], (uint32_t)((ts / STAGES) & 1))) kv_seen = ts;
            if (ts <= kv_seen && mbar_try_wait(&s_free[qt], (uint32_t)((ts - 1) & 1))) {
              tc_fence_after();
              issue_s(qt, ts % STAGES);
              s_next[qt] = ts + 1;
            }
          }
          const int tp = pv_next[qt];
          if (tp < n_tiles && mbar_try_wait(&p_full[qt], (uint32_t)(tp & 1))) {
            tc_fence_after();
            issue_pv(qt, tp % STAGES, tp > 0);
            if (tp + 1 < n_tiles)
              umma_commit(&pv_done[qt]);   // softmax(qt, tp+1) may overwrite P_qt / rescale O_qt once this fires
            else
              umma_commit(&o_full[qt]);
            pv_next[qt] = tp + 1;
            const int done = pv_next[0] < pv_next[1] ? pv_next[0] : pv_next[1];
            while (kv_released < done) {
              umma_commit(&kv_empty[kv_released % STAGES]);   // both P*V(kv_released) issued: stage can be refilled
              ++kv_released;
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: one thread per query row
    const int qt = (warp - 2) >> 2;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t t_s = tmem_base + ((uint32_t)(quarter * 32) << 16) + (qt == 0 ? TMEM_S0 : TMEM_S1);
    const uint32_t t_o = tmem_base + ((uint32_t)(quarter * 32) << 16) + (qt == 0 ? TMEM_O0 : TMEM_O1);
    uint8_t* prow = sP + qt * P_BYTES + r * 128;
    const int rx = r & 7;
    float m_used = -INFINITY, l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[qt], (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, s[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&s_free[qt]);   // S lives in registers now: the tensor core may start S(qt, j+1)
      const int kbase = j * BKV;
      float mx = -INFINITY;
      if (kbase + BKV <= p.T) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[c][i]));
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (kbase + c * 32 + i >= p.T) s[c][i] = __float_as_uint(-INFINITY);
            mx = fmaxf(mx, __uint_as_float(s[c][i]));
          }
      }
      const float m_new = fmaxf(m_used, mx * p.scale_log2e);
      if (j > 0) {
        // P*V(qt, j-1) must have finished reading P_qt and accumulating into O_qt before either is touched again
        mbar_wait(&pv_done[qt], (uint32_t)((j - 1) & 1));
        tc_fence_after();
      }
      // lazy rescale: keep the stale max unless it grew by more than 8 (p stays below 2^8, exact in fp32 / fine in bf16)
      const bool grow = (m_new - m_used) > 8.0f;   // also true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float alpha = grow ? fast_exp2(m_used - m_new) : 1.0f;
        if (grow) m_used = m_new;
        l *= alpha;
        if (j > 0) {
#pragma unroll
          for (int c = 0; c < NV / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(t_o + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_o + c * 16, o);
          }
          tmem_st_wait();
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = fast_exp2(__uint_as_float(s[c][g * 8 + i]) * p.scale_log2e - m_used);
            sum += e[i];
          }
          const int jc = c * 4 + g;   // 16-byte chunk (8 keys) index inside the 128-key row
          uint4 u = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]),
                               pack_bf16x2(e[6], e[7]));
          *reinterpret_cast<uint4*>(prow + (jc >> 3) * (BQ * 128) + (((jc & 7) ^ rx) << 4)) = u;
        }
      }
      l += sum;
      fence_proxy_async();   // make the generic-proxy smem writes visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(&p_full[qt]);
    }
    // ---- epilogue: O / l -> bf16 -> out[(b*T + row)][head*72 + d]
    mbar_wait(&o_full[qt], 0);
    tc_fence_after();
    const int row = q0 + qt * BQ + r;
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    uint32_t o[32];
    bf16* og = p.out + ((long)b * p.T + row) * p.ldo + head * HD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
      if (row < p.T) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u = make_uint4(pack_bf16x2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv),
                               pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv),
                               pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv),
                               pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv));
          *reinterpret_cast<uint4*>(og + c * 32 + g * 8) = u;
        }
      }
    }
    uint32_t o2[16];
    tmem_ld_32x16(t_o + 64, o2);
    tmem_ld_wait();
    if (row < p.T) {
      uint4 u = make_uint4(pack_bf16x2(__uint_as_float(o2[0]) * inv, __uint_as_float(o2[1]) * inv),
                           pack_bf16x2(__uint_as_float(o2[2]) * inv, __uint_as_float(o2[3]) * inv),
                           pack_bf16x2(__uint_as_float(o2[4]) * inv, __uint_as_float(o2[5]) * inv),
                           pack_bf16x2(__uint_as_float(o2[6]) * inv, __uint_as_float(o2[7]) * inv));
      *reinterpret_cast<uint4*>(og + 64) = u;   // columns 64..71; accumulator columns 72..79 are padding
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attention_tc_launch(const AttnTcArgs& a, cudaStream_t stream) {
  IR_REQUIRE(a.head_dim == HD, "attention_tc: head_dim %d unsupported (kernel is specialised for %d)", a.head_dim, HD);
  IR_REQUIRE(a.q && a.k && a.vt && a.out && a.B > 0 && a.H > 0 && a.T > 0, "attention_tc: bad arguments");
  IR_REQUIRE(a.Tp % 8 == 0 && a.Tp >= a.T && a.ldo % 8 == 0, "attention_tc: Tp must be a multiple of 8 and >= T");
  CUtensorMap mq64, mq16, mk64, mk16, mvt;
  {
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)a.T, (uint64_t)a.H, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)a.T * HD * 2, (uint64_t)a.H * a.T * HD * 2};
    const uint32_t box64[4] = {64, BQ, 1, 1};
    const uint32_t box16[4] = {16, BQ, 1, 1};
    IR_TRY(make_tensor_map(&mq64, a.q, 4, dims, strides, box64, 128));
    IR_TRY(make_tensor_map(&mq16, a.q, 4, dims, strides, box16, 32));
    IR_TRY(make_tensor_map(&mk64, a.k, 4, dims, strides, box64, 128));
    IR_TRY(make_tensor_map(&mk16, a.k, 4, dims, strides, box16, 32));
  }
  {
    const uint64_t dims[4] = {(uint64_t)a.T, (uint64_t)HD, (uint64_t)a.H, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)a.Tp * 2, (uint64_t)HD * a.Tp * 2, (uint64_t)a.H * HD * a.Tp * 2};
    const uint32_t box[4] = {64, NV, 1, 1};
    IR_TRY(make_tensor_map(&mvt, a.vt, 4, dims, strides, box, 128));
  }
  static bool configured = false;
  if (!configured) {
    IR_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  AttnTcDev p;
  p.out = a.out;
  p.ldo = a.ldo;
  p.T = a.T;
  p.H = a.H;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  dim3 grid((a.T + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
  const bool prof = prof_enabled();
  if (prof) prof_before(stream);
  IR_CUDA_CHECK(launch_pdl(attn_tc_kernel, grid, dim3(NTHREADS), SMEM_BYTES, stream, mq64, mq16, mk64, mk16, mvt, p));
  if (prof) prof_after(stream, PROF_ATTN, 4.0 * a.B * a.H * (double)a.T * a.T * HD);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

}  // namespace ir
