// tcgen05 / TMEM flash attention for the DiT self-attention (AttentionKVCompress.forward with sr_ratio 1,
// diffusion/model/nets/PixArt_blocks.py:123-158: softmax(q k^T / sqrt(72)) v, 16 heads, no mask).
//
// One CTA owns 256 query rows of one (sample, head) as two 128-row tiles that alternate on the tensor core.
// Shared memory holds only the K / V^T ring: a single-CTA SS MMA with N = 128 reads 8 KB of operands per 64-cycle
// instruction, i.e. the whole 128 B/clk shared-memory port, so both A operands live in TMEM instead (TS MMAs):
//   Q   written once per CTA into TMEM by the softmax threads (bf16 pairs, 40 columns, d 72..79 zero)
//   P   written by the softmax threads over the first 64 columns of their own S accumulator (bf16 pairs)
//   warp 0      TMA producer: rings of K tiles and V^T tiles (128 keys each; separate barriers, K runs one tile ahead)
//   warp 1      MMA issuer (one thread), program order  P V_0(j), S_0(j+1), P V_1(j), S_1(j+1): the tensor pipe executes
//               in issue order, so S(j+1) may be queued right behind the P V(j) that still reads P out of the same columns
//   warps 2-3   idle (keep the softmax warps warpgroup-aligned for setmaxnreg)
//   warps 4-7   softmax of query tile 0, warps 8-11 softmax of query tile 1 (one thread per query row, no shuffles):
//               tcgen05.ld the S row, online max with lazy rescaling of the O accumulator (tcgen05.ld/st, only when the
//               running max grew by more than 2^8), exp2 with packed FFMA2 / FADD2 row math (optionally a compile-time
//               share of the exponentials on the FMA pipe: Cody-Waite split + degree-3 minimax polynomial, relative
//               error 7.5e-5, far below the bf16 rounding of P), tcgen05.st of the bf16 P row.
// The softmax denominator is accumulated by the tensor core: shared-memory row 72 of every V^T tile is a row of ones (one
// of the 8 padding rows of the N = 80 P V MMA; the TMA box covers the 72 real rows only), so O[:, 72] = sum_k P[:, k] in
// fp32, rescaled with O and consistent with the bf16 P the numerator sees (sharp-softmax error 0.025 -> 0.016).
// setmaxnreg moves registers from warps 0-3 to the softmax warps (S row = 128 live registers).
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
constexpr int STAGES = 4;
constexpr int K64_BYTES = BKV * 64 * 2;   // 16384
constexpr int K16_BYTES = BKV * 16 * 2;   // 4096
constexpr int VT_ATOM_BYTES = NV * 64 * 2;  // 10240: [80 rows (d)][64 keys]
constexpr int STAGE_BYTES = K64_BYTES + K16_BYTES + 2 * VT_ATOM_BYTES;  // 40960
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256;
constexpr int NTHREADS = 384;   // warp 0 TMA (+TMEM alloc), warp 1 MMA, warps 2-3 idle, warps 4-11 softmax (2 tiles x 4 lane quarters)
// setmaxnreg moves registers inside the CTA's launch allocation (384 threads x 168 = 64512 registers), so
// 128 * REGS_CTRL + 256 * REGS_SOFTMAX must not exceed that or the last setmaxnreg.inc never returns.
constexpr int REGS_CTRL = 48, REGS_SOFTMAX = 224;
static_assert(128 * REGS_CTRL + 256 * REGS_SOFTMAX <= NTHREADS * 168, "setmaxnreg budget");
// TMEM columns: S (fp32, P aliased over its first 64 columns), O (fp32, 80 of them), Q (bf16 pairs, 40 used of 48)
constexpr int TMEM_S0 = 0, TMEM_S1 = 128, TMEM_O0 = 256, TMEM_O1 = 336, TMEM_Q0 = 416, TMEM_Q1 = 464;

struct AttnTcDev {
  const bf16* q;   // [B][H][T][72]
  bf16* out;
  long ldo;
  int T;        // tokens per sample (queries = keys)
  int H;
  float scale_log2e;
  float inv_scale_log2e;
  long long* trace;   // diagnostics (TRACE instantiation only): [4 roles][TRACE_ITERS][8] SM-clock stamps of CTA (0,0,0)
};
constexpr int TRACE_ITERS = 64;

IR_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 issue once for two lanes of data)
IR_DEVINL uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
IR_DEVINL void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
IR_DEVINL uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
IR_DEVINL uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
template <int N>
IR_DEVINL void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
IR_DEVINL void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// 2^x for a packed pair on the FMA pipe. x must be >= -126 (callers clamp the scores) and < 2^21.
// t = x + 1.5*2^23 rounds x to the nearest integer n and leaves n in the low mantissa bits of t; f = x - n in
// [-0.5, 0.5]; 2^f by a degree-3 minimax polynomial; the exponent is applied by adding n << 23 to the bit pattern.
IR_DEVINL void exp2_poly2(uint64_t x, float& e0, float& e1) {
  const uint64_t magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f);
  const uint64_t neg1 = pack2(-1.f, -1.f);
  const uint64_t c0 = pack2(0.9999280571937561f, 0.9999280571937561f), c1 = pack2(0.6932609677314758f, 0.6932609677314758f);
  const uint64_t c2 = pack2(0.24261091649532318f, 0.24261091649532318f), c3 = pack2(0.05517148599028587f, 0.05517148599028587f);
  const uint64_t t = add2(x, magic);
  const uint64_t n = add2(t, nmagic);
  const uint64_t f = fma2(n, neg1, x);
  uint64_t p = fma2(f, c3, c2);
  p = fma2(p, f, c1);
  p = fma2(p, f, c0);
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

}  // namespace

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (bf16 pairs, one row per lane, 8 columns per k-step of 16) stays in TMEM
IR_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// EMU8: of every 8 consecutive key pairs, EMU8 are exponentiated on the FMA pipe instead of the MUFU (0..4).
// ORDER > 0: the exponential sections of the two query tiles take turns (token passing through two mbarriers): one
// tile's MUFU section runs against the other tile's TMEM load / row max / MMAs. The token is handed over after ORDER
// of the 4 column chunks of the row (4 = strictly exclusive sections).
// TCSUM: the softmax denominator comes out of the tensor core: row 72 of every V^T tile in shared memory is a row of ones
// (rows 72..79 are the padding of the N = 80 MMA; the TMA box covers only the 72 real rows, the padding rows are written
// once per CTA), so O[:, 72] = sum_k P[:, k] accumulates in fp32 beside the output, is rescaled with it, and the softmax
// threads no longer add up their exponentials (64 packed adds per row and tile off the issue-bound exponential section).
template <int EMU8, int ORDER, bool TRACE = false, bool TCSUM = true>
__global__ void __launch_bounds__(NTHREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
               const __grid_constant__ CUtensorMap tmVT, const AttnTcDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                                 // [STAGES][K64 | K16 | VT0 | VT1]; K and V^T rings run separately
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + STAGES * STAGE_BYTES);
  uint64_t* k_full = bars;                    // [STAGES]
  uint64_t* k_empty = k_full + STAGES;        // [STAGES] both S = Q K^T of the tile have retired
  uint64_t* v_full = k_empty + STAGES;        // [STAGES]
  uint64_t* v_empty = v_full + STAGES;        // [STAGES] both P V of the tile have retired
  uint64_t* q_ready = v_empty + STAGES;       // [2] the query tile's rows are in TMEM
  uint64_t* s_full = q_ready + 2;             // [2] S(j) complete (which implies P V(j-1) complete: in-order pipe)
  uint64_t* p_full = s_full + 2;              // [2] P(j) is in TMEM
  uint64_t* o_full = p_full + 2;              // [2]
  uint64_t* order = o_full + 2;               // [2] order[t]: query tile t may run its exponential section
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(order + 2);
  static_assert((4 * STAGES + 10) * 8 + 4 <= 256, "barrier block");

  // warp index through a shuffle: provably warp-uniform, so the role branches below are convergent and the compiler
  // emits the uniform-datapath instructions (UTCHMMA / UTMALDG / UTCBAR) directly instead of an elect-one loop around
  // each of them (which costs ~80 cycles per MMA when the issuing code sits under a divergent `lane == 0`)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int head = blockIdx.y, b = blockIdx.z;
  const int n_tiles = (p.T + BKV - 1) / BKV;
  const bool trace_cta = TRACE && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  const bool tracing = trace_cta && lane == 0;
  auto stamp = [&](int role, int iter, int slot) {
    if (TRACE && tracing && iter < TRACE_ITERS) p.trace[(role * TRACE_ITERS + iter) * 8 + slot] = clock64();
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmK16);
    tma_prefetch_desc(&tmVT);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_ready[i], BQ);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], BQ);
      mbar_init(&o_full[i], 1);
      mbar_init(&order[i], BQ);
    }
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (TCSUM) {
    // rows 72..79 of the 2 * STAGES V^T atoms ([80 rows][64 keys] bf16, 128 B per row): row 72 = ones, the rest zeros.
    // Whole rows of one value, so the 128 B swizzle does not matter. 8 rows x 8 16-byte chunks per atom.
    for (int i = threadIdx.x; i < 2 * STAGES * 64; i += NTHREADS) {
      const int atom = i >> 6, r = (i >> 3) & 7, c = i & 7;
      uint8_t* base = sKV + (atom >> 1) * STAGE_BYTES + K64_BYTES + K16_BYTES + (atom & 1) * VT_ATOM_BYTES;
      const uint32_t w = r == 0 ? 0x3F803F80u : 0u;
      *reinterpret_cast<uint4*>(base + (HD + r) * 128 + c * 16) = make_uint4(w, w, w, w);
    }
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's (async-proxy) operand reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  // setmaxnreg sits at the top of each role branch (ptxas budgets registers per branch from it)
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    setmaxnreg_dec<REGS_CTRL>();
    const bool leader = elect_one_sync();   // the whole warp runs the loop (convergent), one lane issues
    {
      // K runs one tile ahead of V (S(j+1) is queued right behind P V(j)), so the two rings have their own barriers:
      // the K slot of tile j is free as soon as both S(j) retired, one iteration before the V slot.
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles + 1; ++j) {
        if (j < n_tiles) {
          mbar_wait(&k_empty[stage], phase ^ 1);
          uint8_t* st = sKV + stage * STAGE_BYTES;
          if (leader) {
            mbar_arrive_expect_tx(&k_full[stage], K64_BYTES + K16_BYTES);
            tma_load_4d(st, &tmK64, &k_full[stage], 0, j * BKV, head, b);
            tma_load_4d(st + K64_BYTES, &tmK16, &k_full[stage], 64, j * BKV, head, b);
          }
          __syncwarp();
        }
        if (j >= 1) {   // V(j-1)
          const int jv = j - 1;
          const int vs = jv % STAGES;
          const uint32_t vp = (uint32_t)((jv / STAGES) & 1);
          mbar_wait(&v_empty[vs], vp ^ 1);
          uint8_t* st = sKV + vs * STAGE_BYTES + K64_BYTES + K16_BYTES;
          if (leader) {
            mbar_arrive_expect_tx(&v_full[vs], 2 * (TCSUM ? HD * 128 : VT_ATOM_BYTES));
            tma_load_4d(st, &tmVT, &v_full[vs], jv * BKV, 0, head, b);
            tma_load_4d(st + VT_ATOM_BYTES, &tmVT, &v_full[vs], jv * BKV + 64, 0, head, b);
          }
          __syncwarp();
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ MMA issuer (warp 1: convergent, one lane issues)
    setmaxnreg_dec<REGS_CTRL>();
    if (warp == 1) {
      const bool leader = elect_one_sync();
      auto istamp = [&](int iter, int slot) {
        if (TRACE && trace_cta && leader && iter < TRACE_ITERS) p.trace[(2 * TRACE_ITERS + iter) * 8 + slot] = clock64();
      };
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, NV);
      auto issue_s = [&](int qt, int t) {   // S_qt(t) = Q_qt K(t)^T, A = Q from TMEM
        const int stage = t % STAGES;
        mbar_wait(&k_full[stage], (uint32_t)((t / STAGES) & 1));
        tc_fence_after();
        const uint32_t ka = smem_u32(sKV + stage * STAGE_BYTES);
        const uint64_t dk = make_smem_desc_sw128(ka);
        const uint32_t td = tmem_base + (qt == 0 ? TMEM_S0 : TMEM_S1);
        const uint32_t tq = tmem_base + (qt == 0 ? TMEM_Q0 : TMEM_Q1);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(td, tq + 8 * k, dk + (uint64_t)(2 * k), idesc_s, k != 0);
          umma_bf16_ts(td, tq + 32, make_smem_desc_sw32(ka + K64_BYTES), idesc_s, 1);
          umma_commit(&s_full[qt]);
          if (qt == 1) umma_commit(&k_empty[stage]);   // both S of tile t issued: the K slot is free once they retire
        }
        __syncwarp();
      };
      auto issue_pv = [&](int qt, int t) {   // O_qt += P_qt(t) V(t), A = P from TMEM (over S_qt's first 64 columns)
        const int stage = t % STAGES;
        if (qt == 0) mbar_wait(&v_full[stage], (uint32_t)((t / STAGES) & 1));
        mbar_wait(&p_full[qt], (uint32_t)(t & 1));
        istamp(t, 4 * qt + 1);
        tc_fence_after();
        const uint32_t va = smem_u32(sKV + stage * STAGE_BYTES + K64_BYTES + K16_BYTES);
        const uint32_t td = tmem_base + (qt == 0 ? TMEM_O0 : TMEM_O1);
        const uint32_t tp = tmem_base + (qt == 0 ? TMEM_S0 : TMEM_S1);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk) {
            const uint64_t dv = make_smem_desc_sw128(va + (kk >> 2) * VT_ATOM_BYTES) + (uint64_t)(2 * (kk & 3));
            umma_bf16_ts(td, tp + 8 * kk, dv, idesc_o, (t > 0 || kk != 0) ? 1u : 0u);
          }
          if (qt == 1) umma_commit(&v_empty[stage]);
          if (t + 1 == n_tiles) umma_commit(&o_full[qt]);
        }
        __syncwarp();
      };
#pragma unroll 1
      for (int qt = 0; qt < 2; ++qt) {
        mbar_wait(&q_ready[qt], 0);
        issue_s(qt, 0);
      }
#pragma unroll 1
      for (int j = 0; j < n_tiles; ++j) {
#pragma unroll 1
        for (int qt = 0; qt < 2; ++qt) {
          istamp(j, 4 * qt);
          issue_pv(qt, j);
          istamp(j, 4 * qt + 2);
          if (j + 1 < n_tiles) issue_s(qt, j + 1);   // in-order pipe: overwrites S/P only after P V(j) has read P
          istamp(j, 4 * qt + 3);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax: one thread per query row
    setmaxnreg_inc<REGS_SOFTMAX>();
    const int qt = (warp - 4) >> 2;
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_s = lane_base + (qt == 0 ? TMEM_S0 : TMEM_S1);
    const uint32_t t_o = lane_base + (qt == 0 ? TMEM_O0 : TMEM_O1);
    const int row = q0 + qt * BQ + r;
    {
      // this thread's query row -> TMEM as the A operand of S = Q K^T: bf16 pairs, d 72..79 (and rows beyond T) zero
      uint32_t qv[48];
#pragma unroll
      for (int i = 0; i < 48; ++i) qv[i] = 0u;
      if (row < p.T) {
        const uint4* src = reinterpret_cast<const uint4*>(p.q + (((long)b * p.H + head) * p.T + row) * HD);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const uint4 v = __ldg(src + i);
          qv[4 * i] = v.x;
          qv[4 * i + 1] = v.y;
          qv[4 * i + 2] = v.z;
          qv[4 * i + 3] = v.w;
        }
      }
      const uint32_t t_q = lane_base + (qt == 0 ? TMEM_Q0 : TMEM_Q1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t part[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) part[i] = qv[c * 16 + i];
        tmem_st_32x16(t_q + c * 16, part);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_ready[qt]);
    }
    const uint64_t scale2 = pack2(p.scale_log2e, p.scale_log2e);
    float m_used = -INFINITY;
    uint64_t l2 = pack2(0.f, 0.f);   // running row sum, two partial sums (rescaled together)
    for (int j = 0; j < n_tiles; ++j) {
      if (quarter == 0) stamp(qt, j, 0);
      mbar_wait(&s_full[qt], (uint32_t)(j & 1));
      if (quarter == 0) stamp(qt, j, 1);
      tc_fence_after();
      uint32_t s[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, s[c]);
      tmem_ld_wait();
      if (quarter == 0) stamp(qt, j, 2);
      const int kbase = j * BKV;
      if (kbase + BKV > p.T) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kbase + c * 32 + i >= p.T) s[c][i] = __float_as_uint(-INFINITY);
      }
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains of 3-input max
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          mxa[(i >> 1) & 3] = fmaxf(mxa[(i >> 1) & 3], fmaxf(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])));
      const float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
      const float m_new = fmaxf(m_used, mx * p.scale_log2e);
      if (quarter == 0) stamp(qt, j, 3);
      // lazy rescale: keep the stale max unless it grew by more than 8 (p stays below 2^8, exact in fp32 / fine in bf16)
      const bool grow = (m_new - m_used) > 8.0f;   // also true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float alpha = grow ? fast_exp2(m_used - m_new) : 1.0f;
        if (grow) m_used = m_new;
        const uint64_t alpha2 = pack2(alpha, alpha);
        if (!TCSUM) l2 = fma2(l2, alpha2, pack2(0.f, 0.f));
        if (j > 0) {
          // O_qt holds P V(0..j-1): complete, because S(j) (observed through s_full) was issued behind P V(j-1)
#pragma unroll
          for (int c = 0; c < NV / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(t_o + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x16(t_o + c * 16, o);
          }
        }
      }
      const uint64_t negm2 = pack2(-m_used, -m_used);
      // scores whose exponent would fall below 2^-120 are clamped there (only the polynomial path needs it)
      const float s_lo = (m_used - 120.0f) * p.inv_scale_log2e;
      uint64_t sum_a = pack2(0.f, 0.f), sum_b = pack2(0.f, 0.f);
      if (ORDER) mbar_wait(&order[qt], (uint32_t)((j & 1) ^ (qt == 0 ? 1 : 0)));   // tile 0's first turn is free
      if (quarter == 0) stamp(qt, j, 4);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];   // 32 keys of the P row as bf16 pairs -> 16 TMEM columns over S(j), which lives in registers now
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int pr = 0; pr < 4; ++pr) {
            const int pi = (g & 1) * 4 + pr;                 // pair index inside a 16-key window
            const bool emulate = ((pi * EMU8) & 7) < EMU8;   // spreads EMU8 emulated pairs over the window
            float s0 = __uint_as_float(s[c][g * 8 + 2 * pr]), s1 = __uint_as_float(s[c][g * 8 + 2 * pr + 1]);
            if (emulate) {
              s0 = fmaxf(s0, s_lo);
              s1 = fmaxf(s1, s_lo);
            }
            const uint64_t x = fma2(pack2(s0, s1), scale2, negm2);
            float e0, e1;
            if (emulate) {
              exp2_poly2(x, e0, e1);
            } else {
              float x0, x1;
              unpack2(x, x0, x1);
              e0 = fast_exp2(x0);
              e1 = fast_exp2(x1);
            }
            if (!TCSUM) {
              if (pr & 1)
                sum_b = add2(sum_b, pack2(e0, e1));
              else
                sum_a = add2(sum_a, pack2(e0, e1));
            }
            pk[g * 4 + pr] = pack_bf16x2(e0, e1);
          }
        }
        tmem_st_32x16(t_s + c * 16, pk);
        if (ORDER > 0 && c == ORDER - 1) mbar_arrive(&order[qt ^ 1]);   // hand the MUFU over
      }
      if (!TCSUM) l2 = add2(l2, add2(sum_a, sum_b));
      if (quarter == 0) stamp(qt, j, 5);
      tmem_st_wait();        // P (and a rescaled O) are in TMEM
      tc_fence_before();
      mbar_arrive(&p_full[qt]);
      if (quarter == 0) stamp(qt, j, 6);
    }
    float l_lo, l_hi;
    unpack2(l2, l_lo, l_hi);
    float l = l_lo + l_hi;
    // ---- epilogue: O / l -> bf16 -> out[(b*T + row)][head*72 + d]
    mbar_wait(&o_full[qt], 0);
    tc_fence_after();
    uint32_t o2[16];   // accumulator columns 64..79: d = 64..71, then (TCSUM) the row sum in column 72
    tmem_ld_32x16(t_o + 64, o2);
    tmem_ld_wait();
    if (TCSUM) l = __uint_as_float(o2[8]);
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
    if (row < p.T) {
      uint4 u = make_uint4(pack_bf16x2(__uint_as_float(o2[0]) * inv, __uint_as_float(o2[1]) * inv),
                           pack_bf16x2(__uint_as_float(o2[2]) * inv, __uint_as_float(o2[3]) * inv),
                           pack_bf16x2(__uint_as_float(o2[4]) * inv, __uint_as_float(o2[5]) * inv),
                           pack_bf16x2(__uint_as_float(o2[6]) * inv, __uint_as_float(o2[7]) * inv));
      *reinterpret_cast<uint4*>(og + 64) = u;   // columns 64..71; accumulator columns 72..79 are the row sum / padding
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

static long long* g_attn_trace = nullptr;
void attention_tc_set_trace(long long* device_buf) { g_attn_trace = device_buf; }
int attention_tc_trace_len() { return 4 * TRACE_ITERS * 8; }

int attention_tc_launch(const AttnTcArgs& a, cudaStream_t stream) {
  IR_REQUIRE(a.head_dim == HD, "attention_tc: head_dim %d unsupported (kernel is specialised for %d)", a.head_dim, HD);
  IR_REQUIRE(a.q && a.k && a.vt && a.out && a.B > 0 && a.H > 0 && a.T > 0, "attention_tc: bad arguments");
  IR_REQUIRE(a.Tp % 8 == 0 && a.Tp >= a.T && a.ldo % 8 == 0, "attention_tc: Tp must be a multiple of 8 and >= T");
  static const bool tcsum = [] {
    const char* e = debug_env("IR_ATTN_TCSUM");   // A/B switch of debug builds: 0 = row sums on the CUDA cores
    return !(e && e[0] == '0');
  }();
  CUtensorMap mk64, mk16, mvt;
  {
    const uint64_t dims[4] = {(uint64_t)HD, (uint64_t)a.T, (uint64_t)a.H, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)HD * 2, (uint64_t)a.T * HD * 2, (uint64_t)a.H * a.T * HD * 2};
    const uint32_t box64[4] = {64, BKV, 1, 1};
    const uint32_t box16[4] = {16, BKV, 1, 1};
    IR_TRY(make_tensor_map(&mk64, a.k, 4, dims, strides, box64, 128));
    IR_TRY(make_tensor_map(&mk16, a.k, 4, dims, strides, box16, 32));
  }
  {
    const uint64_t dims[4] = {(uint64_t)a.T, (uint64_t)HD, (uint64_t)a.H, (uint64_t)a.B};
    const uint64_t strides[3] = {(uint64_t)a.Tp * 2, (uint64_t)HD * a.Tp * 2, (uint64_t)a.H * HD * a.Tp * 2};
    // TCSUM kernels: the box covers the 72 real rows; rows 72..79 of the shared-memory atoms hold the ones row / zeros
    const uint32_t box[4] = {64, (uint32_t)(tcsum ? HD : NV), 1, 1};
    IR_TRY(make_tensor_map(&mvt, a.vt, 4, dims, strides, box, 128));
  }
  AttnTcDev p;
  p.q = a.q;
  p.out = a.out;
  p.ldo = a.ldo;
  p.T = a.T;
  p.H = a.H;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  p.inv_scale_log2e = 1.0f / p.scale_log2e;
  p.trace = g_attn_trace;
  dim3 grid((a.T + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
  // share of the exponentials evaluated on the FMA pipe: EMU8 of every 8 key pairs. With the row sums on the tensor core
  // (TCSUM) 3 of 8 balances the MUFU and issue-slot budgets of the softmax warps (B8 T4096: 673 us against 699 / 701 us at
  // 2 / 4 of 8; 678 us for the old kernel with register row sums at its own optimum of 2); IR_ATTN_EMU overrides it for
  // measurements (tools/gpu_attn_probe.py, tools/gpu_attn_sweep.sh).
  static const int emu = [] {
    const char* e = debug_env("IR_ATTN_EMU");
    const int v = e ? atoi(e) : 3;
    return v < 0 ? 0 : (v > 4 ? 4 : v);
  }();
  auto launch = [&](auto kernel) -> int {
    IR_TRY(ensure_smem_optin((const void*)kernel, SMEM_BYTES));
    const bool prof = prof_enabled();
    if (prof) prof_before(stream);
    IR_CUDA_CHECK(launch_pdl(kernel, grid, dim3(NTHREADS), SMEM_BYTES, stream, mk64, mk16, mvt, p));
    if (prof) prof_after(stream, PROF_ATTN, 4.0 * a.B * a.H * (double)a.T * a.T * HD, a.B * a.H, a.T, HD);
    return IR_OK;
  };
  static const int order = [] {
    const char* e = debug_env("IR_ATTN_ORDER");
    const int v = e ? atoi(e) : 3;
    return v < 0 ? 0 : (v > 4 ? 4 : v);
  }();
  if (!tcsum) {
    IR_TRY(launch(attn_tc_kernel<2, 3, false, false>));
    IR_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return IR_OK;
  }
  if (g_attn_trace) {
    switch (order) {
      case 0: return launch(attn_tc_kernel<2, 0, true>);
      case 3: return launch(attn_tc_kernel<3, 3, true>);
      default: return launch(attn_tc_kernel<2, 4, true>);
    }
  }
  const int key = order * 10 + emu;
  switch (key) {
    case 0: IR_TRY(launch(attn_tc_kernel<0, 0>)); break;
    case 2: IR_TRY(launch(attn_tc_kernel<2, 0>)); break;
    case 10: IR_TRY(launch(attn_tc_kernel<0, 1>)); break;
    case 12: IR_TRY(launch(attn_tc_kernel<2, 1>)); break;
    case 20: IR_TRY(launch(attn_tc_kernel<0, 2>)); break;
    case 22: IR_TRY(launch(attn_tc_kernel<2, 2>)); break;
    case 23: IR_TRY(launch(attn_tc_kernel<3, 2>)); break;
    case 24: IR_TRY(launch(attn_tc_kernel<4, 2>)); break;
    case 30: IR_TRY(launch(attn_tc_kernel<0, 3>)); break;
    case 32: IR_TRY(launch(attn_tc_kernel<2, 3>)); break;
    case 33: IR_TRY(launch(attn_tc_kernel<3, 3>)); break;
    case 34: IR_TRY(launch(attn_tc_kernel<4, 3>)); break;
    case 40: IR_TRY(launch(attn_tc_kernel<0, 4>)); break;
    case 42: IR_TRY(launch(attn_tc_kernel<2, 4>)); break;
    case 44: IR_TRY(launch(attn_tc_kernel<4, 4>)); break;
    default:
      IR_REQUIRE(false, "attention_tc: no instantiation for IR_ATTN_ORDER=%d IR_ATTN_EMU=%d", order, emu);
  }
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

}  // namespace ir
