// tcgen05 / TMEM cross-attention to the packed caption (MultiHeadCrossAttention.forward,
// diffusion/model/nets/PixArt_blocks.py:43-58: q = q_linear(x) (1, B*T, 16, 72), kv = kv_linear(y) (1, sum L, 2, 16, 72),
// xformers memory_efficient_attention with BlockDiagonalMask.from_seqlens([T]*B, y_lens): sample b's queries attend to
// that sample's L_b valid caption tokens only, scale 72^-1/2).
//
// The problem is tiny in FLOPs (4*T*L*72 per head, L <= 300) and therefore latency- and issue-bound; the design keeps
// the whole caption of one (sample, head) resident and makes the per-query-tile chain as short as possible:
//   * one CTA (4 warps = the 4 TMEM lane quarters, one thread per query row) owns one (sample, head) and walks QT
//     consecutive 128-row query tiles; QT is chosen so that the grid is a whole number of resident waves;
//   * K_h [L][72] arrives once per CTA by TMA straight out of the cached kv_linear output [sum L][2304] (a 64-column
//     SWIZZLE_128B box and a 16-column SWIZZLE_32B box per row block; what the 16-column box reads beyond d = 72 -- the
//     next head's first columns -- meets zeros in Q and contributes nothing); V_h^T [72][L] arrives by TMA from a
//     transposed copy of the V half, written once per caption next to the K/V cache (xattention_transpose_v), as
//     K-major SWIZZLE_128B atoms of 64 keys (rows d = 72..79 are zero-filled by the tensor-map bounds);
//   * both A operands live in TMEM (tcgen05.mma TS form): Q is written by the threads (bf16 pairs, d 72..79 zero), S = Q K^T
//     lands in TMEM, the softmax threads read their S row (L <= 128: once, into registers; longer captions: a max pass and
//     an exp pass in 16-column pieces), write P as bf16 pairs over the S columns they have consumed, and O = P V
//     accumulates over the dead upper half of S; the next tile's Q row is fetched from global memory while the current
//     tile is in flight;
//   * TMEM columns: S [0, L), P [0, L/2), O [~L/2, +80), Q [L, L+40) (Q is dead once S is complete, so O may overlap it):
//     128 columns and ~35 KB of shared memory per CTA for L <= 80, so three to four CTAs share an SM and interleave
//     their chains (TMA / MMA of one under the softmax / stores of the others).
// Keys beyond L_b inside the padded tile (the next sample's tokens, or zero fill beyond sum L) are masked: whole 16-key
// pieces beyond L_b are skipped (P = 0), only the one straddling piece pays per-element compares. P there is exactly 0
// and everything V^T can hold there is finite, so ragged captions need no separate code path.
#include "attention.cuh"

namespace ir {

namespace {

constexpr int XHD = 72;            // head dim
constexpr int XNV = 80;            // head dim padded to a multiple of 16 (UMMA N of P*V, K of Q*K^T)
constexpr int XBQ = 128;           // query rows per tile (UMMA M)
constexpr int XVT_ATOM = XNV * 64 * 2;   // 10240 B: V^T atom [80 rows (d)][64 keys], SWIZZLE_128B K-major
constexpr int XNTHREADS = 128;

struct XAttnDev {
  long ldkv;
  const int* kv_off;   // device [B]: first packed caption row of each sample
  const int* kv_len;   // device [B]: valid caption tokens of each sample
  int T, H;
  int Lp;           // padded key count (multiple of 16, >= every kv_len, <= 384)
  int box_rows;     // rows per K TMA box (Lp, or Lp/2 when Lp > 256)
  int QT;           // query tiles per CTA
  int tmem_cols;    // power of two >= max(o_col + 80, Lp + 40)
  int o_col;        // first TMEM column of the O accumulator (multiple of 16, >= Lp / 2)
  float scale_log2e;
};

IR_DEVINL float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

IR_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

IR_DEVINL void tmem_st_32x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

IR_DEVINL int al1024(int v) { return (v + 1023) & ~1023; }

// TMA store of a shared-memory tile (bulk async-group completion)
IR_DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
IR_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
IR_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
IR_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
IR_DEVINL uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
IR_DEVINL void sts_u4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
constexpr int XQ_TILE = XBQ * XHD * 2;   // 18432 B: one query / output tile [128 rows][144 B], dense (no swizzle)

// NP > 0: Lp == 16 * NP <= 128 known at compile time -- the S row is read from TMEM once and lives in registers (max,
// then exponentials). NP == 0: any Lp <= 384, two passes over TMEM in 16-column pieces.
template <int NP>
__global__ void __launch_bounds__(XNTHREADS, (NP > 0 && NP <= 5) ? 3 : 2)
xattn_tc_kernel(const __grid_constant__ CUtensorMap tmK64, const __grid_constant__ CUtensorMap tmK16,
                const __grid_constant__ CUtensorMap tmVT, const __grid_constant__ CUtensorMap tmQ,
                const __grid_constant__ CUtensorMap tmO, const XAttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Lp = NP > 0 ? NP * 16 : p.Lp;
  const int n_atoms = (Lp + 63) >> 6;
  uint8_t* sK128 = smem;                               // [Lp rows][64 d]  128 B rows, SWIZZLE_128B (TMA)
  uint8_t* sK32 = sK128 + al1024(Lp * 128);            // [Lp rows][16 d]   32 B rows, SWIZZLE_32B  (TMA)
  uint8_t* sVT = sK32 + al1024(Lp * 32);               // [n_atoms][80 d][64 keys] SWIZZLE_128B (TMA)
  uint8_t* sQ = sVT + n_atoms * XVT_ATOM;              // [2][128 rows][144 B]: Q tile in (TMA), O tile out (TMA store)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + 2 * XQ_TILE);
  uint64_t* k_bar = bars;        // K and V^T landed
  uint64_t* s_bar = bars + 1;    // S = Q K^T complete
  uint64_t* o_bar = bars + 2;    // O = P V complete
  uint64_t* q_bar = bars + 3;    // [2] Q tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const int head = blockIdx.y, b = blockIdx.z;
  const int n_qtiles = (p.T + XBQ - 1) / XBQ;
  const int tile0 = blockIdx.x * p.QT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK64);
    tma_prefetch_desc(&tmK16);
    tma_prefetch_desc(&tmVT);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(k_bar, 1);
    mbar_init(s_bar, 1);
    mbar_init(o_bar, 1);
    mbar_init(&q_bar[0], 1);
    mbar_init(&q_bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  // TMA wants the box start 16-byte aligned in the innermost dimension (keys, for V^T): the key window starts at the packed
  // row rounded down to a multiple of 8; the sample's keys are columns [lo, hi) of the window (the host sizes Lp for
  // max_b (kv_off[b] % 8 + kv_len[b]))
  const int koff_raw = p.kv_off[b];
  const int koff = koff_raw & ~7;
  const int lo = koff_raw & 7;
  const int hi = min(lo + p.kv_len[b], Lp);
  const bool leader_warp = warp == 0;
  if (leader_warp) {
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(k_bar, (uint32_t)(Lp * (128 + 32) + n_atoms * XVT_ATOM));
      for (int r0 = 0; r0 < Lp; r0 += p.box_rows) {
        tma_load_2d(sK128 + r0 * 128, &tmK64, k_bar, head * XHD, koff + r0);
        tma_load_2d(sK32 + r0 * 32, &tmK16, k_bar, head * XHD + 64, koff + r0);
      }
      for (int a = 0; a < n_atoms; ++a) tma_load_3d(sVT + a * XVT_ATOM, &tmVT, k_bar, koff + a * 64, 0, head);
      if (tile0 < n_qtiles) {   // first query tile (rows beyond T are zero-filled by the tensor-map bounds)
        mbar_arrive_expect_tx(&q_bar[0], XQ_TILE);
        tma_load_3d(sQ, &tmQ, &q_bar[0], head * XHD, tile0 * XBQ, b);
      }
    }
    __syncwarp();
  }

  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const uint32_t S_COL = 0, O_COL = (uint32_t)p.o_col, Q_COL = (uint32_t)Lp;
  const uint32_t t_s = lane_base + S_COL, t_o = lane_base + O_COL, t_q = lane_base + Q_COL;
  const int n16 = Lp >> 4;
  uint32_t ph = 0;
  const int r_tile = warp * 32 + lane;   // this thread's row inside the tile == its TMEM lane
  for (int it = 0; it < p.QT; ++it) {
    const int qt = tile0 + it;
    if (qt >= n_qtiles) break;   // uniform over the CTA
    const int bq = it & 1;
    uint8_t* sQb = sQ + bq * XQ_TILE;
    if (leader_warp) {
      // prefetch the next query tile into the other buffer, once the TMA store of the output tile staged there has read it
      if (it + 1 < p.QT && qt + 1 < n_qtiles && elect_one_sync()) {
        tma_store_wait_read();
        mbar_arrive_expect_tx(&q_bar[bq ^ 1], XQ_TILE);
        tma_load_3d(sQ + (bq ^ 1) * XQ_TILE, &tmQ, &q_bar[bq ^ 1], head * XHD, (qt + 1) * XBQ, b);
      }
      __syncwarp();
    }
    mbar_wait(&q_bar[bq], (uint32_t)((it >> 1) & 1));
    uint32_t qv[36];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const uint4 v = lds_u4(smem_u32(sQb + r_tile * (XHD * 2) + i * 16));
      qv[4 * i] = v.x;
      qv[4 * i + 1] = v.y;
      qv[4 * i + 2] = v.z;
      qv[4 * i + 3] = v.w;
    }
    {
      // Q row -> TMEM as the A operand of S = Q K^T: 36 bf16 pairs + 4 zero pairs (d 72..79); columns 40..47 unused
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t part[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) part[i] = qv[c * 16 + i];
        tmem_st_32x16(t_q + c * 16, part);
      }
      uint32_t tail[8] = {qv[32], qv[33], qv[34], qv[35], 0u, 0u, 0u, 0u};
      tmem_st_32x8(t_q + 32, tail);
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    if (leader_warp) {
      if (it == 0) mbar_wait(k_bar, 0);
      tc_fence_after();
      if (elect_one_sync()) {
        for (int c0 = 0; c0 < Lp; c0 += 128) {   // S in chunks of <= 128 keys
          const int nk = min(128, Lp - c0);
          const uint32_t idesc = make_idesc_bf16(XBQ, nk);
          const uint64_t dk = make_smem_desc_sw128(smem_u32(sK128 + c0 * 128));
          const uint32_t td = tmem_base + S_COL + (uint32_t)c0;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ts(td, tmem_base + Q_COL + 8 * k, dk + (uint64_t)(2 * k), idesc, k != 0);
          umma_ts(td, tmem_base + Q_COL + 32, make_smem_desc_sw32(smem_u32(sK32 + c0 * 32)), idesc, 1);
        }
        umma_commit(s_bar);
      }
      __syncwarp();
    }
    mbar_wait(s_bar, ph);
    tc_fence_after();
    float l0 = 0.f, l1 = 0.f;
    if (NP > 0) {
      // single pass: the whole S row in registers
      uint32_t s[NP > 0 ? NP : 1][16];
#pragma unroll
      for (int i = 0; i < NP; ++i) tmem_ld_32x16(t_s + i * 16, s[i]);
      tmem_ld_wait();
      // lo / hi are uniform over the CTA: a 16-key piece is either entirely valid (no masking), entirely outside (skipped,
      // P = 0) or straddles an end of [lo, hi) (per-element compares; at most two such pieces)
      float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        if (i * 16 >= lo && i * 16 + 16 <= hi) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mxa[j & 3] = fmaxf(mxa[j & 3], __uint_as_float(s[i][j]));
        } else if (i * 16 + 16 > lo && i * 16 < hi) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (i * 16 + j >= lo && i * 16 + j < hi) mxa[j & 3] = fmaxf(mxa[j & 3], __uint_as_float(s[i][j]));
        }
      }
      const float msc = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])) * p.scale_log2e;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        uint32_t pk[8];
        if (i * 16 >= lo && i * 16 + 16 <= hi) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float e0 = ex2(fmaf(__uint_as_float(s[i][2 * j]), p.scale_log2e, -msc));
            const float e1 = ex2(fmaf(__uint_as_float(s[i][2 * j + 1]), p.scale_log2e, -msc));
            l0 += e0;
            l1 += e1;
            pk[j] = pack_bf16x2(e0, e1);
          }
        } else if (i * 16 + 16 > lo && i * 16 < hi) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k0 = i * 16 + 2 * j;
            const float e0 = (k0 >= lo && k0 < hi) ? ex2(fmaf(__uint_as_float(s[i][2 * j]), p.scale_log2e, -msc)) : 0.f;
            const float e1 = (k0 + 1 >= lo && k0 + 1 < hi) ? ex2(fmaf(__uint_as_float(s[i][2 * j + 1]), p.scale_log2e, -msc)) : 0.f;
            l0 += e0;
            l1 += e1;
            pk[j] = pack_bf16x2(e0, e1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = 0u;
        }
        tmem_st_32x8(t_s + i * 8, pk);
      }
    } else {
      // pass 1: row max over the valid keys
      float mx = -INFINITY;
      for (int i = 0; i < n16; ++i) {
        uint32_t s[16];
        tmem_ld_32x16(t_s + i * 16, s);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (i * 16 + j >= lo && i * 16 + j < hi) mx = fmaxf(mx, __uint_as_float(s[j]));
      }
      const float msc = mx * p.scale_log2e;
      // pass 2: p = 2^(s*scale*log2e - m), row sum, bf16 pairs written over the S columns already consumed
      for (int i = 0; i < n16; ++i) {
        uint32_t s[16];
        tmem_ld_32x16(t_s + i * 16, s);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k0 = i * 16 + 2 * j;
          const float e0 = (k0 >= lo && k0 < hi) ? ex2(fmaf(__uint_as_float(s[2 * j]), p.scale_log2e, -msc)) : 0.f;
          const float e1 = (k0 + 1 >= lo && k0 + 1 < hi) ? ex2(fmaf(__uint_as_float(s[2 * j + 1]), p.scale_log2e, -msc)) : 0.f;
          l0 += e0;
          l1 += e1;
          pk[j] = pack_bf16x2(e0, e1);
        }
        tmem_st_32x8(t_s + i * 8, pk);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (leader_warp) {
      tc_fence_after();
      if (elect_one_sync()) {
        constexpr uint32_t idesc_o = make_idesc_bf16(XBQ, XNV);
        for (int kk = 0; kk < n16; ++kk) {
          const uint64_t dv = make_smem_desc_sw128(smem_u32(sVT + (kk >> 2) * XVT_ATOM)) + (uint64_t)(2 * (kk & 3));
          umma_ts(tmem_base + O_COL, tmem_base + S_COL + 8 * kk, dv, idesc_o, kk != 0);
        }
        umma_commit(o_bar);
      }
      __syncwarp();
    }
    mbar_wait(o_bar, ph);
    tc_fence_after();
    const float l = l0 + l1;
    const float inv = l > 0.f ? 1.0f / l : 0.f;
    // O / l -> bf16 -> this thread's row of the tile buffer (the Q tile in it was consumed before the first barrier of this
    // iteration), then one TMA store of the [128][72] tile: rows beyond T are clipped by the tensor-map bounds
    const uint32_t orow = smem_u32(sQb + r_tile * (XHD * 2));
    uint32_t o[32];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tmem_ld_32x32(t_o + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g)
        sts_u4(orow + c * 64 + g * 16,
               make_uint4(pack_bf16x2(__uint_as_float(o[g * 8]) * inv, __uint_as_float(o[g * 8 + 1]) * inv),
                          pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv),
                          pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv),
                          pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv)));
    }
    uint32_t o2[16];
    tmem_ld_32x16(t_o + 64, o2);
    tmem_ld_wait();
    sts_u4(orow + 128, make_uint4(pack_bf16x2(__uint_as_float(o2[0]) * inv, __uint_as_float(o2[1]) * inv),
                                  pack_bf16x2(__uint_as_float(o2[2]) * inv, __uint_as_float(o2[3]) * inv),
                                  pack_bf16x2(__uint_as_float(o2[4]) * inv, __uint_as_float(o2[5]) * inv),
                                  pack_bf16x2(__uint_as_float(o2[6]) * inv, __uint_as_float(o2[7]) * inv)));   // d 64..71; accumulator columns 72..79 are padding
    fence_proxy_async();   // generic-proxy writes of the tile -> async-proxy (TMA) read
    tc_fence_before();
    __syncthreads();
    if (leader_warp) {
      if (elect_one_sync()) {
        tma_store_3d(&tmO, sQb, head * XHD, qt * XBQ, b);
        tma_store_commit();
      }
      __syncwarp();
    }
    ph ^= 1;
  }

  if (leader_warp) {
    if (elect_one_sync()) tma_store_wait_all();   // the output tiles are in global memory before the CTA retires
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

// V half of one or more blocks' kv_linear output [nblk][sumL][ldkv] -> vt [nblk][H][72][sumLp] (keys contiguous): the
// K-major B operand of O = P V, fetched by TMA in 64-key atoms. Runs once per caption (the K/V are cached per run).
__global__ void xattn_transpose_v_kernel(const bf16* __restrict__ kv, bf16* __restrict__ vt, int nblk, int H, int sumL,
                                         int sumLp, long ldkv) {
  const long total = (long)nblk * H * XHD * sumL;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int tok = (int)(i % sumL);
    long r = i / sumL;
    const int d = (int)(r % XHD);
    r /= XHD;
    const int h = (int)(r % H);
    const int blk = (int)(r / H);
    vt[(((long)blk * H + h) * XHD + d) * sumLp + tok] = kv[((long)blk * sumL + tok) * ldkv + (long)H * XHD + h * XHD + d];
  }
}

long xattention_vt_elems(int H, int sumL) { return (long)H * XHD * ((sumL + 7) / 8 * 8); }

int xattention_transpose_v(const bf16* kv, bf16* vt, int nblk, int H, int sumL, long ldkv, cudaStream_t stream) {
  IR_REQUIRE(kv && vt && nblk > 0 && H > 0 && sumL > 0, "xattention_transpose_v: bad arguments");
  const long total = (long)nblk * H * XHD * sumL;
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  xattn_transpose_v_kernel<<<grid, 256, 0, stream>>>(kv, vt, nblk, H, sumL, (sumL + 7) / 8 * 8, ldkv);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

int xattention_tc_launch(const XAttnTcArgs& a, cudaStream_t stream) {
  IR_REQUIRE(a.head_dim == XHD, "xattention_tc: head_dim %d unsupported (kernel is specialised for %d)", a.head_dim, XHD);
  IR_REQUIRE(a.q && a.kv && a.vt && a.out && a.kv_off && a.kv_len, "xattention_tc: null pointer");
  IR_REQUIRE(a.B > 0 && a.H > 0 && a.T > 0 && a.sumL > 0, "xattention_tc: bad shape");
  IR_REQUIRE(a.max_len > 0 && a.max_len <= 384, "xattention_tc: %d caption tokens per sample unsupported (1..384)", a.max_len);
  IR_REQUIRE(a.ldq % 8 == 0 && a.ldkv % 8 == 0 && a.ldo % 8 == 0 && a.ldkv >= 2L * a.H * XHD, "xattention_tc: strides must keep 16 B rows");
  XAttnDev p;
  p.ldkv = a.ldkv;
  p.kv_off = a.kv_off;
  p.kv_len = a.kv_len;
  p.T = a.T;
  p.H = a.H;
  p.Lp = (a.max_len + 15) / 16 * 16;
  p.box_rows = p.Lp <= 256 ? p.Lp : p.Lp / 2;
  p.o_col = (p.Lp / 2 + 15) / 16 * 16;
  {
    const int need = (p.o_col + XNV > p.Lp + 40) ? p.o_col + XNV : p.Lp + 40;
    p.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
  }
  p.scale_log2e = a.scale * 1.4426950408889634f;
  CUtensorMap mk64, mk16, mvt, mq, mo;
  {
    // q / out tiles [128 rows][72] of one head, dense in shared memory; per-sample row bounds (T) so that a partial last
    // tile neither reads nor writes the next sample's rows
    const uint64_t dims[3] = {(uint64_t)a.H * XHD, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t sq[2] = {(uint64_t)a.ldq * 2, (uint64_t)a.T * a.ldq * 2};
    const uint64_t so[2] = {(uint64_t)a.ldo * 2, (uint64_t)a.T * a.ldo * 2};
    const uint32_t box[3] = {(uint32_t)XHD, (uint32_t)XBQ, 1};
    IR_TRY(make_tensor_map(&mq, a.q, 3, dims, sq, box, 0));
    IR_TRY(make_tensor_map(&mo, a.out, 3, dims, so, box, 0));
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.H * XHD, (uint64_t)a.sumL};
    const uint64_t strides[1] = {(uint64_t)a.ldkv * 2};
    const uint32_t box64[2] = {64, (uint32_t)p.box_rows};
    const uint32_t box16[2] = {16, (uint32_t)p.box_rows};
    IR_TRY(make_tensor_map(&mk64, a.kv, 2, dims, strides, box64, 128));
    IR_TRY(make_tensor_map(&mk16, a.kv, 2, dims, strides, box16, 32));
  }
  {
    const uint64_t sumLp = (uint64_t)((a.sumL + 7) / 8 * 8);
    const uint64_t dims[3] = {(uint64_t)a.sumL, (uint64_t)XHD, (uint64_t)a.H};
    const uint64_t strides[2] = {sumLp * 2, (uint64_t)XHD * sumLp * 2};
    const uint32_t box[3] = {64, (uint32_t)XNV, 1};
    IR_TRY(make_tensor_map(&mvt, a.vt, 3, dims, strides, box, 128));
  }
  const int n_atoms = (p.Lp + 63) / 64;
  const int smem = 1024 + ((p.Lp * 128 + 1023) & ~1023) + ((p.Lp * 32 + 1023) & ~1023) + n_atoms * XVT_ATOM + 2 * XQ_TILE + 64;
  constexpr int SMEM_MAX = 1024 + 384 * 128 + 384 * 32 + 6 * XVT_ATOM + 2 * XQ_TILE + 64;   // Lp = 384
  const int n_qtiles = (a.T + XBQ - 1) / XBQ;
  dim3 grid;
  auto launch = [&](auto kern) -> int {
    IR_TRY(ensure_smem_optin((const void*)kern, SMEM_MAX));
    // resident CTAs per SM: registers / shared memory (occupancy query) and TMEM columns
    static int occ_cache[32] = {};   // per instantiation (static of this generic lambda), indexed by Lp / 16
    int& occ_c = occ_cache[(p.Lp / 16) & 31];
    if (occ_c == 0) {
      // resident CTAs per SM from the kernel's own attributes (64 K registers allocated per warp in units of 256, 228 KB of
      // shared memory with 1 KB reserved per CTA); cudaOccupancyMaxActiveBlocksPerMultiprocessor answered 1 here on B200
      cudaFuncAttributes fa;
      IR_CUDA_CHECK(cudaFuncGetAttributes(&fa, kern));
      const int regs_cta = ((fa.numRegs * 32 + 255) / 256 * 256) * (XNTHREADS / 32);
      const int by_regs = 65536 / (regs_cta > 0 ? regs_cta : 1);
      const int by_smem = (228 * 1024) / (smem + 1024);
      int o = by_regs < by_smem ? by_regs : by_smem;
      occ_c = o < 1 ? 1 : (o > 16 ? 16 : o);
    }
    int occ = occ_c;
    const int by_tmem = 512 / p.tmem_cols;
    occ = occ < 1 ? 1 : (occ > by_tmem ? by_tmem : occ);
    const long slots = (long)device_num_sms() * occ;
    // query tiles per CTA: minimise (waves * (QT tiles + one tile-time of per-CTA set-up)); ties go to the larger QT
    int best_qt = 1;
    double best = 1e30;
    for (int qt = 1; qt <= 8 && qt <= n_qtiles; ++qt) {
      const long ctas = (long)((n_qtiles + qt - 1) / qt) * a.H * a.B;
      const long waves = (ctas + slots - 1) / slots;
      const double cost = (double)waves * (qt + 1.0);
      if (cost <= best) {
        best = cost;
        best_qt = qt;
      }
    }
    p.QT = best_qt;
    grid = dim3((n_qtiles + best_qt - 1) / best_qt, a.H, a.B);
    IR_CUDA_CHECK(launch_pdl(kern, grid, dim3(XNTHREADS), (size_t)smem, stream, mk64, mk16, mvt, mq, mo, p));
    return IR_OK;
  };
  const bool prof = prof_enabled();
  if (prof) prof_before(stream);
  switch (p.Lp <= 128 ? p.Lp / 16 : 0) {
    case 1: IR_TRY(launch(xattn_tc_kernel<1>)); break;
    case 2: IR_TRY(launch(xattn_tc_kernel<2>)); break;
    case 3: IR_TRY(launch(xattn_tc_kernel<3>)); break;
    case 4: IR_TRY(launch(xattn_tc_kernel<4>)); break;
    case 5: IR_TRY(launch(xattn_tc_kernel<5>)); break;
    case 6: IR_TRY(launch(xattn_tc_kernel<6>)); break;
    case 7: IR_TRY(launch(xattn_tc_kernel<7>)); break;
    case 8: IR_TRY(launch(xattn_tc_kernel<8>)); break;
    default: IR_TRY(launch(xattn_tc_kernel<0>)); break;
  }
  if (prof) prof_after(stream, PROF_XATTN, 4.0 * a.H * (double)a.T * XHD * (double)a.kv_total, a.B * a.H, a.T, a.max_len);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

}  // namespace ir
