// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA (128B-swizzled tiles) -> shared memory ring ->
// tcgen05.mma with fp32 accumulators in TMEM (double-buffered) -> tcgen05.ld epilogue with fused
// bias / GELU-tanh / gate*(.)+fp32 residual. The same kernel runs the VAE decoder's 3x3 convolutions as an
// implicit GEMM: the A tile of one (tap, channel-chunk) K-block is a 4-D TMA box over the NHWC activation
// whose out-of-bounds elements (the zero padding) are filled by the TMA unit.
//
// Replaces, on the reference side, every nn.Linear of diffusion/model/nets/PixArt_blocks.py:47-55,130,156,
// PixArtMS.py:66-67 (timm Mlp), pixart_controlnet.py:31-36 and the Conv2d 3x3 / 1x1 layers of
// ldm/modules/diffusionmodules/model.py:57-61,103-129,160-179.
#include "gemm.cuh"

#include <cstdio>
#include <cstdlib>
#include <mutex>

namespace ir {

static constexpr int BM = 128;       // rows per CTA tile == UMMA M == TMEM lanes
static constexpr int BK = 64;        // bf16 elements per K-block == one 128 B swizzle row
static constexpr int UMMA_K = 16;    // K per tcgen05.mma for 16-bit inputs
static constexpr int CONV_BW = 16;   // conv tile: 16 x 8 output pixels
static constexpr int CONV_BH = 8;
static constexpr int STG_LD = 36;    // staging row stride in floats (32 + 4 pad: conflict-free v4 stores)
static constexpr int NUM_THREADS_MAX = 384;

template <int BN, int CG, bool CONV, int EPI = EPI_BF16>
struct GemmCfg {
  // CG = 1: one CTA per 128 x BN tile. CG = 2: a CTA pair (cta_group::2) per 256 x BN tile; each CTA stages its own
  // 128 rows of A and BN/2 rows of B, which halves the L2 -> shared-memory traffic of the B operand per FLOP.
  // CONV: one pipeline stage = (channel chunk, kx): a (16 x 10)-pixel halo box of the activation (160 rows of 128 B)
  // serves the three ky taps through descriptor row offsets (+16 rows = +2048 B each, swizzle-phase preserving), and
  // three weight tiles (one per ky). The activation is fetched 3x per channel chunk instead of 9x.
  static constexpr int A_ROWS = CONV ? CONV_BW * (CONV_BH + 2) : BM;
  static constexpr int A_BYTES = A_ROWS * BK * 2;
  static constexpr int B_ROWS = BN / CG;
  static constexpr int B_TAP_BYTES = B_ROWS * BK * 2;
  static constexpr int B_BYTES = (CONV ? 3 : 1) * B_TAP_BYTES;
  // epilogue warps: 8 for plain GEMMs (two per TMEM lane quarter, interleaved 32-column chunks): the residual / scatter
  // epilogues are latency-bound, more warps = more loads in flight. Convs: 8 where the second staging area does not cost
  // a pipeline stage (per-CTA weight tiles of <= 64 rows: the 256x128 pair tiles and 128x64), 4 otherwise. The
  // full-resolution N = 128 convs are epilogue-bound with 4 warps: 367 us (K = 1152) vs 526 us (K = 2304) at 1024^2 is a
  // marginal MMA cost of 158 us per 309 GFLOP on top of ~209 us that does not scale with K (~6500 cycles per tile).
  static constexpr int EW = CONV ? ((BN / CG <= 64) ? 8 : 4) : 8;
  static constexpr int NUM_THREADS = 128 + 32 * EW;
  // staging per epilogue warp: 4608 B for the transposing epilogues (32 x 36 floats); the row-owner epilogues of the linear
  // GEMMs stage two 4 KB TMA boxes instead: bf16 outputs ping-pong between two [64 col][32 row] boxes (a box is rewritten
  // only after the store two boxes back has read it), the fp32-residual epilogue holds the two fp32 [32 col][32 row] boxes
  // of a column pair; 1024 B aligned (TMA SWIZZLE_128B pattern)
  static constexpr int STG_WARP = CONV ? 32 * STG_LD * 4 : 8192;
  static constexpr int STG_BYTES = EW * STG_WARP;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGES_MAX = (227 * 1024 - 1024 - STG_BYTES - BAR_BYTES) / (A_BYTES + B_BYTES);
  static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
  static constexpr int SMEM_BYTES = 1024 + STAGES * (A_BYTES + B_BYTES) + STG_BYTES + BAR_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
};

struct GemmDev {
  int M, N, K;
  int k_blocks;    // K-blocks per tile
  int raster_n;    // tile order: 1 = N blocks fastest, 0 = M blocks fastest
  int gelu_erf;    // EPI_BF16_GELU: 1 = exact erf GELU (launched as the EPI_BF16_GELU_ERF instantiation), 0 = tanh approximation
  int m_blocks;    // 128-row M tiles per batch entry (per image for conv)
  int m_units;     // scheduling units per batch entry: m_blocks (CG 1) or ceil(m_blocks / 2) pairs (CG 2)
  int n_blocks;
  int batch;
  int num_tiles;
  // conv geometry
  int H, Wd, tiles_x, c_blocks;
  int ntap, off_y, off_x;          // taps per axis (3 or 2), halo origin offset
  int s2;                          // single-tap stages: one pipeline stage = (channel chunk, tap), A = ONE [8][16]-pixel box.
                                   // Serves the stride-2 3x3 conv (pad right / bottom; the box is gathered at pixel stride 2
                                   // by the tensor map's elementStrides) and 1x1 convs on the conv epilogue
  int tap_w, tap_stride;           // single-tap stages: taps per axis (3 or 1), source pixel = stride * (y, x) + (ky, kx)
  int o_scale, o_oy, o_ox, oH, oW; // output pixel mapping (phase of an upsample-folded conv)
  int gn_slot_off, gn_slots_img;
  int wres;      // conv, weight-resident mode: number of A stages (0 = off); the whole per-CTA weight slab stays in smem
  int w_tiles;   // weight tiles [B_ROWS x 64] of the slab = ntap^2 * c_blocks, tile t = (ky*ntap + kx)*c_blocks + cb
  // epilogue
  float alpha;
  const float* bias;
  long stride_bias;
  int a_shared;  // all batch entries read A at batch coordinate 0
  int dbg_nomma; // experiment: skip the MMAs (times the TMA/L2 path alone; results are garbage)
  bf16* out_bf16;
  const bf16* resid_bf16;
  long ldo_b, stride_ob;
  float* out_f32;
  const float* resid_f32;
  long ldo_f, stride_of;
  const float* gate;
  long gate_ld;
  int rows_per_gate;
  // fused GroupNorm statistics
  float* gn_partial;
  int gn_cpg, gn_rows_per_img;
  // EPI_QKV
  bf16 *q_heads, *k_heads, *vt_heads;
  int qkv_T, qkv_Tp, qkv_H, qkv_hd;
  // EPI_ATTN (row-owner epilogue only): 1 = per-(row, 64-column pair) maxima of acc, no matrix output; 2 = out_bf16 =
  // exp2(alpha * acc - att_row[row]) + per-(row, pair) sums; 3 = out_bf16 = acc * att_row[row]. att_out[pair * M + row]
  int att_mode;
  const float* att_row;
  float* att_out;
  int red_add;        // EPI_F32 row-owner epilogue, in-place update without a bf16 copy: x += gate * (acc + bias) leaves as a
                      // TMA reduce-add store (the residual is never read by the SM)
  int sm_limit;       // host only: SM budget of the launch (0 = all)
  int lean_depth2;    // lean fp32-residual epilogue: residual rows of the first two chunks requested before the accumulator wait
  int row_path;       // linear GEMM: row-owner epilogue with TMA-store boxes (host-checked alignment), else the transposing one
  long long* trace;   // IR_DEBUG builds: %globaltimer stamps of the roles of CTA 0 ([16] int64), else unused
};

#ifdef IR_DEBUG
static constexpr bool kDebugBuild = true;
#else
static constexpr bool kDebugBuild = false;
#endif
#ifdef IR_DEBUG
IR_DEVINL long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define IR_STAMP(slot)                                                                  \
  do {                                                                                  \
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) p.trace[slot] = gtimer(); \
  } while (0)
#else
#define IR_STAMP(slot) \
  do {                 \
  } while (0)
#endif


IR_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 256-bit global accesses of the direct row-owner epilogue (sm_100: LDG / STG .256): one full 32-byte sector per thread.
// asm volatile keeps them in program order among themselves (the fp32 output may alias the residual, which is only read
// through ldg_v8); no memory clobber, so that the compiler may hoist the read-only bias / gate loads across the stores.
IR_DEVINL void ldg_v8(const float* src, float* r) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(src));
}
IR_DEVINL void stg_v8(float* dst, const float (&x)[8]) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(x[0]), "f"(x[1]), "f"(x[2]),
               "f"(x[3]), "f"(x[4]), "f"(x[5]), "f"(x[6]), "f"(x[7]));
}
IR_DEVINL void stg_v8(bf16* dst, const uint32_t (&x)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(x[0]), "r"(x[1]), "r"(x[2]),
               "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
}

// ---------------------------------------------------------------------------------------------- conv epilogue helpers
// TMA store / bulk-group primitives and swizzled staging access of the conv epilogue (CONV && EPI_BF16)
IR_DEVINL void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
IR_DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// TMA reduce-store: global[box] += shared[box] (fp32 add performed in L2; every element is touched once per launch)
IR_DEVINL void tma_reduce_add_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
IR_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
IR_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
IR_DEVINL void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
IR_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
IR_DEVINL uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
IR_DEVINL void sts_u4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Sum NV per-lane values over the 32 lanes of a warp with a butterfly reduce-scatter: log2(NV) steps in which partner
// lanes split the value array (NV/2, NV/4, ... shuffles) followed by plain xor steps on the remaining lane bits --
// NV - 1 + (5 - log2 NV) shuffles instead of 5 * NV, in a FIXED order (deterministic). Afterwards lane L holds the total
// of value index idx(L) = the bits (4, 3, ...) of L it used for splitting, most significant first; v[0] is that total.
template <int NV>
IR_DEVINL void warp_reduce_scatter(float (&v)[NV], int lane) {
  int off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float give = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, give, off);
    }
  }
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}
// value index lane L holds after warp_reduce_scatter<NV>, and whether L is the lane that writes it
template <int NV>
IR_DEVINL int reduce_scatter_index(int lane, bool& writer) {
  int idx = 0, bit = 4;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, --bit) idx = (idx << 1) | ((lane >> bit) & 1);
  writer = (lane & ((1 << (bit + 1)) - 1)) == 0;
  return idx;
}

// GroupNorm partials of one 32-column chunk: this thread holds one pixel x 32 consecutive channels (fp32, before the
// bf16 rounding -- what the reference normalises); CPG channels per group. The warp's 32 pixels are summed in a fixed
// order and lanes write (sum, sum of squares) of group `g0 + idx / 2` to partial[(slot * 32 + group) * 2 + which].
template <int CPG>
IR_DEVINL void gn_chunk_partials(const float (&a)[32], int lane, float* __restrict__ partial, long slot, int g0, bool valid) {
  constexpr int GPC = 32 / CPG, NV = 2 * GPC;
  float v[NV];
#pragma unroll
  for (int g = 0; g < GPC; ++g) {
    float s_ = 0.f, q_ = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      s_ += a[g * CPG + j];
      q_ = fmaf(a[g * CPG + j], a[g * CPG + j], q_);
    }
    v[2 * g] = s_;
    v[2 * g + 1] = q_;
  }
  warp_reduce_scatter<NV>(v, lane);
  bool writer;
  const int idx = reduce_scatter_index<NV>(lane, writer);
  if (writer && valid) partial[(slot * 32 + g0 + (idx >> 1)) * 2 + (idx & 1)] = v[0];
}

template <int BN, int EPI, bool CONV, int CG>
__global__ void __launch_bounds__(NUM_THREADS_MAX, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmF, const GemmDev p) {
  using Cfg = GemmCfg<BN, CG, CONV, EPI>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(STAGES >= 2, "tile configuration does not fit a double-buffered pipeline");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Weight-resident convs (p.wres): when a CTA's whole weight slab fits beside the A ring it is loaded ONCE per
  // (persistent) CTA instead of once per tile: [slab: w_tiles x B_TAP_BYTES][A ring: wres stages]. The full-resolution
  // N = 128, C = 128 convs are bound by L2 -> shared-memory traffic, more than half of which is the re-streamed weights.
  const bool wres = CONV && p.wres > 0;
  const int nstages = wres ? p.wres : STAGES;
  const uint32_t slab_bytes = wres ? (uint32_t)p.w_tiles * Cfg::B_TAP_BYTES : 0u;
  const uint32_t ring_bytes = wres ? slab_bytes + (uint32_t)p.wres * Cfg::A_BYTES : (uint32_t)(STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
  uint8_t* smA = wres ? smem + slab_bytes : smem;
  uint8_t* smB = wres ? smem : smem + STAGES * Cfg::A_BYTES;
  float* staging = reinterpret_cast<float*>(smem + ring_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_bytes + Cfg::STG_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;       // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* w_full = bars + 2 * STAGES + 6;   // weight slab landed (resident mode)
  uint64_t* r_bar = bars + 2 * STAGES + 7;    // [EW] conv epilogue: residual tile of warp w landed
  static_assert((2 * STAGES + 7 + Cfg::EW) * 8 <= Cfg::BAR_BYTES, "barrier block");

  // warp index through a shuffle: provably warp-uniform, so the producer / issuer branches are convergent and the
  // uniform-datapath instructions (UTMALDG, UTCHMMA, UTCBAR) are emitted directly; under a divergent `lane == 0` the
  // compiler wraps every one of them in an elect-one loop (~80 cycles per MMA: more than a 128 x 128 x 16 MMA takes)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0) IR_STAMP(0);   // kernel entry

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], CG);    // CG 2: the leader's expect_tx arrive + the peer producer's remote arrive
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], Cfg::EW * CG);  // epilogue warps of every CTA of the pair release the accumulator
    }
    mbar_init(w_full, CG);
    for (int i = 0; i < Cfg::EW; ++i) mbar_init(&r_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    if (CG == 2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail
  if (warp == 0) IR_STAMP(1);   // prologue done
  pdl_wait();
  pdl_launch();
  if (warp == 0) IR_STAMP(2);   // predecessor complete

  const int mb_total = p.m_units * p.batch;
  const int unit0 = (int)blockIdx.x / CG, unit_stride = (int)gridDim.x / CG;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (convergent warp, one lane issues)
    const bool leader = elect_one_sync();
    if (wres && unit0 < p.num_tiles) {
      // the CTA's weight slab (n_blocks == 1: rows [cta_rank * B_ROWS, +B_ROWS) of W, every K tile), once
      if (leader) {
        if (CG == 2) {
          if (cta_rank == 0)
            mbar_arrive_expect_tx(w_full, 2 * slab_bytes);
          else
            mbar_arrive_cluster(w_full, 0);
        } else {
          mbar_arrive_expect_tx(w_full, slab_bytes);
        }
        for (int t = 0; t < p.w_tiles; ++t) {
          if (CG == 2)
            tma_load_3d_2sm(smB + t * Cfg::B_TAP_BYTES, &tmW, w_full, t * BK, (int)cta_rank * Cfg::B_ROWS, 0);
          else
            tma_load_3d(smB + t * Cfg::B_TAP_BYTES, &tmW, w_full, t * BK, 0, 0);
        }
      }
      __syncwarp();
    }
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit0; tile < p.num_tiles; tile += unit_stride) {
        // raster_n: consecutive tiles walk the N blocks of one M block (A streamed from HBM once, W stays in L2)
        const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
        const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
        const int b = mb / p.m_units;
        const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;   // may exceed m_blocks (odd tail): all OOB -> zeros
        int ty = 0, tx = 0;
        if (CONV) {
          ty = m_blk / p.tiles_x;
          tx = m_blk - ty * p.tiles_x;
        }
        const int n_row0 = n_blk * BN + (int)cta_rank * Cfg::B_ROWS;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) {
          // conv: one halo box + ntap weight tiles per stage (ntap = 3, or 2 for an upsample-folded phase conv)
          const uint32_t stage_tx = CONV ? (p.s2 ? (uint32_t)(BM * BK * 2 + Cfg::B_TAP_BYTES)
                                                 : (uint32_t)(Cfg::A_BYTES + (wres ? 0 : p.ntap * Cfg::B_TAP_BYTES)))
                                         : (uint32_t)(Cfg::A_BYTES + Cfg::B_BYTES);
          if (CG == 2) {
            if (cta_rank == 0)
              mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_tx);
            else
              mbar_arrive_cluster(&full_bar[stage], 0);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          }
          if (CONV && p.s2) {
            // single-tap stage kb = (channel chunk cb, tap). Stride-2 conv: source pixel (2 y + ky, 2 x + kx), pixels beyond
            // the right / bottom edge (the reference's F.pad(x, (0,1,0,1))) are zero-filled by the TMA unit. 1x1 conv: one tap
            const int ntaps = p.tap_w * p.tap_w;
            const int cb = kb / ntaps;
            const int tap = kb - cb * ntaps;
            const int ky = tap / p.tap_w, kx = tap - ky * p.tap_w;
            const int kcol = (tap * p.c_blocks + cb) * BK;
            const int sx = p.tap_stride * tx * CONV_BW + kx, sy = p.tap_stride * ty * CONV_BH + ky;
            if (CG == 2) {
              tma_load_4d_2sm(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * BK, sx, sy, b);
              tma_load_3d_2sm(smB + stage * Cfg::B_BYTES, &tmW, &full_bar[stage], kcol, n_row0, 0);
            } else {
              tma_load_4d(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * BK, sx, sy, b);
              tma_load_3d(smB + stage * Cfg::B_BYTES, &tmW, &full_bar[stage], kcol, n_row0, 0);
            }
          } else if (CONV) {
            // stage kb = (channel chunk cb, kx): one halo box + the ky weight tiles
            const int cb = kb / p.ntap;
            const int kx = kb - cb * p.ntap;
            if (CG == 2)
              tma_load_4d_2sm(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * BK, tx * CONV_BW + kx + p.off_x,
                              ty * CONV_BH + p.off_y, b);
            else
              tma_load_4d(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], cb * BK, tx * CONV_BW + kx + p.off_x,
                          ty * CONV_BH + p.off_y, b);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              if (ky >= p.ntap || wres) break;
              const int kcol = ((ky * p.ntap + kx) * p.c_blocks + cb) * BK;
              if (CG == 2)
                tma_load_3d_2sm(smB + stage * Cfg::B_BYTES + ky * Cfg::B_TAP_BYTES, &tmW, &full_bar[stage], kcol, n_row0, 0);
              else
                tma_load_3d(smB + stage * Cfg::B_BYTES + ky * Cfg::B_TAP_BYTES, &tmW, &full_bar[stage], kcol, n_row0, 0);
            }
          } else {
            if (CG == 2) {
              tma_load_3d_2sm(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM, p.a_shared ? 0 : b);
              tma_load_3d_2sm(smB + stage * Cfg::B_BYTES, &tmW, &full_bar[stage], kb * BK, n_row0, b);
            } else {
              tma_load_3d(smA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * BK, m_blk * BM, p.a_shared ? 0 : b);
              tma_load_3d(smB + stage * Cfg::B_BYTES, &tmW, &full_bar[stage], kb * BK, n_row0, b);
            }
          }
          }   // leader
          __syncwarp();
          if (tile == unit0 && kb == 0) IR_STAMP(3);   // first stage issued
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      IR_STAMP(4);   // producer done
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (convergent warp, one lane issues)
    if (cta_rank == 0) {
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (wres && unit0 < p.num_tiles) {
        mbar_wait(w_full, 0);
        tc_fence_after();
      }
      for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (it == 0 && kb == 0) IR_STAMP(5);   // first operands landed
          const int TAPS = CONV ? (p.s2 ? 1 : p.ntap) : 1;
          if (CG == 1 && p.dbg_nomma) {
            if (leader) {
              mbar_arrive(&empty_bar[stage]);
              if (kb == p.k_blocks - 1) mbar_arrive(&tfull_bar[buf]);
            }
            __syncwarp();
            if (++stage == nstages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          if (leader) {
#pragma unroll
          for (int ky = 0; ky < (CONV ? 3 : 1); ++ky) {
            if (ky >= TAPS) break;
            // conv: tap ky reads rows [16*ky, 16*ky + 128) of the halo box (+2048 B keeps the 1024 B swizzle phase)
            const uint64_t da = make_smem_desc_sw128(smem_u32(smA + stage * Cfg::A_BYTES + ky * (CONV_BW * BK * 2)));
            // resident slab: tile (ky, kx, cb) with kb = cb * ntap + kx; streamed: the stage's ky-th weight tile
            const uint8_t* bt = wres ? smB + ((ky * p.ntap + (kb % p.ntap)) * p.c_blocks + kb / p.ntap) * Cfg::B_TAP_BYTES
                                     : smB + stage * Cfg::B_BYTES + ky * Cfg::B_TAP_BYTES;
            const uint64_t db = make_smem_desc_sw128(smem_u32(bt));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 32 B (16 bf16) inside the swizzle row: +2 in the 16 B-granular address field
              if (CG == 2)
                umma_bf16_2sm(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | ky | k) != 0);
              else
                umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | ky | k) != 0);
            }
          }
          if (CG == 2) {
            umma_commit_2sm(&empty_bar[stage], 3);  // frees the smem slot in both CTAs
            if (kb == p.k_blocks - 1) umma_commit_2sm(&tfull_bar[buf], 3);
          } else {
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            if (kb == p.k_blocks - 1) umma_commit(&tfull_bar[buf]);
          }
          }   // leader
          __syncwarp();
          if (kb == p.k_blocks - 1) IR_STAMP(6 + (it < 3 ? it : 3));   // tile `it` fully issued (slots 6..9)
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global memory
    // Which code serves which launch (the host picks; p.row_path: 0 generic, 1 TMA-box row-owner, 2 direct row-owner, 3 lean
    // transposing). Every form computes the same per-element arithmetic in fp32 from the same accumulator, so a result never
    // depends on the tile configuration or the batch it was computed in.
    //   conv, bf16 output                pixel-owner epilogue with TMA-store boxes, fused residual / GroupNorm partials
    //   qkv head scatter (EPI_QKV)       direct row-owner (2): TMEM -> registers -> global, no shared memory; generic (0) when
    //                                    head_dim % 8 != 0 or the outputs are not 16-byte aligned
    //   plain bf16 (q_linear, caption)   TMA-box row-owner (1); lean (3) / generic (0) when the output cannot be a TMA box
    //   GELU (fc1), fp32 residual        lean transposing (3): coalesced 128-byte row segments through a per-warp staging
    //   (proj, cross-proj, fc2, zero-    tile, per-tile offsets and masks; generic (0) for N % 32 != 0, misaligned rows,
    //   linears)                         a bf16 copy on another leading dimension
    //   bf16 + residual / GroupNorm      generic transposing (0): the VAE's 1x1 convs as GEMMs, conv instantiations other
    //   partials, fp32 conv outputs      than bf16
    //   EPI_ATTN (VAE mid-attention)     TMA-box row-owner (1) only
    constexpr int EW = Cfg::EW;
    constexpr int CSTEP = EW / 4;           // warps per TMEM lane quarter = chunk interleave
    const int q = warp & 3;                 // TMEM lane quarter this warp may access: lanes [32q, 32q+32)
    const int chalf = (warp - 4) >> 2;      // which interleaved set of 32-column chunks this warp owns
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(staging) + (warp - 4) * Cfg::STG_WARP);
    const uint32_t stg_s = smem_u32(stg);
    const int col4 = (lane & 7) * 4;
    const int rsub = lane >> 3;
    // fused GroupNorm statistics: per-lane (sum, sum of squares) of its 4 channels per 32-column chunk over the rows of
    // ONE tile, reduced by a fixed shuffle tree and written to the slot of (image, tile, warp). Per-tile partials do not
    // depend on how tiles are scheduled or batched, so the statistics are bit-reproducible for any sharding.
    if constexpr (CONV && EPI == EPI_BF16) {
      // -------- conv, bf16 output: TMEM -> registers (one pixel x 32 channels per thread) -> bias / residual / GroupNorm
      // partials -> bf16 -> this warp's 4 KB staging box in the TMA SWIZZLE_128B layout -> ONE TMA store per 64-channel
      // box ([64 ch][16 x][2 y]: the warp's TMEM lane quarter is two rows of the 16 x 8 pixel patch). No shared-memory
      // transpose, no per-row address arithmetic or bounds predicates (the tensor map clips partial tiles and carries the
      // phase mapping of the upsample-folded convs in its strides), stores drain asynchronously. The residual tile
      // (ResnetBlock x + h, model.py:151) is fetched by TMA into the same box while the tile's MMAs still run. GroupNorm
      // partials are per (image, tile, lane quarter): no barrier between the epilogue warps.
      uint8_t* box = reinterpret_cast<uint8_t*>(staging) + (warp - 4) * 4096;
      const uint32_t box_s = smem_u32(box);
      const uint32_t my_row = box_s + (uint32_t)lane * 128u;
      uint64_t* rb = &r_bar[warp - 4];
      const bool has_resid = p.resid_bf16 != nullptr;
      const bool gn_on = p.gn_partial != nullptr;
      constexpr int PAIRS = BN / 64;   // 64-channel boxes per tile
      uint32_t rphase = 0;
      int it = 0;
      for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
        const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
        const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
        const int b = mb / p.m_units;
        const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        const int ty = m_blk / p.tiles_x, tx = m_blk - ty * p.tiles_x;
        const int px = tx * CONV_BW, py = ty * CONV_BH + 2 * q;
        const bool tile_ok = m_blk < p.m_blocks;
        const long slot = ((long)b * (p.gn_slots_img ? p.gn_slots_img : p.m_blocks) + p.gn_slot_off + m_blk) * 4 + q;
        const bool box_any = tile_ok && px < p.Wd && py < p.H;   // a box entirely outside the image neither loads nor stores
        const bool resid_t = has_resid && box_any;
        if (resid_t && chalf < PAIRS && n_blk * BN + chalf * 64 < p.N) {   // first box of this warp: prefetch its residual under the tile's mainloop
          if (lane == 0) {
            bulk_wait_read();   // the previous store out of this box has been read
            mbar_arrive_expect_tx(rb, 4096);
            tma_load_4d(box, &tmR, rb, n_blk * BN + chalf * 64, px, py, b);
          }
          __syncwarp();
        }
        mbar_wait(&tfull_bar[buf], use & 1);
        tc_fence_after();
        if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));
        if (chalf >= PAIRS) {   // narrow tiles (BN = 64 with 8 epilogue warps): this warp owns no box, it only releases the buffer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
          }
          continue;
        }
#pragma unroll 1
        for (int pr = chalf; pr < PAIRS; pr += CSTEP) {
          const int ch0 = n_blk * BN + pr * 64;
          const bool last_pair = pr + CSTEP >= PAIRS;
          if (ch0 >= p.N) {   // Cout tail: this 64-channel box does not exist; only the accumulator hand-back remains
            if (last_pair) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
              }
            }
            continue;
          }
          if (resid_t) {
            if (pr != chalf) {
              if (lane == 0) {
                bulk_wait_read();
                mbar_arrive_expect_tx(rb, 4096);
                tma_load_4d(box, &tmR, rb, ch0, px, py, b);
              }
              __syncwarp();
            }
            mbar_wait(rb, rphase);
            rphase ^= 1;
          } else {
            if (lane == 0) bulk_wait_read();
            __syncwarp();
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + pr * 64 + half * 32), v);
            tmem_ld_wait();
            if (last_pair && half == 1) {
              // accumulator fully read: hand the TMEM buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
              }
            }
            float a[32];
            const float* bp = p.bias + ch0 + half * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b4 = (p.bias && ch0 + half * 32 + 4 * j < p.N) ? __ldg(reinterpret_cast<const float4*>(bp) + j)
                                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
              a[4 * j] = __uint_as_float(v[4 * j]) * p.alpha + b4.x;
              a[4 * j + 1] = __uint_as_float(v[4 * j + 1]) * p.alpha + b4.y;
              a[4 * j + 2] = __uint_as_float(v[4 * j + 2]) * p.alpha + b4.z;
              a[4 * j + 3] = __uint_as_float(v[4 * j + 3]) * p.alpha + b4.w;
            }
            if (resid_t) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 r4 = lds_u4(my_row + (uint32_t)(((half * 4 + j) ^ (lane & 7)) << 4));
                const float2 r0 = unpack_bf16x2(r4.x), r1 = unpack_bf16x2(r4.y), r2 = unpack_bf16x2(r4.z), r3 = unpack_bf16x2(r4.w);
                a[8 * j] += r0.x; a[8 * j + 1] += r0.y; a[8 * j + 2] += r1.x; a[8 * j + 3] += r1.y;
                a[8 * j + 4] += r2.x; a[8 * j + 5] += r2.y; a[8 * j + 6] += r3.x; a[8 * j + 7] += r3.y;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts_u4(my_row + (uint32_t)(((half * 4 + j) ^ (lane & 7)) << 4),
                     make_uint4(pack_bf16x2(a[8 * j], a[8 * j + 1]), pack_bf16x2(a[8 * j + 2], a[8 * j + 3]),
                                pack_bf16x2(a[8 * j + 4], a[8 * j + 5]), pack_bf16x2(a[8 * j + 6], a[8 * j + 7])));
            if (gn_on) {
              // pixels outside the image (partial tiles) hold conv results of zero-padded input: they are not part of the
              // tensor and must not enter the statistics
              const bool inside = (px + (lane & 15) < p.Wd) && (py + (lane >> 4) < p.H);
              if (!inside) {
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = 0.f;
              }
              const int g0 = (ch0 + half * 32) / p.gn_cpg;
              if (p.gn_cpg == 4) gn_chunk_partials<4>(a, lane, p.gn_partial, slot, g0, tile_ok && ch0 + half * 32 < p.N);
              else if (p.gn_cpg == 8) gn_chunk_partials<8>(a, lane, p.gn_partial, slot, g0, tile_ok && ch0 + half * 32 < p.N);
              else gn_chunk_partials<16>(a, lane, p.gn_partial, slot, g0, tile_ok && ch0 + half * 32 < p.N);
            }
          }
          fence_proxy_async();   // generic-proxy writes of the box -> async-proxy (TMA) read
          __syncwarp();
          if (lane == 0 && box_any) {
            tma_store_4d(&tmO, box, ch0, px, py, b);
            bulk_commit();
          }
          __syncwarp();
        }
        if (warp == 4) IR_STAMP(12 + (it < 1 ? 0 : 1));
      }
      if (lane == 0) bulk_wait_all();   // the boxes are in global memory before the CTA retires
      __syncwarp();
    } else {
    bool row_path_done = false;
    // release builds carry the direct epilogue only where it is the default (the qkv scatter); the variants that measured a tie
    // or a loss exist in IR_DEBUG builds for A/B runs (IR_GEMM_DIRECT)
    if constexpr (!CONV && (EPI == EPI_QKV || (kDebugBuild && EPI != EPI_ATTN))) {
      if (p.row_path == 2) {
        // -------- linear GEMMs, direct row-owner epilogue: a thread owns one output row (its TMEM lane) and walks its
        // warp's 32-column chunks straight from TMEM to global memory -- no shared-memory staging at all. The transposing
        // epilogue moves every fp32 element through shared memory twice (256 KB per 128 x 256 tile) while the mainloop
        // already runs the 128 B/clk port at its limit, and exposes the TMEM / shared-memory latencies of each chunk with
        // only two warps per scheduler to hide them. Here the next chunk's accumulator (bf16 outputs) or residual row
        // (fp32-residual epilogue) is in flight while the current chunk is computed, and a thread writes whole 32-byte
        // sectors of its own row (256-bit stores), so the L2 sees full-sector writes in 32 different lines per instruction.
        // Host-checked: N % 32 == 0, 32-byte aligned rows.
        row_path_done = true;
        constexpr int NCH = (BN / 32 + CSTEP - 1) / CSTEP;   // chunks per warp and tile
        const uint32_t lane_tmem = tmem_base + ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
          const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
          const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
          const int b = mb / p.m_units;
          const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;
          const int buf = it & 1;
          const uint32_t use = (uint32_t)(it >> 1);
          const int row = m_blk * BM + q * 32 + lane;
          const bool row_ok = row < p.M;
          const int colw = n_blk * BN;   // first column of the tile
          auto chunk_col = [&](int i) { return colw + (chalf + i * CSTEP) * 32; };
          auto chunk_ok = [&](int i) { return chalf + i * CSTEP < BN / 32 && chunk_col(i) < p.N; };
          auto release = [&]() {   // accumulator fully read: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
            }
          };
          auto tload = [&](uint32_t (&v)[32], int i) {
            tmem_ld_32x32(lane_tmem + (uint32_t)(buf * BN + (chalf + i * CSTEP) * 32), v);
          };
          // the chunk's 32 bias values (every lane reads the same addresses: broadcast loads, L1-resident), requested before
          // the wait for the accumulator chunk so that their latency is not exposed in front of the first FMA
          auto bload = [&](float4 (&bs)[8], int i) {
            const bool on = p.bias != nullptr && chunk_ok(i);
            const float4* bp = reinterpret_cast<const float4*>(p.bias + (long)b * p.stride_bias + chunk_col(i));
#pragma unroll
            for (int j = 0; j < 8; ++j) bs[j] = on ? __ldg(bp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          };
          // alpha * acc + bias of one 8-column piece
          auto lin8 = [&](const uint32_t (&v)[32], int j8, const float4 (&bs)[8], float (&x)[8]) {
            const float4 b0 = bs[2 * j8], b1 = bs[2 * j8 + 1];
            x[0] = fmaf(__uint_as_float(v[8 * j8]), p.alpha, b0.x);
            x[1] = fmaf(__uint_as_float(v[8 * j8 + 1]), p.alpha, b0.y);
            x[2] = fmaf(__uint_as_float(v[8 * j8 + 2]), p.alpha, b0.z);
            x[3] = fmaf(__uint_as_float(v[8 * j8 + 3]), p.alpha, b0.w);
            x[4] = fmaf(__uint_as_float(v[8 * j8 + 4]), p.alpha, b1.x);
            x[5] = fmaf(__uint_as_float(v[8 * j8 + 5]), p.alpha, b1.y);
            x[6] = fmaf(__uint_as_float(v[8 * j8 + 6]), p.alpha, b1.z);
            x[7] = fmaf(__uint_as_float(v[8 * j8 + 7]), p.alpha, b1.w);
          };

          if constexpr (EPI == EPI_F32) {
            // x += gate * (alpha * acc + bias) on the fp32 residual stream (PixArtMS.py:71-79), optional bf16 copy. The residual
            // row piece of the NEXT chunk is requested before the current one is computed; the first one before the wait for
            // the accumulator, so that its latency runs under the tile's own MMAs (every element is read and written by the
            // same thread of the same tile: nothing these loads can see is still to be written).
            const int gate_row = row_ok ? row / p.rows_per_gate : 0;
            float ra[32], rb[32];
            float4 bs[8];
            auto rload = [&](float (&r)[32], int i) {
              if (p.resid_f32 && row_ok && chunk_ok(i)) {
                const float* src = p.resid_f32 + (long)b * p.stride_of + (long)row * p.ldo_f + chunk_col(i);
#pragma unroll
                for (int j = 0; j < 4; ++j) ldg_v8(src + 8 * j, &r[8 * j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0.f;
              }
            };
            auto emit = [&](const uint32_t (&v)[32], const float (&r)[32], int i) {
              if (!row_ok || !chunk_ok(i)) return;
              const int colc = chunk_col(i);
              const float* gp = p.gate ? p.gate + (long)gate_row * p.gate_ld + colc : nullptr;
              float* of = p.out_f32 + (long)b * p.stride_of + (long)row * p.ldo_f + colc;
              bf16* ob = p.out_bf16 ? p.out_bf16 + (long)b * p.stride_ob + (long)row * p.ldo_b + colc : nullptr;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                uint32_t pk[8];
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                  const int j8 = 2 * h + jj;
                  float x[8];
                  lin8(v, j8, bs, x);
                  if (gp) {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gp) + 2 * j8);
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gp) + 2 * j8 + 1);
                    x[0] *= g0.x; x[1] *= g0.y; x[2] *= g0.z; x[3] *= g0.w;
                    x[4] *= g1.x; x[5] *= g1.y; x[6] *= g1.z; x[7] *= g1.w;
                  }
#pragma unroll
                  for (int e = 0; e < 8; ++e) x[e] += r[8 * j8 + e];
                  stg_v8(of + 8 * j8, x);
#pragma unroll
                  for (int e = 0; e < 4; ++e) pk[4 * jj + e] = pack_bf16x2(x[2 * e], x[2 * e + 1]);
                }
                if (ob) stg_v8(ob + 16 * h, pk);
              }
            };
            rload(ra, 0);
            mbar_wait(&tfull_bar[buf], use & 1);
            tc_fence_after();
            if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));
            uint32_t v[32];
#pragma unroll 1
            for (int i = 0; i < NCH; i += 2) {
              tload(v, i);
              bload(bs, i);
              tmem_ld_wait();
              if (NCH > 1) rload(rb, i + 1); else release();
              emit(v, ra, i);
              if constexpr (NCH > 1) {
                tload(v, i + 1);
                bload(bs, i + 1);
                tmem_ld_wait();
                if (i + 2 < NCH) rload(ra, i + 2); else release();
                emit(v, rb, i + 1);
              }
            }
          } else {
            const int qT = (EPI == EPI_QKV) ? p.qkv_T : 1;
            const int bb = (EPI == EPI_QKV) ? row / qT : 0;
            const int tt = (EPI == EPI_QKV) ? row - bb * qT : 0;
            float4 bs[8];
            auto emit = [&](const uint32_t (&v)[32], int i) {
              if (!row_ok || !chunk_ok(i)) return;
              const int colc = chunk_col(i);
              if constexpr (EPI == EPI_QKV) {
                // head-major scatter of the qkv projection (q, k: [b][head][t][hd]; v transposed: [b][head][hd][Tp]); H * hd is a
                // multiple of 32, so a chunk is all-q, all-k or all-v (warp-uniform), and hd a multiple of 8, so an 8-column
                // piece never straddles a head
                const int D1 = p.qkv_H * p.qkv_hd;
                const int which = colc / D1;
                const int cl = colc - which * D1;
                if (which == 2) {
                  // lanes are consecutive tokens: every column is one 64-byte run of the transposed layout
                  bf16* dst = p.vt_heads + ((long)bb * D1 + cl) * p.qkv_Tp + tt;
#pragma unroll
                  for (int j8 = 0; j8 < 4; ++j8) {
                    float x[8];
                    lin8(v, j8, bs, x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) dst[(long)(8 * j8 + e) * p.qkv_Tp] = __float2bfloat16(x[e]);
                  }
                } else {
                  bf16* base = which == 0 ? p.q_heads : p.k_heads;
                  int head = cl / p.qkv_hd, d = cl - head * p.qkv_hd;
#pragma unroll
                  for (int j8 = 0; j8 < 4; ++j8) {
                    float x[8];
                    lin8(v, j8, bs, x);
                    bf16* dst = base + (((long)bb * p.qkv_H + head) * p.qkv_T + tt) * p.qkv_hd + d;
                    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                                                                pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                    d += 8;
                    if (d >= p.qkv_hd) {
                      d -= p.qkv_hd;
                      ++head;
                    }
                  }
                }
              } else {
                bf16* ob = p.out_bf16 + (long)b * p.stride_ob + (long)row * p.ldo_b + colc;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  uint32_t pk[8];
#pragma unroll
                  for (int jj = 0; jj < 2; ++jj) {
                    float x[8];
                    lin8(v, 2 * h + jj, bs, x);
                    if constexpr (EPI == EPI_BF16_GELU_ERF) {   // exact GELU (nn.GELU default: SwinIR's Mlp)
#pragma unroll
                      for (int e = 0; e < 8; ++e) x[e] = 0.5f * x[e] * (1.0f + erff(x[e] * 0.70710678118654752f));
                    } else if constexpr (EPI == EPI_BF16_GELU) {   // tanh approximation (PixArt's Mlp, approximate="tanh")
#pragma unroll
                      for (int e = 0; e < 8; ++e) x[e] = gelu_tanh_fast(x[e]);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) pk[4 * jj + e] = pack_bf16x2(x[2 * e], x[2 * e + 1]);
                  }
                  stg_v8(ob + 16 * h, pk);
                }
              }
            };
            mbar_wait(&tfull_bar[buf], use & 1);
            tc_fence_after();
            if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));
            uint32_t va[32], vb[32];
            tload(va, 0);
#pragma unroll 1
            for (int i = 0; i < NCH; i += 2) {
              bload(bs, i);
              tmem_ld_wait();
              if (NCH > 1) tload(vb, i + 1); else release();
              emit(va, i);
              if constexpr (NCH > 1) {
                bload(bs, i + 1);
                tmem_ld_wait();
                if (i + 2 < NCH) tload(va, i + 2); else release();
                emit(vb, i + 1);
              }
            }
          }
          if (warp == 4) IR_STAMP(12 + (it < 1 ? 0 : 1));
        }
      }
    }
    if constexpr (!CONV && EPI == EPI_F32) {
      if (p.row_path == 3) {
        // -------- fp32-residual epilogue, lean transposing form: x += gate * (alpha * acc + bias) in place on the fp32 stream
        // (PixArtMS.py:71-79) with an optional bf16 copy. Same data movement as the generic transposing epilogue below
        // (TMEM -> registers -> per-warp staging -> coalesced 128-byte row segments), but nothing is recomputed inside the
        // row loop: the ncu instruction mix of the generic code on the proj GEMM (M 4096, N = K = 1152) was 18.6 k warp
        // instructions per 128 x 128 tile against ~2 k of arithmetic and memory instructions -- per-row 64-bit index
        // arithmetic, null-pointer and tail predicates, register copies of the prefetched residual -- i.e. the epilogue was
        // bound by issue slots, not by the 160 KB it moves. Here a lane keeps eight 32-bit element offsets and a validity
        // mask per tile, the residual of the next chunk ping-pongs between two register sets, and the per-tile conditions
        // (bias, gate, per-row gates of a tile that straddles two samples, bf16 copy) are warp-uniform branches around whole
        // chunk bodies. Host-checked: N % 32 == 0, 16-byte aligned rows, one leading dimension for the fp32 stream and its
        // bf16 copy, offsets below 2^31.
        row_path_done = true;
        constexpr int NCH = (BN / 32 + CSTEP - 1) / CSTEP;   // chunks per warp and tile
        const uint32_t lane_tmem = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool has_resid = p.resid_f32 != nullptr;
        const bool has_copy = p.out_bf16 != nullptr;
        int it = 0;
        for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
          const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
          const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
          const int b = mb / p.m_units;
          const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;
          const int buf = it & 1;
          const uint32_t use = (uint32_t)(it >> 1);
          // rows this lane touches after the transpose: q * 32 + i * 4 + rsub, i = 0..7
          const int gm0 = m_blk * BM + q * 32 + rsub;
          uint32_t off[8];
          uint32_t valid = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int gm = gm0 + 4 * i;
            off[i] = (uint32_t)gm * (uint32_t)p.ldo_f + (uint32_t)(n_blk * BN + col4);
            valid |= (gm < p.M ? 1u : 0u) << i;
          }
          const float* rbase = has_resid ? p.resid_f32 + (long)b * p.stride_of : nullptr;
          float* fbase = p.out_f32 + (long)b * p.stride_of;
          bf16* bbase = has_copy ? p.out_bf16 + (long)b * p.stride_ob : nullptr;
          const float* biasp = p.bias ? p.bias + (long)b * p.stride_bias + n_blk * BN + col4 : nullptr;
          // gate: one row of the table per sample; a tile whose rows belong to one sample reads it once per chunk
          const int row_lo = m_blk * BM, row_hi = min(row_lo + BM, p.M) - 1;
          const bool any_rows = row_lo < p.M;   // the odd tail of a CTA pair owns no rows: it reads no gate row either
          const int g_lo = any_rows ? row_lo / p.rows_per_gate : 0;
          const bool gate_uniform = p.gate == nullptr || !any_rows || g_lo == row_hi / p.rows_per_gate;
          const float* gatep = p.gate ? p.gate + (long)g_lo * p.gate_ld + n_blk * BN + col4 : nullptr;
          auto ccol = [&](int k) { return (chalf + k * CSTEP) * 32; };   // column offset of this warp's k-th chunk in the tile
          auto cok = [&](int k) { return n_blk * BN + ccol(k) < p.N; };
          auto rload = [&](float4 (&r)[8], int k) {
            if (has_resid && cok(k)) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                r[i] = (valid >> i & 1u) ? *reinterpret_cast<const float4*>(rbase + off[i] + ccol(k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          };
          auto chunk = [&](const float4 (&r)[8], int k, bool last) {
            uint32_t v[32];
            tmem_ld_32x32(lane_tmem + (uint32_t)(buf * BN + ccol(k)), v);
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), gate4 = make_float4(1.f, 1.f, 1.f, 1.f);
            const bool ok = cok(k);
            if (ok) {
              if (biasp) bias4 = __ldg(reinterpret_cast<const float4*>(biasp + ccol(k)));
              if (gatep) gate4 = __ldg(reinterpret_cast<const float4*>(gatep + ccol(k)));
            }
            tmem_ld_wait();
            if (last) {   // accumulator fully read: hand the TMEM buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_f4(stg_s + (uint32_t)(lane * STG_LD + 4 * j) * 4u,
                     make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3])));
            __syncwarp();
            if (ok) {
              float4 av[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) av[i] = lds_f4(stg_s + (uint32_t)((i * 4 + rsub) * STG_LD + col4) * 4u);
              // premultiplied: x = r + (alpha * gate) * acc + gate * bias
              const float4 ag = make_float4(p.alpha * gate4.x, p.alpha * gate4.y, p.alpha * gate4.z, p.alpha * gate4.w);
              const float4 gb = make_float4(gate4.x * bias4.x, gate4.y * bias4.y, gate4.z * bias4.z, gate4.w * bias4.w);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (!(valid >> i & 1u)) continue;
                float4 a;
                if (gate_uniform) {
                  a.x = fmaf(av[i].x, ag.x, gb.x) + r[i].x;
                  a.y = fmaf(av[i].y, ag.y, gb.y) + r[i].y;
                  a.z = fmaf(av[i].z, ag.z, gb.z) + r[i].z;
                  a.w = fmaf(av[i].w, ag.w, gb.w) + r[i].w;
                } else {   // the tile straddles two samples: this row's own gate
                  const int gm = gm0 + 4 * i;
                  const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gate + (long)(gm / p.rows_per_gate) * p.gate_ld +
                                                                          n_blk * BN + col4 + ccol(k)));
                  // the same arithmetic as the uniform case (a row's result must not depend on which tile it lands in: results
                  // are bit-identical for every batch composition / sharding)
                  a.x = fmaf(av[i].x, p.alpha * g4.x, g4.x * bias4.x) + r[i].x;
                  a.y = fmaf(av[i].y, p.alpha * g4.y, g4.y * bias4.y) + r[i].y;
                  a.z = fmaf(av[i].z, p.alpha * g4.z, g4.z * bias4.z) + r[i].z;
                  a.w = fmaf(av[i].w, p.alpha * g4.w, g4.w * bias4.w) + r[i].w;
                }
                *reinterpret_cast<float4*>(fbase + off[i] + ccol(k)) = a;
                if (has_copy)
                  *reinterpret_cast<uint2*>(bbase + off[i] + ccol(k)) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
              }
            }
            __syncwarp();   // the staging area is rewritten by the next chunk
          };
          float4 ra[8], rb[8];
          // the residual rows of the first TWO chunks are requested before the wait for the accumulator: their latency (the
          // stream may have left the L2 since the previous block touched it) runs under the tile's own MMAs, and the second
          // chunk does not wait out a DRAM round trip behind the short first one
          rload(ra, 0);
          if (NCH > 1 && p.lean_depth2) rload(rb, 1);
          mbar_wait(&tfull_bar[buf], use & 1);
          tc_fence_after();
          if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));
#pragma unroll
          for (int k = 0; k < NCH; k += 2) {
            if (k + 1 < NCH && !(p.lean_depth2 && k == 0)) rload(rb, k + 1);
            chunk(ra, k, k + 1 >= NCH);
            if (k + 1 < NCH) {
              if (k + 2 < NCH) rload(ra, k + 2);
              chunk(rb, k + 1, k + 2 >= NCH);
            }
          }
          if (warp == 4) IR_STAMP(12 + (it < 1 ? 0 : 1));
        }
      }
    }
    if constexpr (!CONV && (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_BF16_GELU_ERF)) {
      if (p.row_path == 3) {
        // -------- bf16 / GELU epilogue, lean transposing form (see the fp32-residual one above): out = act(alpha * acc + bias).
        // The accumulator chunk of the next iteration is in flight (second register set) while the current one is staged,
        // transposed and stored; per-tile offsets and validity mask instead of per-row index arithmetic.
        row_path_done = true;
        constexpr int NCH = (BN / 32 + CSTEP - 1) / CSTEP;
        const uint32_t lane_tmem = tmem_base + ((uint32_t)(q * 32) << 16);
        int it = 0;
        for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
          const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
          const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
          const int b = mb / p.m_units;
          const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;
          const int buf = it & 1;
          const uint32_t use = (uint32_t)(it >> 1);
          const int gm0 = m_blk * BM + q * 32 + rsub;
          uint32_t off[8];
          uint32_t valid = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int gm = gm0 + 4 * i;
            off[i] = (uint32_t)gm * (uint32_t)p.ldo_b + (uint32_t)(n_blk * BN + col4);
            valid |= (gm < p.M ? 1u : 0u) << i;
          }
          bf16* bbase = p.out_bf16 + (long)b * p.stride_ob;
          const float* biasp = p.bias ? p.bias + (long)b * p.stride_bias + n_blk * BN + col4 : nullptr;
          auto ccol = [&](int k) { return (chalf + k * CSTEP) * 32; };
          auto cok = [&](int k) { return n_blk * BN + ccol(k) < p.N; };
          auto tload = [&](uint32_t (&v)[32], int k) { tmem_ld_32x32(lane_tmem + (uint32_t)(buf * BN + ccol(k)), v); };
          auto release = [&]() {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
            }
          };
          auto chunk = [&](const uint32_t (&v)[32], int k) {
            const bool ok = cok(k);
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && biasp) bias4 = __ldg(reinterpret_cast<const float4*>(biasp + ccol(k)));
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts_f4(stg_s + (uint32_t)(lane * STG_LD + 4 * j) * 4u,
                     make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3])));
            __syncwarp();
            if (ok) {
              float4 av[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) av[i] = lds_f4(stg_s + (uint32_t)((i * 4 + rsub) * STG_LD + col4) * 4u);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (!(valid >> i & 1u)) continue;
                float4 a;
                a.x = fmaf(av[i].x, p.alpha, bias4.x);
                a.y = fmaf(av[i].y, p.alpha, bias4.y);
                a.z = fmaf(av[i].z, p.alpha, bias4.z);
                a.w = fmaf(av[i].w, p.alpha, bias4.w);
                if constexpr (EPI == EPI_BF16_GELU_ERF) {   // exact GELU (nn.GELU default: SwinIR's Mlp)
                  a.x = 0.5f * a.x * (1.0f + erff(a.x * 0.70710678118654752f));
                  a.y = 0.5f * a.y * (1.0f + erff(a.y * 0.70710678118654752f));
                  a.z = 0.5f * a.z * (1.0f + erff(a.z * 0.70710678118654752f));
                  a.w = 0.5f * a.w * (1.0f + erff(a.w * 0.70710678118654752f));
                } else if constexpr (EPI == EPI_BF16_GELU) {   // tanh approximation (PixArt's Mlp, approximate="tanh")
                  a.x = gelu_tanh_fast(a.x);
                  a.y = gelu_tanh_fast(a.y);
                  a.z = gelu_tanh_fast(a.z);
                  a.w = gelu_tanh_fast(a.w);
                }
                *reinterpret_cast<uint2*>(bbase + off[i] + ccol(k)) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
              }
            }
            __syncwarp();   // the staging area is rewritten by the next chunk
          };
          mbar_wait(&tfull_bar[buf], use & 1);
          tc_fence_after();
          if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));
          uint32_t va[32], vb[32];
          tload(va, 0);
#pragma unroll
          for (int k = 0; k < NCH; k += 2) {
            tmem_ld_wait();
            if (k + 1 < NCH) tload(vb, k + 1); else release();
            chunk(va, k);
            if (k + 1 < NCH) {
              tmem_ld_wait();
              if (k + 2 < NCH) tload(va, k + 2); else release();
              chunk(vb, k + 1);
            }
          }
          if (warp == 4) IR_STAMP(12 + (it < 1 ? 0 : 1));
        }
      }
    }
    if constexpr (!CONV && EPI != EPI_QKV) {
      if (p.row_path == 1) {
        // -------- linear GEMMs, row-owner epilogue: a thread owns one output row (its TMEM lane) and 32 consecutive columns
        // per chunk. bf16 outputs: two chunks fill one 128 B row of a [64 col][32 row] box in the TMA SWIZZLE_128B layout and
        // leave by ONE TMA store per box. fp32-residual epilogue (x += gate * (A W^T + b), PixArtMS.py:71-79): the residual
        // rows arrive by TMA into two fp32 boxes ([32 col][32 row], fetched while the tile's MMAs still run), are updated in
        // place and stored back by TMA, the bf16 copy of the new stream leaves through a third box. No shared-memory
        // transpose, no per-row address arithmetic or predicates (the tensor maps clip the M / N tails), stores drain
        // asynchronously behind the next chunk.
        row_path_done = true;
        constexpr bool F32 = (EPI == EPI_F32);
        uint8_t* wbase = reinterpret_cast<uint8_t*>(staging) + (warp - 4) * Cfg::STG_WARP;
        const uint32_t rowf0 = smem_u32(wbase) + (uint32_t)lane * 128u, rowf1 = rowf0 + 4096u;
        int nbox = 0;   // bf16 outputs: boxes alternate
        uint64_t* rb = &r_bar[warp - 4];
        const bool red_add = F32 && p.red_add;
        const bool has_resid = F32 && p.resid_f32 != nullptr && !red_add;
        const bool want_bf16 = !F32 || p.out_bf16 != nullptr;
        constexpr int PAIRS = BN / 64;
        uint32_t rphase = 0;
        int it = 0;
        for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
          const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
          const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
          const int b = mb / p.m_units;
          const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;
          const int buf = it & 1;
          const uint32_t use = (uint32_t)(it >> 1);
          const int row0 = m_blk * BM + q * 32;
          const bool row_ok = row0 + lane < p.M;   // M tail: rows beyond M are zero-filled on load and clipped on store
          const int gate_row = row_ok ? (row0 + lane) / p.rows_per_gate : 0;
          const bool rows_any = row0 < p.M;           // a lane quarter entirely beyond M neither loads nor stores
          const bool resid_t = has_resid && rows_any;
          float att_r = 0.f;   // EPI_ATTN: this row's shift (mode 2) or scale (mode 3)
          if (EPI == EPI_ATTN && p.att_row && row_ok) att_r = __ldg(p.att_row + row0 + lane);
          if (resid_t && chalf < PAIRS && n_blk * BN + chalf * 64 < p.N) {   // first column pair of this warp: residual rows fetched under the mainloop
            if (lane == 0) {
              const int c0 = n_blk * BN + chalf * 64;
              const bool two = c0 + 32 < p.N;   // a box that starts beyond N is neither fetched nor stored
              bulk_wait_read();
              mbar_arrive_expect_tx(rb, two ? 8192 : 4096);
              tma_load_3d(wbase, &tmR, rb, c0, row0, b);
              if (two) tma_load_3d(wbase + 4096, &tmR, rb, c0 + 32, row0, b);
            }
            __syncwarp();
          }
          mbar_wait(&tfull_bar[buf], use & 1);
          tc_fence_after();
          if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));
          if (chalf >= PAIRS) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
            }
            continue;
          }
#pragma unroll 1
          for (int pr = chalf; pr < PAIRS; pr += CSTEP) {
            const int col0 = n_blk * BN + pr * 64;
            const bool last_pair = pr + CSTEP >= PAIRS;
            if (col0 >= p.N) {   // N tail: this column pair does not exist; only the accumulator hand-back remains
              if (last_pair) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
                }
              }
              continue;
            }
            if (resid_t) {
              if (pr != chalf) {
                if (lane == 0) {
                  const bool two = col0 + 32 < p.N;
                  bulk_wait_read();
                  mbar_arrive_expect_tx(rb, two ? 8192 : 4096);
                  tma_load_3d(wbase, &tmR, rb, col0, row0, b);
                  if (two) tma_load_3d(wbase + 4096, &tmR, rb, col0 + 32, row0, b);
                }
                __syncwarp();
              }
              mbar_wait(rb, rphase);
              rphase ^= 1;
            } else {
              if (lane == 0) {
                if (F32) bulk_wait_read(); else bulk_wait_read1();   // bf16: only the store two boxes back must have drained
              }
              __syncwarp();
            }
            uint8_t* boxb = wbase + (nbox & 1) * 4096;
            const uint32_t rowb = smem_u32(boxb) + (uint32_t)lane * 128u;
            ++nbox;
            float att_acc = (EPI == EPI_ATTN && p.att_mode == 1) ? -INFINITY : 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t v[32];
              tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + pr * 64 + half * 32), v);
              tmem_ld_wait();
              if (last_pair && half == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
                }
              }
              const int colh = col0 + half * 32;
              const float* bp = p.bias + (long)b * p.stride_bias + colh;
              const float* gp = p.gate + (long)gate_row * p.gate_ld + colh;
              const uint32_t rowf = half ? rowf1 : rowf0;
              float a[32];
              if constexpr (EPI == EPI_ATTN) {
                // AttnBlock softmax (model.py:195-197) split over three GEMM passes; columns beyond N do not exist
                if (p.att_mode == 1) {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (colh + j < p.N) att_acc = fmaxf(att_acc, __uint_as_float(v[j]));
                } else if (p.att_mode == 2) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) {
                    const float e = (colh + j < p.N) ? ex2_approx(fmaf(__uint_as_float(v[j]), p.alpha, -att_r)) : 0.f;
                    att_acc += e;
                    a[j] = e;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]) * att_r;
                }
              } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const bool cok = colh + 4 * j < p.N;
                const float4 b4 = (p.bias && cok) ? __ldg(reinterpret_cast<const float4*>(bp) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                float x0 = __uint_as_float(v[4 * j]) * p.alpha + b4.x, x1 = __uint_as_float(v[4 * j + 1]) * p.alpha + b4.y;
                float x2 = __uint_as_float(v[4 * j + 2]) * p.alpha + b4.z, x3 = __uint_as_float(v[4 * j + 3]) * p.alpha + b4.w;
                if constexpr (EPI == EPI_BF16_GELU_ERF) {   // exact GELU (nn.GELU default: SwinIR's Mlp)
                  x0 = 0.5f * x0 * (1.0f + erff(x0 * 0.70710678118654752f));
                  x1 = 0.5f * x1 * (1.0f + erff(x1 * 0.70710678118654752f));
                  x2 = 0.5f * x2 * (1.0f + erff(x2 * 0.70710678118654752f));
                  x3 = 0.5f * x3 * (1.0f + erff(x3 * 0.70710678118654752f));
                } else if constexpr (EPI == EPI_BF16_GELU) {   // tanh approximation (PixArt's Mlp, approximate="tanh")
                  x0 = gelu_tanh_fast(x0);
                  x1 = gelu_tanh_fast(x1);
                  x2 = gelu_tanh_fast(x2);
                  x3 = gelu_tanh_fast(x3);
                }
                if (F32) {
                  if (p.gate) {
                    const float4 g4 = cok ? __ldg(reinterpret_cast<const float4*>(gp) + j) : make_float4(1.f, 1.f, 1.f, 1.f);
                    x0 *= g4.x; x1 *= g4.y; x2 *= g4.z; x3 *= g4.w;
                  }
                  const uint32_t sa = rowf + (uint32_t)((j ^ (lane & 7)) << 4);
                  if (resid_t) {
                    const float4 r4 = lds_f4(sa);
                    x0 += r4.x; x1 += r4.y; x2 += r4.z; x3 += r4.w;
                  }
                  sts_f4(sa, make_float4(x0, x1, x2, x3));
                }
                a[4 * j] = x0; a[4 * j + 1] = x1; a[4 * j + 2] = x2; a[4 * j + 3] = x3;
              }
              }   // !EPI_ATTN
              if (want_bf16 && !(EPI == EPI_ATTN && p.att_mode == 1)) {
                if (F32) {
                  // bf16 copy of the updated stream (next GEMM's A operand): 64 contiguous bytes of this thread's row
                  if (row_ok) {
                    bf16* ob = p.out_bf16 + (long)b * p.stride_ob + (long)(row0 + lane) * p.ldo_b + colh;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      if (colh + 8 * j < p.N)
                        *reinterpret_cast<uint4*>(ob + 8 * j) =
                            make_uint4(pack_bf16x2(a[8 * j], a[8 * j + 1]), pack_bf16x2(a[8 * j + 2], a[8 * j + 3]),
                                       pack_bf16x2(a[8 * j + 4], a[8 * j + 5]), pack_bf16x2(a[8 * j + 6], a[8 * j + 7]));
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    sts_u4(rowb + (uint32_t)(((half * 4 + j) ^ (lane & 7)) << 4),
                           make_uint4(pack_bf16x2(a[8 * j], a[8 * j + 1]), pack_bf16x2(a[8 * j + 2], a[8 * j + 3]),
                                      pack_bf16x2(a[8 * j + 4], a[8 * j + 5]), pack_bf16x2(a[8 * j + 6], a[8 * j + 7])));
                }
              }
            }
            if (EPI == EPI_ATTN && p.att_mode != 3 && row_ok) p.att_out[(long)(col0 >> 6) * p.M + row0 + lane] = att_acc;
            fence_proxy_async();   // generic-proxy writes of the boxes -> async-proxy (TMA) reads
            __syncwarp();
            if (lane == 0 && rows_any && !(EPI == EPI_ATTN && p.att_mode == 1)) {
              if (F32 && red_add) {
                tma_reduce_add_3d(&tmF, wbase, col0, row0, b);
                if (col0 + 32 < p.N) tma_reduce_add_3d(&tmF, wbase + 4096, col0 + 32, row0, b);
              } else if (F32) {
                tma_store_3d(&tmF, wbase, col0, row0, b);
                if (col0 + 32 < p.N) tma_store_3d(&tmF, wbase + 4096, col0 + 32, row0, b);
              }
              if (!F32) tma_store_3d(&tmO, boxb, col0, row0, b);
              bulk_commit();
            }
            __syncwarp();
          }
          if (warp == 4) IR_STAMP(12 + (it < 1 ? 0 : 1));
        }
        if (lane == 0) bulk_wait_all();
        __syncwarp();
      }
    }
    if (!row_path_done) {
    float gn_s[BN / 32 / CSTEP], gn_q[BN / 32 / CSTEP];
    const bool gn_on = (EPI == EPI_BF16) && p.gn_partial != nullptr;
    int it = 0;
    for (int tile = unit0; tile < p.num_tiles; tile += unit_stride, ++it) {
      const int n_blk = p.raster_n ? tile % p.n_blocks : tile / mb_total;
      const int mb = p.raster_n ? tile / p.n_blocks : tile - n_blk * mb_total;
      const int b = mb / p.m_units;
      const int m_blk = (mb - b * p.m_units) * CG + (int)cta_rank;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      if (gn_on) {
#pragma unroll
        for (int c = 0; c < BN / 32 / CSTEP; ++c) {
          gn_s[c] = 0.f;
          gn_q[c] = 0.f;
        }
      }

      // per-lane row bookkeeping for the 8 rows this lane touches after the transpose
      long row_off[8];   // row index into the output (rows of ldo elements), -1 if masked
      int gate_row[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = q * 32 + i * 4 + rsub;
        if (CONV) {
          const int ty = m_blk / p.tiles_x, tx = m_blk - ty * p.tiles_x;
          const int y = ty * CONV_BH + (r >> 4), x = tx * CONV_BW + (r & 15);
          row_off[i] = (y < p.H && x < p.Wd)
                           ? ((long)b * p.oH + y * p.o_scale + p.o_oy) * p.oW + x * p.o_scale + p.o_ox : -1;
          gate_row[i] = 0;
        } else {
          const int gm = m_blk * BM + r;
          row_off[i] = (gm < p.M) ? (long)gm : -1;
          gate_row[i] = gm / p.rows_per_gate;
        }
      }

      // Residual reads are software-pipelined one 32-column chunk ahead (two chunks of loads in flight per warp): the
      // epilogue is latency-bound on these loads, not bandwidth-bound. The first chunk's residual is requested BEFORE the
      // wait for the accumulator, so its latency runs under the tile's own MMAs (each output element is read and written
      // by the same thread of the same tile, so nothing this load can see is still to be written).
      float4 res_cur[8], res_nxt[8];
      uint2 resb_cur[8], resb_nxt[8];
      auto load_resid = [&](int cc, float4 (&r4)[8], uint2 (&rb)[8]) {
        const int ccol = n_blk * BN + cc * 32 + col4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = row_off[i] >= 0 && ccol < p.N;
          if (EPI == EPI_F32) {
            r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && p.resid_f32)
              r4[i] = *reinterpret_cast<const float4*>(p.resid_f32 + (long)b * p.stride_of + row_off[i] * p.ldo_f + ccol);
          } else if (EPI == EPI_BF16) {
            rb[i] = make_uint2(0u, 0u);
            if (ok && p.resid_bf16)
              rb[i] = *reinterpret_cast<const uint2*>(p.resid_bf16 + (long)b * p.stride_ob + row_off[i] * p.ldo_b + ccol);
          }
        }
      };
      if (EPI == EPI_F32 || EPI == EPI_BF16) load_resid(chalf, res_cur, resb_cur);
      const bool gate_uniform = gate_row[0] == gate_row[7];

      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      if (warp == 4) IR_STAMP(10 + (it < 1 ? 0 : 1));   // accumulator of tile `it` complete (slot 10: first, 11: last seen)

      // generic path: the accumulator chunk of the NEXT iteration is requested from TMEM as soon as the current one has
      // been parked in shared memory, so that its latency runs under the current chunk's math and stores
      // (the fp32-residual epilogue already keeps gate, residual and next residual in registers: prefetching there spills)
      // Linear GEMMs only: measured on B200, fc1 (GELU) 39.5 -> 35.1 us; the conv instantiations (GroupNorm partials, fewer
      // epilogue warps) lost ~3 % with it and keep the plain order.
      constexpr bool PREFETCH_ACC = !CONV && (EPI == EPI_BF16 || EPI == EPI_BF16_GELU || EPI == EPI_BF16_GELU_ERF);
      uint32_t vacc[32];
      if (PREFETCH_ACC) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + chalf * 32), vacc);
#pragma unroll 1
      for (int c = chalf; c < BN / 32; c += CSTEP) {
        const bool last_chunk = c + CSTEP >= BN / 32;
        if (EPI == EPI_QKV) {
          // head-major scatter of the qkv projection; D1 = H*hd is a multiple of 32, so a chunk is all-q, all-k or all-v
          const int D1 = p.qkv_H * p.qkv_hd;
          const int cbase = n_blk * BN + c * 32;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), v);
          tmem_ld_wait();
          if (last_chunk) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
            }
          }
          if (cbase >= p.N) continue;
          const int which = cbase / D1;
          if (which == 2) {
            // v is stored transposed ([b][head][d][t]): stage the 32 x 32 chunk, then lane = channel reads its
            // column (conflict-free) and writes 32 consecutive tokens = 64 contiguous bytes
            const int gm0 = m_blk * BM + q * 32;
            const int bb0 = gm0 / p.qkv_T, t0 = gm0 - bb0 * p.qkv_T;
            const bool fast = (gm0 + 32 <= p.M) && (t0 + 32 <= p.qkv_T) && ((t0 & 7) == 0) && ((p.qkv_Tp & 7) == 0);
            if (fast) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 f = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                sts_f4(stg_s + (uint32_t)(lane * STG_LD + 4 * j) * 4u, f);
              }
              __syncwarp();
              const int cl = cbase + lane - 2 * D1;
              const int head = cl / p.qkv_hd, d = cl - head * p.qkv_hd;
              const float bv = p.bias[cbase + lane];
              bf16* dst = p.vt_heads + (((long)bb0 * p.qkv_H + head) * p.qkv_hd + d) * p.qkv_Tp + t0;
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                float e[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) e[i] = lds_f1(stg_s + (uint32_t)((g * 8 + i) * STG_LD + lane) * 4u) + bv;
                *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]),
                                                                    pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
              }
              __syncwarp();
            } else {
              const int gm = gm0 + lane;
              if (gm < p.M) {
                const int bb = gm / p.qkv_T, t = gm - bb * p.qkv_T;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int cl = cbase + j - 2 * D1;
                  const int head = cl / p.qkv_hd, d = cl - head * p.qkv_hd;
                  const float val = __uint_as_float(v[j]) + p.bias[cbase + j];
                  p.vt_heads[(((long)bb * p.qkv_H + head) * p.qkv_hd + d) * p.qkv_Tp + t] = __float2bfloat16(val);
                }
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 f = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                     __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
              sts_f4(stg_s + (uint32_t)(lane * STG_LD + 4 * j) * 4u, f);
            }
            __syncwarp();
            const int cl = cbase + col4 - which * D1;
            const int head = cl / p.qkv_hd, d = cl - head * p.qkv_hd;
            const float4 bias4 = *reinterpret_cast<const float4*>(p.bias + cbase + col4);
            bf16* dst = which == 0 ? p.q_heads : p.k_heads;
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = lds_f4(stg_s + (uint32_t)((i * 4 + rsub) * STG_LD + col4) * 4u);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (row_off[i] < 0) continue;
              const int gm = (int)row_off[i];
              const int bb = gm / p.qkv_T, t = gm - bb * p.qkv_T;
              const float4 a = av[i];
              *reinterpret_cast<uint2*>(dst + (((long)bb * p.qkv_H + head) * p.qkv_T + t) * p.qkv_hd + d) =
                  make_uint2(pack_bf16x2(a.x + bias4.x, a.y + bias4.y), pack_bf16x2(a.z + bias4.z, a.w + bias4.w));
            }
            __syncwarp();
          }
          continue;
        }
        const int col = n_blk * BN + c * 32 + col4;
        const bool col_ok = col < p.N;
        // Everything the epilogue reads from global memory is issued BEFORE the TMEM load (and the residual of the NEXT
        // chunk now): the output may alias the residual (in-place x += ...), so loads placed after the first store
        // could not be hoisted by the compiler.
        float4 gate_u = make_float4(1.f, 1.f, 1.f, 1.f);   // gate of the whole tile when its rows share one (the common case)
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        float chunk_s = 0.f, chunk_q = 0.f;
        if (!last_chunk && (EPI == EPI_F32 || EPI == EPI_BF16)) load_resid(c + CSTEP, res_nxt, resb_nxt);
        if (col_ok) {
          if (p.bias) bias4 = *reinterpret_cast<const float4*>(p.bias + (long)b * p.stride_bias + col);
          if (EPI == EPI_F32 && p.gate && gate_uniform)
            gate_u = *reinterpret_cast<const float4*>(p.gate + (long)gate_row[0] * p.gate_ld + col);
        }

        if (!PREFETCH_ACC) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c * 32), vacc);
        tmem_ld_wait();
        if (last_chunk) {
          // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(&tempty_bar[buf], 0); else mbar_arrive(&tempty_bar[buf]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 f = make_float4(__uint_as_float(vacc[4 * j]), __uint_as_float(vacc[4 * j + 1]),
                                 __uint_as_float(vacc[4 * j + 2]), __uint_as_float(vacc[4 * j + 3]));
          sts_f4(stg_s + (uint32_t)(lane * STG_LD + 4 * j) * 4u, f);
        }
        if (PREFETCH_ACC && !last_chunk)   // the registers are free again (the stores above read them at issue): prefetch the next chunk
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + (c + CSTEP) * 32), vacc);
        __syncwarp();

        if (col_ok) {
          // all eight shared-memory reads of the chunk are issued back to back (the asm volatile loads keep their program
          // order, so inside the row loop each would wait out its own latency behind the previous row's stores)
          float4 av[8];
          if (!CONV) {
#pragma unroll
            for (int i = 0; i < 8; ++i) av[i] = lds_f4(stg_s + (uint32_t)((i * 4 + rsub) * STG_LD + col4) * 4u);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (row_off[i] < 0) continue;
            float4 a = !CONV ? av[i] : lds_f4(stg_s + (uint32_t)((i * 4 + rsub) * STG_LD + col4) * 4u);
            a.x = a.x * p.alpha + bias4.x;
            a.y = a.y * p.alpha + bias4.y;
            a.z = a.z * p.alpha + bias4.z;
            a.w = a.w * p.alpha + bias4.w;
            if constexpr (EPI == EPI_BF16_GELU_ERF) {   // exact GELU (nn.GELU default: SwinIR's Mlp)
              a.x = 0.5f * a.x * (1.0f + erff(a.x * 0.70710678118654752f));
              a.y = 0.5f * a.y * (1.0f + erff(a.y * 0.70710678118654752f));
              a.z = 0.5f * a.z * (1.0f + erff(a.z * 0.70710678118654752f));
              a.w = 0.5f * a.w * (1.0f + erff(a.w * 0.70710678118654752f));
            } else if constexpr (EPI == EPI_BF16_GELU) {   // tanh approximation (PixArt's Mlp, approximate="tanh")
              a.x = gelu_tanh_fast(a.x);
              a.y = gelu_tanh_fast(a.y);
              a.z = gelu_tanh_fast(a.z);
              a.w = gelu_tanh_fast(a.w);
            }
            if (EPI == EPI_F32) {
              const long o = (long)b * p.stride_of + row_off[i] * p.ldo_f + col;
              float4 g4 = gate_u;
              if (p.gate && !gate_uniform)   // tile straddles two samples: per-row gate (the gate never aliases the output)
                g4 = __ldg(reinterpret_cast<const float4*>(p.gate + (long)gate_row[i] * p.gate_ld + col));
              a.x = a.x * g4.x + res_cur[i].x;
              a.y = a.y * g4.y + res_cur[i].y;
              a.z = a.z * g4.z + res_cur[i].z;
              a.w = a.w * g4.w + res_cur[i].w;
              *reinterpret_cast<float4*>(p.out_f32 + o) = a;
              if (p.out_bf16) {
                const long ob = (long)b * p.stride_ob + row_off[i] * p.ldo_b + col;
                *reinterpret_cast<uint2*>(p.out_bf16 + ob) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
              }
            } else {
              const long ob = (long)b * p.stride_ob + row_off[i] * p.ldo_b + col;
              if (EPI == EPI_BF16) {
                const float2 r0 = unpack_bf16x2(resb_cur[i].x), r1 = unpack_bf16x2(resb_cur[i].y);
                a.x += r0.x;
                a.y += r0.y;
                a.z += r1.x;
                a.w += r1.y;
              }
              *reinterpret_cast<uint2*>(p.out_bf16 + ob) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
              if (gn_on) {   // statistics of the fp32 values (what the reference normalises), before bf16 rounding
                chunk_s += (a.x + a.y) + (a.z + a.w);
                chunk_q = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, chunk_q))));
              }
            }
          }
        }
        __syncwarp();
        if (gn_on) {
#pragma unroll
          for (int cc = 0; cc < BN / 32 / CSTEP; ++cc) {   // static indexing keeps the accumulators in registers
            if (cc == c / CSTEP) {
              gn_s[cc] += chunk_s;
              gn_q[cc] += chunk_q;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          res_cur[i] = res_nxt[i];
          resb_cur[i] = resb_nxt[i];
        }
      }
      if (gn_on) {
        // reduce over the lanes that share a group (fixed tree), park the per-warp group sums in this warp's staging
        // area, combine the 4 epilogue warps in fixed order and write ONE partial per (image, tile, group)
        const int tiles_per_img = CONV ? p.m_blocks : p.gn_rows_per_img / BM;
        const int img = CONV ? b : m_blk / tiles_per_img;
        const int tile_in_img = CONV ? m_blk : m_blk - img * tiles_per_img;
        const int gpc = 32 / p.gn_cpg;   // groups per 32-column chunk
#pragma unroll
        for (int lc = 0; lc < BN / 32 / CSTEP; ++lc) {
          const int c = lc * CSTEP + chalf;   // global 32-column chunk index of this warp's lc-th chunk
          float s_ = gn_s[lc], q_ = gn_q[lc];
          s_ += __shfl_xor_sync(0xffffffffu, s_, 8);
          q_ += __shfl_xor_sync(0xffffffffu, q_, 8);
          s_ += __shfl_xor_sync(0xffffffffu, s_, 16);
          q_ += __shfl_xor_sync(0xffffffffu, q_, 16);
          if (p.gn_cpg >= 8) {
            s_ += __shfl_xor_sync(0xffffffffu, s_, 1);
            q_ += __shfl_xor_sync(0xffffffffu, q_, 1);
          }
          if (p.gn_cpg >= 16) {
            s_ += __shfl_xor_sync(0xffffffffu, s_, 2);
            q_ += __shfl_xor_sync(0xffffffffu, q_, 2);
          }
          if (rsub == 0 && (col4 % p.gn_cpg) == 0) {
            const int gi = c * gpc + col4 / p.gn_cpg;
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stg_s + (uint32_t)gi * 8u), "f"(s_), "f"(q_) : "memory");
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        if (warp == 4 && m_blk < p.m_blocks) {
          const uint32_t stg0 = smem_u32(staging);
          for (int gi = lane; gi < (BN / 32) * gpc; gi += 32) {
            const int gcol = n_blk * BN + gi * p.gn_cpg;
            if (gcol < p.N) {
              // the group lives in chunk gi / gpc, owned by column-set (chunk % CSTEP): sum its 4 lane-quarter warps
              const int owner = (gi / gpc) % CSTEP;
              float s_ = 0.f, q_ = 0.f;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                float a0, a1;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a0), "=f"(a1)
                             : "r"(stg0 + (uint32_t)((owner * 4 + w) * Cfg::STG_WARP + gi * 8)) : "memory");
                s_ += a0;
                q_ += a1;
              }
              const long slot = (long)img * (CONV && p.gn_slots_img ? p.gn_slots_img : tiles_per_img) +
                                (CONV ? p.gn_slot_off : 0) + tile_in_img;
              *reinterpret_cast<float2*>(p.gn_partial + (slot * 32 + gcol / p.gn_cpg) * 2) = make_float2(s_, q_);
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");   // staging is reused by the next tile's first chunk
      }
      if (warp == 4) IR_STAMP(12 + (it < 1 ? 0 : 1));   // epilogue of tile `it` done (12: first, 13: last seen)
    }
    }   // !row_path_done
    }   // !(CONV && EPI_BF16)
  }

  if (warp == 0) IR_STAMP(14);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();   // the peer may still read this CTA's operands / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (CG == 2)
      tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else
      tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (warp == 0) IR_STAMP(15);   // exit
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// bf16 tensor map, innermost dimension first; strides in bytes for dims 1..rank-1; 128B swizzle, zero OOB fill.
int make_tensor_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int swizzle_bytes, int elem_bytes, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point not available (driver too old or no GPU)");
    return IR_ERR_DRIVER;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                     : (swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d, dims %llu,%llu,%llu base %p)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                   base);
    return IR_ERR_DRIVER;
  }
  return IR_OK;
}

static int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  return make_tensor_map(m, base, rank, dims, strides_bytes, box, 128);
}

static int num_sms() { return device_num_sms(); }

static constexpr int RED_ADD_DEFAULT = 0;   // measured on B200 (interleaved A/B, 1024^2 step): cross-proj / after_proj 23.2 -> 22.8 us, fc2 47.1 -> 47.4 us, step 21.6 vs 21.5-21.8 ms: a wash -> off
static constexpr int ROW_PATH_DEFAULT = 1;   // measured on B200: plain bf16 gains (qkv-like 29.3 -> 27.4 us), GELU and fp32-residual lose (34.9 -> 40.8, 18.1 -> 22.6 us)
static constexpr int DIRECT_DEFAULT = 8;   // direct row-owner epilogue (see gemm_launch), measured on B200 against the incumbents (isolated launches, M 4096 / 25600 / 1024): qkv scatter 30.4 -> 29.4, 160.8 -> 154.4, 15.7 -> 13.1 us: on; plain bf16 a tie (13.7 vs 13.6, 54.2 vs 55.7), GELU 31.9 vs 32.3, fp32-residual 16.2 -> 19.5 us (one 32-byte L2 request per thread and sector: 3x the write requests of the coalesced stores): off
static long long* g_gemm_trace = nullptr;
static int g_gemm_trace_slots = 1, g_gemm_trace_next = 0;
void gemm_set_trace(long long* device_buf, int slots) {
  g_gemm_trace = device_buf;
  g_gemm_trace_slots = slots > 0 ? slots : 1;
  g_gemm_trace_next = 0;
}
int gemm_conv_tiles_per_image(int H, int W) { return ((W + CONV_BW - 1) / CONV_BW) * ((H + CONV_BH - 1) / CONV_BH); }

static bool conv_wres_enabled() {
  static const bool on = [] {
    // Opt-in ("1"): measured neutral on B200 (interleaved A/B at 1024^2: 24.41 vs 24.30 ms; conv class 7.57 vs 7.57 ms), so
    // the full-resolution N = 128 convs are NOT bound by re-streaming their weights; kept for experiments.
    const char* e = debug_env("IR_CONV_WRES");
    return e && e[0] == '1';
  }();
  return on;
}

template <int BN, int EPI, bool CONV, int CG>
static int launch_inst(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const CUtensorMap& tr,
                       const CUtensorMap& tf, const GemmDev& p_in, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG, CONV, EPI>;
  constexpr int SMEM_MAX = 227 * 1024;
  auto kern = gemm_tc_kernel<BN, EPI, CONV, CG>;
  IR_TRY(ensure_smem_optin((const void*)kern, CONV ? SMEM_MAX : Cfg::SMEM_BYTES));
  GemmDev p = p_in;
  const int slots = ((p.sm_limit > 0 && p.sm_limit < num_sms()) ? p.sm_limit : num_sms()) / CG;
  int smem_bytes = Cfg::SMEM_BYTES;
  p.wres = 0;
  p.w_tiles = 0;
  if (CONV && !p.s2 && p.n_blocks == 1 && conv_wres_enabled()) {
    // weight-resident mode: the CTA's slab of ntap^2 * c_blocks weight tiles beside >= 3 A stages, and enough tiles per
    // persistent CTA to amortise loading it
    const long slab = (long)p.ntap * p.ntap * p.c_blocks * Cfg::B_TAP_BYTES;
    long sa = (SMEM_MAX - 1024 - Cfg::STG_BYTES - Cfg::BAR_BYTES - slab) / Cfg::A_BYTES;
    if (sa > Cfg::STAGES) sa = Cfg::STAGES;
    const long tiles_per_cta = (p.num_tiles + slots - 1) / slots;
    if (sa >= 3 && tiles_per_cta >= 3) {
      p.wres = (int)sa;
      p.w_tiles = p.ntap * p.ntap * p.c_blocks;
      smem_bytes = 1024 + (int)slab + (int)sa * Cfg::A_BYTES + Cfg::STG_BYTES + Cfg::BAR_BYTES;
    }
  }
  const int grid = (p.num_tiles < slots ? p.num_tiles : slots) * CG;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = CG;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = (CG == 2) ? 2 : 1;
  const bool prof = prof_enabled();
  if (prof) prof_before(stream);
  IR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ta, tw, to, tr, tf, p));
  if (prof) prof_after(stream, CONV ? PROF_CONV : PROF_GEMM, 2.0 * (double)p.M * p.N * p.K * (CONV ? 1 : p.batch), p.M * (CONV ? 1 : p.batch), p.N, p.K);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

template <int BN, bool CONV, int CG>
static int launch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& to, const CUtensorMap& tr,
                      const CUtensorMap& tf, const GemmDev& p, cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_inst<BN, EPI_BF16, CONV, CG>(ta, tw, to, tr, tf, p, s);
    case EPI_BF16_GELU:   // the exact (erf) variant is its own instantiation: inlined erff bloats the unrolled epilogue code
      if (p.gelu_erf) return launch_inst<BN, EPI_BF16_GELU_ERF, CONV, CG>(ta, tw, to, tr, tf, p, s);
      return launch_inst<BN, EPI_BF16_GELU, CONV, CG>(ta, tw, to, tr, tf, p, s);
    case EPI_F32: return launch_inst<BN, EPI_F32, CONV, CG>(ta, tw, to, tr, tf, p, s);
    case EPI_QKV:
      if (!CONV) return launch_inst<BN, EPI_QKV, false, CG>(ta, tw, to, tr, tf, p, s);
      break;
    case EPI_ATTN:
      if (!CONV) return launch_inst<BN, EPI_ATTN, false, CG>(ta, tw, to, tr, tf, p, s);
      break;
  }
  set_last_error("gemm: unknown epilogue %d", epi);
  return IR_ERR_INVALID;
}

struct TileCfg {
  int cg, bn;
};

// Tile configuration: CTA-pair tiles (cta_group::2, 256 x BN) or single-CTA tiles (128 x BN).
// cost = waves * k_steps * (time per K-block step of one scheduling unit) + the exposed epilogue of the last tile;
// waves = units / (SMs / CG). Step times measured on B200 with the convergent issue path (tools/gpu_kernel_check.py
// perf: 8192^3, M4096 x {N4608 K1152, N1152 K4608, N1152 K1152}): per FLOP the 128-wide pair tile costs about what the
// 256-wide one does, so it wins whenever it quantises better (N = 1152 = 9 x 128).
static TileCfg pick_cfg(long m_blocks_total, long m_pairs_total, int N, int k_steps, int forced, bool conv, int sm_limit = 0) {
  static int env_cg = -1, env_bn = 0;
  if (env_cg < 0) {
    const char* e = debug_env("IR_GEMM_CFG");  // "cg,bn", e.g. "2,256"
    env_cg = 0;
    if (e) sscanf(e, "%d,%d", &env_cg, &env_bn);
  }
  if (forced >= 1000) return TileCfg{forced / 1000, forced % 1000};   // force_bn = cg*1000 + bn
  if (forced == 64 || forced == 128 || forced == 256) return TileCfg{1, forced};
  if ((env_cg == 1 || env_cg == 2) && (env_bn == 64 || env_bn == 128 || env_bn == 256) && !(env_cg == 2 && env_bn == 64))
    return TileCfg{env_cg, env_bn};
  const int sms = (sm_limit > 0 && sm_limit < num_sms()) ? sm_limit : num_sms();
  const TileCfg cands[5] = {{2, 256}, {2, 128}, {1, 256}, {1, 128}, {1, 64}};
  const double unit_cost[5] = {0.415, 0.24, 0.50, 0.285, 0.245};   // us per K-block step
  const double epi_cost[5] = {2.0, 1.0, 2.0, 1.0, 0.6};            // us, last tile's epilogue (not overlapped)
  TileCfg best = cands[3];
  double best_cost = 1e30;
  for (int i = 0; i < 5; ++i) {
    const int bn = cands[i].bn, cg = cands[i].cg;
    if (bn > 64 && N <= bn / 2) continue;  // tile mostly empty
    if (conv && cg == 1 && bn == 256) continue;  // halo-box stages do not fit twice in shared memory
    const long units = (cg == 2 ? m_pairs_total : m_blocks_total) * ((N + bn - 1) / bn);
    const long slots = sms / cg;
    const long waves = (units + slots - 1) / slots;
    const double cost = (double)waves * k_steps * unit_cost[i] + epi_cost[i];
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = cands[i];
    }
  }
  return best;
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream) {
  IR_REQUIRE(a.A && a.W, "gemm: null operand");
  IR_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && a.batch > 0, "gemm: bad shape M=%d N=%d K=%d batch=%d", a.M, a.N, a.K,
             a.batch);
  IR_REQUIRE(a.N % 4 == 0, "gemm: N=%d must be a multiple of 4", a.N);
  IR_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  if (a.epi == EPI_QKV) {
    IR_REQUIRE(a.q_heads && a.k_heads && a.vt_heads && a.bias && !a.conv && a.batch == 1, "gemm: EPI_QKV needs q/k/vt outputs and a bias");
    IR_REQUIRE(a.qkv_hd % 4 == 0 && (a.qkv_H * a.qkv_hd) % 32 == 0 && a.N == 3 * a.qkv_H * a.qkv_hd && a.qkv_T > 0 &&
                   a.M % a.qkv_T == 0 && a.qkv_Tp >= a.qkv_T,
               "gemm: EPI_QKV shape mismatch");
  } else if (a.epi == EPI_ATTN) {
    IR_REQUIRE(!a.conv && a.batch == 1 && a.att_mode >= 1 && a.att_mode <= 3 && !a.bias && !a.resid_bf16 && !a.gn_partial,
               "gemm: EPI_ATTN is a plain single-batch GEMM without bias / residual / statistics");
    IR_REQUIRE(a.att_mode == 1 || (a.out_bf16 && a.ldo_b % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out_bf16) & 15) == 0),
               "gemm: EPI_ATTN modes 2, 3 need a 16-byte aligned out_bf16 with ld %% 8 == 0");
    IR_REQUIRE((a.att_mode == 3 || a.att_out) && (a.att_mode == 1 || a.att_row), "gemm: EPI_ATTN row vectors missing");
  } else if (a.epi == EPI_F32) {
    IR_REQUIRE(a.out_f32 && a.ldo_f % 4 == 0, "gemm: EPI_F32 needs out_f32 with ld %% 4 == 0");
    IR_REQUIRE(!a.out_bf16 || a.ldo_b % 4 == 0, "gemm: bf16 copy needs ld %% 4 == 0");
  } else {
    IR_REQUIRE(a.out_bf16 && a.ldo_b % 4 == 0, "gemm: bf16 epilogue needs out_bf16 with ld %% 4 == 0");
  }
  IR_REQUIRE(!a.gate || a.gate_ld % 4 == 0, "gemm: gate_ld must be a multiple of 4");
  if (a.gn_partial) {
    IR_REQUIRE(a.epi == EPI_BF16 && (a.gn_cpg == 4 || a.gn_cpg == 8 || a.gn_cpg == 16) && a.N == 32 * a.gn_cpg,
               "gemm: fused GroupNorm statistics need EPI_BF16 and N == 32 * cpg");
    IR_REQUIRE(a.conv || (a.batch == 1 && a.gn_rows_per_img % BM == 0), "gemm: fused GroupNorm statistics need rows_per_img %% 128 == 0");
  }

  GemmDev p{};
  p.M = a.M;
  p.N = a.N;
  p.K = a.K;
  p.batch = a.batch;
  p.alpha = a.alpha;
  p.bias = a.bias;
  p.stride_bias = a.conv ? 0 : a.stride_bias;
  p.a_shared = (!a.conv && a.batch > 1 && a.strideA == 0) ? 1 : 0;
  {
    static const int nomma = [] { const char* e = debug_env("IR_GEMM_NOMMA"); return (e && e[0] == '1') ? 1 : 0; }();
    p.dbg_nomma = nomma;
  }
  p.gelu_erf = a.gelu_erf;
  p.att_mode = a.att_mode;
  p.att_row = a.att_row;
  p.att_out = a.att_out;
  p.out_bf16 = a.out_bf16;
  p.resid_bf16 = a.resid_bf16;
  p.ldo_b = a.ldo_b;
  p.stride_ob = a.stride_ob;
  p.out_f32 = a.out_f32;
  p.resid_f32 = a.resid_f32;
  p.ldo_f = a.ldo_f;
  p.stride_of = a.stride_of;
  p.gate = a.gate;
  p.gate_ld = a.gate_ld;
  p.rows_per_gate = a.rows_per_gate > 0 ? a.rows_per_gate : 1;
  p.gn_partial = a.gn_partial;
  p.gn_cpg = a.gn_cpg;
  p.gn_rows_per_img = a.gn_rows_per_img > 0 ? a.gn_rows_per_img : 1;
  p.q_heads = a.q_heads;
  p.k_heads = a.k_heads;
  p.vt_heads = a.vt_heads;
  p.qkv_T = a.qkv_T;
  p.qkv_Tp = a.qkv_Tp;
  p.qkv_H = a.qkv_H;
  p.qkv_hd = a.qkv_hd;
  p.trace = g_gemm_trace ? g_gemm_trace + 16L * (g_gemm_trace_next++ % g_gemm_trace_slots) : nullptr;   // one 16-stamp record per launch

  CUtensorMap ta, tw;
  long m_blocks_total;
  if (a.conv) {
    IR_REQUIRE(a.C % BK == 0 || (a.conv_taps == 1 && a.C < BK && a.C % 8 == 0), "conv: C=%d must be a multiple of %d", a.C, BK);
    IR_REQUIRE(a.conv_taps == 3 || a.conv_taps == 2 || a.conv_taps == 1, "conv: %d taps per axis unsupported", a.conv_taps);
    IR_REQUIRE(a.K == a.conv_taps * a.conv_taps * a.C, "conv: K=%d must equal taps^2*C=%d", a.K, a.conv_taps * a.conv_taps * a.C);
    IR_REQUIRE(a.o_scale >= 1 && a.o_oy >= 0 && a.o_oy < a.o_scale && a.o_ox >= 0 && a.o_ox < a.o_scale, "conv: bad output mapping");
    IR_REQUIRE(a.M == a.nimg * a.H * a.Wd, "conv: M mismatch");
    IR_REQUIRE(a.conv_stride == 1 || (a.conv_stride == 2 && a.conv_taps == 3 && a.o_scale == 1),
               "conv: stride %d unsupported (1, or 2 with a 3x3 kernel)", a.conv_stride);
    IR_REQUIRE(a.conv_taps != 1 || a.o_scale == 1, "conv: a 1x1 conv has no output mapping");
    IR_REQUIRE(a.batch == 1, "conv: batch must be 1 (images are folded into M)");
    p.H = a.H;
    p.Wd = a.Wd;
    p.tiles_x = (a.Wd + CONV_BW - 1) / CONV_BW;
    const int tiles_y = (a.H + CONV_BH - 1) / CONV_BH;
    p.m_blocks = p.tiles_x * tiles_y;
    p.batch = a.nimg;  // one "batch" entry per image
    p.c_blocks = (a.C + BK - 1) / BK;   // C < 64 (1x1 only): one chunk, the TMA unit zero-fills channels C..63
    p.ntap = a.conv_taps;
    p.off_y = a.conv_off_y;
    p.off_x = a.conv_off_x;
    p.o_scale = a.o_scale;
    p.o_oy = a.o_oy;
    p.o_ox = a.o_ox;
    p.oH = a.H * a.o_scale;
    p.oW = a.Wd * a.o_scale;
    p.gn_slot_off = a.gn_slot_off;
    p.gn_slots_img = a.gn_slots_img;
    p.k_blocks = p.ntap * p.c_blocks;   // pipeline steps: (channel chunk, kx), ntap ky taps each
    p.s2 = (a.conv_stride == 2 || a.conv_taps == 1) ? 1 : 0;
    p.tap_w = a.conv_taps;
    p.tap_stride = a.conv_stride;
    if (p.s2) p.k_blocks = p.tap_w * p.tap_w * p.c_blocks;   // (channel chunk, ky, kx), one tap each
    // the conv epilogue indexes the output by pixel; batch strides are folded into the pixel index
    p.stride_ob = 0;
    p.stride_of = 0;
    const uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)a.Wd, (uint64_t)a.H, (uint64_t)a.nimg};
    const uint64_t strides[3] = {(uint64_t)a.C * 2, (uint64_t)a.Wd * a.C * 2, (uint64_t)a.H * a.Wd * a.C * 2};
    const uint32_t box[4] = {(uint32_t)BK, (uint32_t)CONV_BW, (uint32_t)(CONV_BH + 2), 1};   // halo box: ky = 0..2
    if (p.s2) {
      // (H, Wd) is the OUTPUT grid; the input is (st*H, st*Wd). The box names the traversed extent: at element stride 2,
      // 32 x 16 source pixels land as a dense [8][16]-pixel tile (tools/probes/tma_elem_stride_probe.cu: 16384 B completed,
      // row (i, j) = pixel (y0 + 2i, x0 + 2j), zero fill beyond the edges)
      const uint64_t st = (uint64_t)a.conv_stride;
      const uint64_t dims2[4] = {(uint64_t)a.C, (uint64_t)a.Wd * st, (uint64_t)a.H * st, (uint64_t)a.nimg};
      const uint64_t strides2[3] = {(uint64_t)a.C * 2, (uint64_t)a.Wd * st * a.C * 2, (uint64_t)a.H * st * a.Wd * st * a.C * 2};
      const uint32_t box2[4] = {(uint32_t)BK, (uint32_t)(st * CONV_BW), (uint32_t)(st * CONV_BH), 1};
      const uint32_t es2[4] = {1, (uint32_t)st, (uint32_t)st, 1};
      IR_TRY(make_tensor_map(&ta, a.A, 4, dims2, strides2, box2, 128, 2, es2));
    } else {
      IR_TRY(make_map(&ta, a.A, 4, dims, strides, box));
    }
    m_blocks_total = (long)p.m_blocks * a.nimg;
  } else {
    IR_REQUIRE(a.lda % 8 == 0 && a.strideA % 8 == 0, "gemm: lda/strideA must be multiples of 8 elements");
    p.m_blocks = (a.M + BM - 1) / BM;
    p.k_blocks = (a.K + BK - 1) / BK;
    const int ab = p.a_shared ? 1 : a.batch;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.M, (uint64_t)ab};
    const uint64_t strides[2] = {(uint64_t)a.lda * 2, (uint64_t)(ab > 1 ? a.strideA : (long)a.M * a.lda) * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BM, 1};
    IR_TRY(make_map(&ta, a.A, 3, dims, strides, box));
    m_blocks_total = (long)p.m_blocks * a.batch;
  }
  IR_REQUIRE(a.ldw % 8 == 0 && a.strideW % 8 == 0, "gemm: ldw/strideW must be multiples of 8 elements");

  const long m_pairs = (p.m_blocks + 1) / 2;
  TileCfg tc = pick_cfg(m_blocks_total, m_pairs * p.batch, a.N, (a.conv && !p.s2) ? p.ntap * p.k_blocks : p.k_blocks, a.force_bn, a.conv != 0, a.sm_limit);
  p.sm_limit = a.sm_limit;
  if (a.conv && tc.cg == 1 && tc.bn == 256) tc.bn = 128;
  const int bn = tc.bn;
  p.m_units = tc.cg == 2 ? (int)m_pairs : p.m_blocks;
  p.n_blocks = (a.N + bn - 1) / bn;
  p.num_tiles = (int)((long)p.m_units * p.batch * p.n_blocks);
  // the operand that is re-read by the slower-moving tile index should be the one that fits in L2: walk N fastest when
  // the activations outweigh the weights (token GEMMs at M > N: M25600 N1152 K4608 243 -> 211 us), M fastest otherwise
  // (convs measured neutral-to-worse with N fastest, so they keep the M-fastest order)
  p.raster_n = (!a.conv && a.batch == 1 && a.M > a.N) ? 1 : 0;
  {
    const int wb = a.conv ? 1 : a.batch;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)a.N, (uint64_t)wb};
    const uint64_t strides[2] = {(uint64_t)a.ldw * 2, (uint64_t)(wb > 1 ? a.strideW : (long)a.N * a.ldw) * 2};
    const uint32_t box[3] = {(uint32_t)BK, (uint32_t)(bn / tc.cg), 1};
    IR_TRY(make_map(&tw, a.W, 3, dims, strides, box));
  }

  // conv with bf16 output: tensor maps of the output (and residual) tiles for the TMA-store epilogue. Box = [64 ch][16 x][2 y]
  // (one TMEM lane quarter of the 16 x 8 pixel patch); the output mapping (y * o_scale + o_oy, x * o_scale + o_ox) of the
  // upsample-folded phase convs is carried by the base offset and the strides.
  CUtensorMap to = ta, tr = ta;
  if (a.conv && a.epi == EPI_BF16) {
    IR_REQUIRE(a.ldo_b == a.N && a.N % 8 == 0, "conv: bf16 output must be dense NHWC with Cout %% 8 == 0 (ldo %ld, Cout %d)", a.ldo_b, a.N);
    IR_REQUIRE((reinterpret_cast<uintptr_t>(a.out_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.resid_bf16) & 15) == 0,
               "conv: output / residual must be 16-byte aligned");
    const uint64_t N = (uint64_t)a.N, oW = (uint64_t)p.oW, oH = (uint64_t)p.oH, sc = (uint64_t)a.o_scale;
    const uint64_t dims[4] = {N, (uint64_t)a.Wd, (uint64_t)a.H, (uint64_t)a.nimg};
    const uint64_t strides[3] = {sc * N * 2, sc * oW * N * 2, oH * oW * N * 2};
    const uint32_t box[4] = {64, (uint32_t)CONV_BW, 2, 1};
    const long base_off = ((long)a.o_oy * p.oW + a.o_ox) * a.N;
    IR_TRY(make_map(&to, a.out_bf16 + base_off, 4, dims, strides, box));
    if (a.resid_bf16) IR_TRY(make_map(&tr, a.resid_bf16 + base_off, 4, dims, strides, box));
  }

  // linear GEMMs: tensor maps of the row-owner epilogue (bf16 box [64 col][32 row], fp32 boxes [32 col][32 row]) when the
  // outputs meet TMA's alignment rules; otherwise the transposing epilogue serves the launch
  CUtensorMap tf = ta;
  p.row_path = 0;
  if (!a.conv && a.epi != EPI_QKV && !a.resid_bf16 && !a.gn_partial) {
    const bool f32 = a.epi == EPI_F32;
    const bool want_b = (!f32 || a.out_bf16 != nullptr) && !(a.epi == EPI_ATTN && a.att_mode == 1);
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    bool ok = true;
    if (want_b) ok = ok && al16(a.out_bf16) && a.ldo_b % 8 == 0 && (a.batch == 1 || a.stride_ob % 8 == 0);
    if (f32) ok = ok && al16(a.out_f32) && al16(a.resid_f32) && a.ldo_f % 4 == 0 && (a.batch == 1 || a.stride_of % 4 == 0);
    if (ok) {
      const uint64_t dims[3] = {(uint64_t)a.N, (uint64_t)a.M, (uint64_t)a.batch};
      if (want_b) {
        const uint64_t st[2] = {(uint64_t)a.ldo_b * 2, (uint64_t)(a.batch > 1 ? a.stride_ob : (long)a.M * a.ldo_b) * 2};
        const uint32_t box[3] = {64, 32, 1};
        IR_TRY(make_map(&to, a.out_bf16, 3, dims, st, box));
      }
      if (f32) {
        const uint64_t st[2] = {(uint64_t)a.ldo_f * 4, (uint64_t)(a.batch > 1 ? a.stride_of : (long)a.M * a.ldo_f) * 4};
        const uint32_t box[3] = {32, 32, 1};
        IR_TRY(make_tensor_map(&tf, a.out_f32, 3, dims, st, box, 128, 4));
        if (a.resid_f32) IR_TRY(make_tensor_map(&tr, a.resid_f32, 3, dims, st, box, 128, 4));
      }
      // which epilogues take the row-owner path: bit 0 plain bf16, bit 1 GELU, bit 2 fp32-residual (IR_GEMM_ROWPATH: A/B
      // switch of debug builds)
      static const int row_mask = [] {
        const char* e = debug_env("IR_GEMM_ROWPATH");
        return e ? atoi(e) : ROW_PATH_DEFAULT;
      }();
      const int bit = a.epi == EPI_BF16 ? 1 : (a.epi == EPI_BF16_GELU ? 2 : 4);
      p.row_path = (row_mask & bit) ? 1 : 0;
      // in-place residual update with no bf16 copy (cross-attention proj, fc2, after_proj: PixArtMS.py:76-79,
      // pixart_controlnet.py:240): row-owner epilogue whose boxes leave by TMA reduce-add (IR_GEMM_REDADD=0: A/B switch
      // of debug builds)
      static const int red_on = [] {
        const char* e = debug_env("IR_GEMM_REDADD");
        return e ? atoi(e) : RED_ADD_DEFAULT;
      }();
      if (f32 && red_on && a.resid_f32 && a.resid_f32 == a.out_f32 && !a.out_bf16) {
        p.row_path = 1;
        p.red_add = 1;
      }
      if (a.epi == EPI_ATTN) p.row_path = 1;   // the only epilogue that implements it
    }
  }
  IR_REQUIRE(a.epi != EPI_ATTN || p.row_path, "gemm: EPI_ATTN needs the row-owner epilogue");
  // direct row-owner epilogue (TMEM -> registers -> 256-bit global stores, no shared-memory staging): whole 32-column chunks
  // and 32-byte aligned rows. Bits of the mask: 1 plain bf16, 2 GELU, 4 fp32-residual, 8 qkv scatter (IR_GEMM_DIRECT: A/B
  // switch of debug builds)
  if (!a.conv && a.epi != EPI_ATTN && !a.resid_bf16 && !a.gn_partial && !p.red_add && a.N % 32 == 0) {
    static const int direct_mask = [] {
      const char* e = debug_env("IR_GEMM_DIRECT");
      return e ? atoi(e) : DIRECT_DEFAULT;
    }();
    auto al = [](const void* q, uintptr_t n) { return (reinterpret_cast<uintptr_t>(q) & (n - 1)) == 0; };
    const int bit = a.epi == EPI_BF16 ? 1 : a.epi == EPI_BF16_GELU ? 2 : a.epi == EPI_F32 ? 4 : 8;
    bool ok = (direct_mask & bit) != 0 && (kDebugBuild || a.epi == EPI_QKV) && (!a.bias || (al(a.bias, 16) && a.stride_bias % 4 == 0));
    const bool ob_ok = al(a.out_bf16, 32) && a.ldo_b % 16 == 0 && (a.batch == 1 || a.stride_ob % 16 == 0);
    if (a.epi == EPI_QKV) {
      ok = ok && al(a.q_heads, 16) && al(a.k_heads, 16) && a.qkv_hd % 8 == 0;
    } else if (a.epi == EPI_F32) {
      ok = ok && al(a.out_f32, 32) && al(a.resid_f32, 32) && a.ldo_f % 8 == 0 && (a.batch == 1 || a.stride_of % 8 == 0) &&
           (!a.out_bf16 || ob_ok) && (!a.gate || al(a.gate, 16));
    } else {
      ok = ok && ob_ok;
    }
    if (ok) p.row_path = 2;
  }
  // lean transposing fp32-residual epilogue (see the kernel): whole 32-column chunks, 16-byte aligned rows, the bf16 copy on the
  // stream's own leading dimension, 32-bit element offsets (IR_GEMM_LEAN=0: A/B switch of debug builds)
  if (!a.conv && a.epi == EPI_F32 && p.row_path != 2 && !p.red_add && a.N % 32 == 0) {
    static const int lean_on = [] {
      const char* e = debug_env("IR_GEMM_LEAN");
      return e ? atoi(e) : 2;
    }();
    p.lean_depth2 = lean_on >= 2 ? 1 : 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool ok = lean_on && al16(a.out_f32) && al16(a.resid_f32) && a.ldo_f % 4 == 0 && a.stride_of % 4 == 0 &&
                    (long)a.M * a.ldo_f < (1L << 31) &&
                    (!a.out_bf16 || (al16(a.out_bf16) && a.ldo_b == a.ldo_f && a.stride_ob % 8 == 0)) &&
                    (!a.bias || (al16(a.bias) && a.stride_bias % 4 == 0)) && (!a.gate || al16(a.gate));
    if (ok) p.row_path = 3;
  }
  // the same lean form for the GELU epilogue (and for plain bf16 outputs that cannot take the TMA-store path)
  if (!a.conv && (a.epi == EPI_BF16 || a.epi == EPI_BF16_GELU) && p.row_path == 0 && !a.resid_bf16 && !a.gn_partial && a.N % 32 == 0) {
    static const int lean_on = [] {
      const char* e = debug_env("IR_GEMM_LEAN_BF16");
      return e ? atoi(e) : 1;
    }();
    auto al = [](const void* q, uintptr_t n) { return (reinterpret_cast<uintptr_t>(q) & (n - 1)) == 0; };
    const bool ok = lean_on && al(a.out_bf16, 8) && a.ldo_b % 4 == 0 && a.stride_ob % 4 == 0 && (long)a.M * a.ldo_b < (1L << 31) &&
                    (!a.bias || (al(a.bias, 16) && a.stride_bias % 4 == 0));
    if (ok) p.row_path = 3;
  }

  if (a.conv) {
    if (tc.cg == 2) {
      if (bn == 256) return launch_epi<256, true, 2>(a.epi, ta, tw, to, tr, tf, p, stream);
      return launch_epi<128, true, 2>(a.epi, ta, tw, to, tr, tf, p, stream);
    }
    if (bn == 64) return launch_epi<64, true, 1>(a.epi, ta, tw, to, tr, tf, p, stream);
    return launch_epi<128, true, 1>(a.epi, ta, tw, to, tr, tf, p, stream);
  }
  if (tc.cg == 2) {
    if (bn == 256) return launch_epi<256, false, 2>(a.epi, ta, tw, to, tr, tf, p, stream);
    return launch_epi<128, false, 2>(a.epi, ta, tw, to, tr, tf, p, stream);
  }
  if (bn == 64) return launch_epi<64, false, 1>(a.epi, ta, tw, to, tr, tf, p, stream);
  if (bn == 128) return launch_epi<128, false, 1>(a.epi, ta, tw, to, tr, tf, p, stream);
  return launch_epi<256, false, 1>(a.epi, ta, tw, to, tr, tf, p, stream);
}

int gemm_conv_gn_slots_per_tile() { return 4; }

}  // namespace ir
