// Fused flash-attention forward on the warp-level mma.sync path: softmax(Q K^T / sqrt(d)) V without materialising the
// score matrix. In the product path it serves
//   * cross-attention to the packed caption tokens (MultiHeadCrossAttention.forward, PixArt_blocks.py:43-58,
//     BlockDiagonalMask.from_seqlens([T]*B, y_lens)): sample b attends to kv rows [kv_off[b], kv_off[b]+kv_len[b]) --
//     <= 300 keys, latency-bound: a CTA walks 2-4 consecutive query tiles with the caption K/V staged once and the next
//     Q tile prefetched, the grid sized to one resident wave;
// and, behind IR_ATTN_LEGACY=1 only, self-attention (AttentionKVCompress.forward, PixArt_blocks.py:123-158), whose
// product kernel is the tcgen05 / TMEM one in attention_tc.cu.
// head_dim is 72 (1152/16): rows are 144 B in global memory, staged to shared memory rows of 88 elements with the
// K dimension zero-padded to 80 for the m16n8k16 tensor-core MMAs. fp32 online softmax (exp2 domain).
#include "attention.cuh"

#include <cstdlib>

namespace ir {

static constexpr int HD = 72;        // head dim
static constexpr int HDK = 80;       // head dim padded to a multiple of 16 (QK^T reduction)
static constexpr int LDS = 88;       // smem row stride in elements (176 B: conflict-free ldmatrix)
static constexpr int BKV = 64;       // keys per tile
static constexpr int NT_S = BKV / 8; // score n-tiles per warp
static constexpr int NT_O = HD / 8;  // output n-tiles per warp (9)

IR_DEVINL void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
IR_DEVINL void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
IR_DEVINL void ldsm_x2_t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
IR_DEVINL void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

struct AttnDev {
  const bf16* q;
  const bf16* k;
  const bf16* v;
  bf16* out;
  long ldq, ldk, ldv, ldo;
  int QT;               // query tiles per CTA (consecutive tiles of one (sample, head); K/V staged once when they fit)
  int Tq;               // queries per sample
  int Tk;               // keys per sample when kv_len == nullptr
  const int* kv_off;    // [B] first kv row of sample b (nullptr: b*Tk)
  const int* kv_len;    // [B] number of kv rows of sample b (nullptr: Tk)
  float scale_log2e;    // d^-1/2 * log2(e)
};

template <int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) flash_attn_kernel(const AttnDev p) {
  constexpr int BQ = NWARPS * 16;
  constexpr int NTHREADS = NWARPS * 32;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  bf16* sQ = reinterpret_cast<bf16*>(smem_attn);   // [2][BQ][LDS]: the next query tile is prefetched under the current one
  bf16* sK = sQ + 2 * BQ * LDS;    // [2][BKV][LDS]
  bf16* sV = sK + 2 * BKV * LDS;   // [2][BKV][LDS]

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int n_qtiles = (p.Tq + BQ - 1) / BQ;
  const int qt0 = blockIdx.x * p.QT;
  const int nq = min(p.QT, n_qtiles - qt0);

  pdl_wait();
  pdl_launch();
  const int kv_start = p.kv_off ? p.kv_off[b] : b * p.Tk;
  const int kv_n = p.kv_len ? p.kv_len[b] : p.Tk;
  const int n_tiles = (kv_n + BKV - 1) / BKV;
  // captions (<= 128 valid tokens in practice) fit the two K/V buffers: staged once per CTA, reused by every query tile
  const bool resident = n_tiles <= 2;

  const bf16* qg = p.q + (long)b * p.Tq * p.ldq + head * HD;
  const bf16* kg = p.k + (long)kv_start * p.ldk + head * HD;
  const bf16* vg = p.v + (long)kv_start * p.ldv + head * HD;

  // zero the K-padding columns [72, 80) of both Q and both K buffers (cp.async never writes them)
  for (int r = tid; r < 2 * BQ + 2 * BKV; r += NTHREADS) {
    bf16* row = (r < 2 * BQ) ? (sQ + r * LDS) : (sK + (r - 2 * BQ) * LDS);
    *reinterpret_cast<uint4*>(row + HD) = make_uint4(0, 0, 0, 0);
  }

  // Q tile: rows beyond Tq are zero-filled
  auto load_q = [&](int qt, int qbuf) {
    const int q0 = qt * BQ;
    bf16* dst = sQ + qbuf * BQ * LDS;
    for (int i = tid; i < BQ * (HD / 8); i += NTHREADS) {
      const int r = i / (HD / 8), ch = i % (HD / 8);
      const bool ok = (q0 + r) < p.Tq;
      cp_async_16(dst + r * LDS + ch * 8, qg + (long)(ok ? q0 + r : 0) * p.ldq + ch * 8, ok);
    }
  };
  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * BKV;
    for (int i = tid; i < BKV * (HD / 8); i += NTHREADS) {
      const int r = i / (HD / 8), ch = i % (HD / 8);
      const bool ok = (k0 + r) < kv_n;
      const long gr = ok ? (k0 + r) : 0;
      cp_async_16(sK + (buf * BKV + r) * LDS + ch * 8, kg + gr * p.ldk + ch * 8, ok);
      cp_async_16(sV + (buf * BKV + r) * LDS + ch * 8, vg + gr * p.ldv + ch * 8, ok);
    }
  };
  // One cp.async group per step: the prologue group holds Q(0) and K/V(0) (and K/V(1) when resident); step (qi, t)
  // commits the group of the data the NEXT steps need and waits for everything older.
  load_q(qt0, 0);
  if (n_tiles > 0) load_kv(0, 0);
  if (resident && n_tiles > 1) load_kv(1, 1);
  cp_async_commit();

  int kvc = 0;   // K/V tiles consumed so far: tile buffers alternate across query tiles when K/V is streamed
  for (int qi = 0; qi < nq; ++qi) {
    const int q0 = (qt0 + qi) * BQ;
    float o_acc[NT_O][4];
#pragma unroll
    for (int i = 0; i < NT_O; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) o_acc[i][j] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    uint32_t qf[HDK / 16][4];

    for (int t = 0; t < n_tiles; ++t, ++kvc) {
      const int buf = resident ? t : (kvc & 1);
      if (t == 0 && qi + 1 < nq) load_q(qt0 + qi + 1, (qi + 1) & 1);
      if (!resident) {
        if (t + 1 < n_tiles)
          load_kv(t + 1, buf ^ 1);
        else if (qi + 1 < nq)
          load_kv(0, buf ^ 1);
      }
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();

      if (t == 0) {
        // Q fragments stay in registers for the whole KV sweep
        const bf16* sq = sQ + (qi & 1) * BQ * LDS;
        const int r = warp * 16 + (lane & 15);
        const int cbase = (lane >> 4) * 8;
#pragma unroll
        for (int kk = 0; kk < HDK / 16; ++kk)
          ldsm_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], smem_u32(sq + r * LDS + kk * 16 + cbase));
      }

      // ---- S = Q K^T
      float s[NT_S][4];
#pragma unroll
      for (int i = 0; i < NT_S; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
      const bf16* kb = sK + buf * BKV * LDS;
#pragma unroll
      for (int kk = 0; kk < HDK / 16; ++kk) {
#pragma unroll
        for (int np = 0; np < NT_S / 2; ++np) {
          // matrices: (keys 16np+0..7, d lo), (keys 0..7, d hi), (keys 8..15, d lo), (keys 8..15, d hi)
          const int kr = np * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int kc = kk * 16 + ((lane >> 3) & 1) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4(b0, b1, b2, b3, smem_u32(kb + kr * LDS + kc));
          mma_bf16(s[2 * np], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
          mma_bf16(s[2 * np + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
        }
      }

      // ---- mask the tail keys, online softmax
      const int kbase = t * BKV + (lane & 3) * 2;
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < NT_S; ++i) {
        const int key = kbase + i * 8;
        if (key >= kv_n) s[i][0] = s[i][2] = -INFINITY;
        if (key + 1 >= kv_n) s[i][1] = s[i][3] = -INFINITY;
        mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
      }
      float alpha[2], msc[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        const float m_new = fmaxf(m_run[h], mx[h]);
        // every tile holds at least one valid key, so m_new is finite
        alpha[h] = exp2f((m_run[h] - m_new) * p.scale_log2e);
        m_run[h] = m_new;
        msc[h] = m_new * p.scale_log2e;
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int i = 0; i < NT_S; ++i) {
        s[i][0] = exp2f(s[i][0] * p.scale_log2e - msc[0]);
        s[i][1] = exp2f(s[i][1] * p.scale_log2e - msc[0]);
        s[i][2] = exp2f(s[i][2] * p.scale_log2e - msc[1]);
        s[i][3] = exp2f(s[i][3] * p.scale_log2e - msc[1]);
        rs[0] += s[i][0] + s[i][1];
        rs[1] += s[i][2] + s[i][3];
      }
      l_run[0] = l_run[0] * alpha[0] + rs[0];
      l_run[1] = l_run[1] * alpha[1] + rs[1];
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
        o_acc[i][0] *= alpha[0];
        o_acc[i][1] *= alpha[0];
        o_acc[i][2] *= alpha[1];
        o_acc[i][3] *= alpha[1];
      }

      // ---- O += P V
      const bf16* vb = sV + buf * BKV * LDS;
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk) {
        const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        const uint32_t a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        // matrices (transposed on load): (keys lo, d tile n), (keys hi, d tile n), (keys lo, d tile n+1), (keys hi, n+1)
        const int vr = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int np = 0; np < NT_O / 2; ++np) {
          const int vc = np * 16 + (lane >> 4) * 8;
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(b0, b1, b2, b3, smem_u32(vb + vr * LDS + vc));
          mma_bf16(o_acc[2 * np], a0, a1, a2, a3, b0, b1);
          mma_bf16(o_acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
        }
        if (NT_O & 1) {
          uint32_t b0, b1;
          ldsm_x2_t(b0, b1, smem_u32(vb + vr * LDS + (NT_O - 1) * 8));
          mma_bf16(o_acc[NT_O - 1], a0, a1, a2, a3, b0, b1);
        }
      }
      __syncthreads();  // all warps are done with this step's buffers before they are refilled
    }

    // ---- finalise: O / l, bf16, (row, head*72 + col)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float l = l_run[h];
      l += __shfl_xor_sync(0xffffffffu, l, 1);
      l += __shfl_xor_sync(0xffffffffu, l, 2);
      const float inv = l > 0.f ? 1.0f / l : 0.f;
      const int row = q0 + warp * 16 + (lane >> 2) + h * 8;
      if (row < p.Tq) {
        bf16* og = p.out + ((long)b * p.Tq + row) * p.ldo + head * HD + (lane & 3) * 2;
#pragma unroll
        for (int i = 0; i < NT_O; ++i) {
          const uint32_t u = pack_bf16x2(o_acc[i][2 * h] * inv, o_acc[i][2 * h + 1] * inv);
          *reinterpret_cast<uint32_t*>(og + i * 8) = u;
        }
      }
    }
  }
  cp_async_wait<0>();
}

template <int NWARPS>
static int launch_attn(AttnDev p, int B, int heads, double flops, cudaStream_t stream) {
  constexpr int BQ = NWARPS * 16;
  constexpr int smem = (2 * BQ + 4 * BKV) * LDS * 2;
  auto kern = flash_attn_kernel<NWARPS>;
  IR_TRY(ensure_smem_optin((const void*)kern, smem));
  // One resident wave when possible: the kernel is latency-bound (load -> compute -> store per query tile), so a CTA
  // walks QT consecutive query tiles of its (sample, head) with the next Q tile prefetched under the current one
  // instead of leaving the tail of the grid to a second, mostly empty wave.
  const int n_qtiles = (p.Tq + BQ - 1) / BQ;
  const long total = (long)n_qtiles * heads * B;
  const long slots = (long)device_num_sms() * (NWARPS == 8 ? 2 : 3);   // resident CTAs (registers / shared memory)
  static const int forced_qt = [] { const char* e = debug_env("IR_XATTN_QT"); return e ? atoi(e) : 0; }();
  int qt = (int)((total + slots - 1) / slots);
  qt = qt < 1 ? 1 : (qt > 4 ? 4 : qt);
  if (forced_qt > 0) qt = forced_qt;
  p.QT = qt;
  dim3 grid((n_qtiles + qt - 1) / qt, heads, B);
  const bool prof = prof_enabled();
  if (prof) prof_before(stream);
  IR_CUDA_CHECK(launch_pdl(kern, grid, dim3(NWARPS * 32), smem, stream, p));
  if (prof) prof_after(stream, PROF_XATTN, flops);
  IR_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return IR_OK;
}

int attention_launch(const AttnArgs& a, cudaStream_t stream) {
  IR_REQUIRE(a.head_dim == HD, "attention: head_dim %d unsupported (kernel is specialised for %d)", a.head_dim, HD);
  IR_REQUIRE(a.q && a.k && a.v && a.out, "attention: null pointer");
  IR_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 2 == 0, "attention: strides must keep 16 B rows");
  IR_REQUIRE(a.B > 0 && a.heads > 0 && a.Tq > 0, "attention: bad shape");
  AttnDev p;
  p.q = a.q;
  p.k = a.k;
  p.v = a.v;
  p.out = a.out;
  p.ldq = a.ldq;
  p.ldk = a.ldk;
  p.ldv = a.ldv;
  p.ldo = a.ldo;
  p.Tq = a.Tq;
  p.Tk = a.Tk;
  p.kv_off = a.kv_off;
  p.kv_len = a.kv_len;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  // small problems: 64-row CTAs fill the 148 SMs better
  const long ctas128 = (long)((a.Tq + 127) / 128) * a.heads * a.B;
  // algorithmic FLOPs: 4 * Tq * Tk * d per (sample, head); Tk is not known on the host for ragged captions (counted 0)
  const double flops = a.kv_len ? 0.0 : 4.0 * a.B * a.heads * (double)a.Tq * a.Tk * a.head_dim;
  if (ctas128 >= 2 * 148) return launch_attn<8>(p, a.B, a.heads, flops, stream);
  return launch_attn<4>(p, a.B, a.heads, flops, stream);
}

}  // namespace ir
