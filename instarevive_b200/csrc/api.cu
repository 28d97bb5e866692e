// C ABI (include/instarevive_b200.h) over the C++ launchers. No torch types cross this boundary.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "../../include/instarevive_b200.h"
#include "dit.cuh"
#include "tiles.cuh"
#include "swinir.cuh"
#include "vae.cuh"

namespace ir {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count_value() { return g_launches.load(std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] {
    const char* e = debug_env("IR_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

int ensure_smem_optin(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  IR_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({dev, kernel})) return IR_OK;
  IR_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.insert({dev, kernel});
  return IR_OK;
}

int device_num_sms() {
  static std::atomic<int> sms[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  int n = sms[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

struct ProfRec {
  cudaEvent_t a, b;
  int klass;
  double flops;
  int M, N, K;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_pending = nullptr;

bool prof_enabled() { return g_prof_on; }
void prof_before(cudaStream_t s) {
  cudaEventCreate(&g_prof_pending);
  cudaEventRecord(g_prof_pending, s);
}
void prof_after(cudaStream_t s, int klass, double flops, int M, int N, int K) {
  ProfRec r;
  r.a = g_prof_pending;
  cudaEventCreate(&r.b);
  cudaEventRecord(r.b, s);
  r.klass = klass;
  r.flops = flops;
  r.M = M;
  r.N = N;
  r.K = K;
  g_prof.push_back(r);
}

}  // namespace ir

using namespace ir;

struct ir_dit {
  Dit* d;
};
struct ir_vae {
  Vae* v;
};

extern "C" {

const char* ir_last_error(void) { return g_err; }
const char* ir_version(void) { return "instarevive_b200 0.1 (sm_100a)"; }
long long ir_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void ir_profile_begin(void) {
  g_prof.clear();
  g_prof_on = true;
}

long long ir_profile_records(int* klass, int* M, int* N, int* K, float* ms, long long cap) {
  // per-launch records of the current profile pass in launch order (call between begin and end, after the work was
  // enqueued); synchronises the device. Returns the number of records (may exceed cap: only cap are written).
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  long long i = 0;
  for (const ProfRec& r : g_prof) {
    if (i < cap) {
      float t = 0.f;
      cudaEventElapsedTime(&t, r.a, r.b);
      if (klass) klass[i] = r.klass;
      if (M) M[i] = r.M;
      if (N) N[i] = r.N;
      if (K) K[i] = r.K;
      if (ms) ms[i] = t;
    }
    ++i;
  }
  return i;
}

int ir_profile_end(double* ms_by_class, double* flops_by_class, long long* launches_by_class) {
  g_prof_on = false;
  cudaError_t e = cudaDeviceSynchronize();
  for (int i = 0; i < PROF_NUM; ++i) {
    ms_by_class[i] = 0.0;
    flops_by_class[i] = 0.0;
    launches_by_class[i] = 0;
  }
  for (ProfRec& r : g_prof) {
    float ms = 0.f;
    if (e == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.klass >= 0 && r.klass < PROF_NUM) {
      ms_by_class[r.klass] += ms;
      flops_by_class[r.klass] += r.flops;
      launches_by_class[r.klass] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
  if (e != cudaSuccess) {
    set_last_error("ir_profile_end: %s", cudaGetErrorString(e));
    return IR_ERR_CUDA;
  }
  return IR_OK;
}

__global__ void profile_empty_kernel() {}

int ir_profile_calibrate(int n, void* stream) {
  // The floor of the event-pair measurement: n empty one-warp kernels, each between its own pair of events exactly like a
  // profiled launch. What ir_profile_end then reports for class 4 (ms / launches) is what an event pair adds to a kernel's
  // duration -- the launch no longer hides behind its predecessor (no programmatic overlap across an event) plus the two
  // event records -- measured on this device in this process.
  if (!g_prof_on || n <= 0) {
    set_last_error("ir_profile_calibrate: call between ir_profile_begin and ir_profile_end with n > 0");
    return IR_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < n; ++i) {
    prof_before(s);
    profile_empty_kernel<<<1, 32, 0, s>>>();
    prof_after(s, PROF_CAL, 0.0);
  }
  IR_CUDA_CHECK(cudaGetLastError());
  return IR_OK;
}

int ir_dit_create(const ir_dit_config* cfg, ir_dit** out) {
  if (!cfg || !out) {
    set_last_error("ir_dit_create: null argument");
    return IR_ERR_INVALID;
  }
  DitConfig c;
  c.depth = cfg->depth;
  c.copy_blocks = cfg->copy_blocks;
  c.hidden = cfg->hidden;
  c.heads = cfg->heads;
  c.patch = cfg->patch;
  c.in_ch = cfg->in_channels;
  c.out_ch = cfg->out_channels;
  c.caption_ch = cfg->caption_channels;
  c.mlp_ratio = cfg->mlp_ratio;
  c.base_size = cfg->base_size;
  c.pe_interpolation = cfg->pe_interpolation;
  Dit* d = nullptr;
  int st = dit_create(c, &d);
  if (st != IR_OK) return st;
  *out = new ir_dit{d};
  return IR_OK;
}

void ir_dit_destroy(ir_dit* h) {
  if (!h) return;
  dit_destroy(h->d);
  delete h;
}

int ir_dit_num_params(const ir_dit* h) { return h ? (int)h->d->params.size() : 0; }

int ir_dit_param_info(const ir_dit* h, int i, char* name, int name_cap, long long* numel, int* rows, int* cols) {
  if (!h || i < 0 || i >= (int)h->d->params.size()) {
    set_last_error("ir_dit_param_info: index %d out of range", i);
    return IR_ERR_INVALID;
  }
  const ParamEntry& e = h->d->params[i];
  if (name && name_cap > 0) {
    strncpy(name, e.name.c_str(), (size_t)name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (numel) *numel = e.numel;
  if (rows) *rows = e.rows;
  if (cols) *cols = e.cols;
  return IR_OK;
}

int ir_dit_load_param(ir_dit* h, const char* name, const float* src_dev, long long numel, void* stream) {
  if (!h || !name || !src_dev) {
    set_last_error("ir_dit_load_param: null argument");
    return IR_ERR_INVALID;
  }
  return dit_load_param(h->d, name, src_dev, (long)numel, (cudaStream_t)stream);
}

size_t ir_dit_workspace_bytes(const ir_dit* h, int B, int H, int W, int sum_l) {
  return h ? dit_workspace_bytes(h->d, B, H, W, sum_l) : 0;
}

int ir_dit_set_graphs(ir_dit* h, int enable) {
  if (!h) {
    set_last_error("ir_dit_set_graphs: null handle");
    return IR_ERR_INVALID;
  }
  dit_set_graphs(h->d, enable != 0);
  return IR_OK;
}

int ir_dit_set_dual_chain(ir_dit* h, int enable) {
  if (!h) {
    set_last_error("ir_dit_set_dual_chain: null handle");
    return IR_ERR_INVALID;
  }
  dit_set_dual_chain(h->d, enable != 0);
  return IR_OK;
}

int ir_dit_reserve(ir_dit* h, int max_tokens, int max_sum_l) {
  if (!h) {
    set_last_error("ir_dit_reserve: null handle");
    return IR_ERR_INVALID;
  }
  return dit_reserve(h->d, max_tokens, max_sum_l);
}

int ir_dit_forward(ir_dit* h, const float* x, const float* c, const float* timestep, const float* y,
                   const int32_t* y_index, const int32_t* kv_off, const int32_t* kv_len, const float* img_hw,
                   const float* aspect, float* out, int B, int H, int W, int sum_l, int max_l, long long kv_total,
                   int reuse_caption, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) {
    set_last_error("ir_dit_forward: null handle");
    return IR_ERR_INVALID;
  }
  DitForwardArgs a;
  a.x = x;
  a.c = c;
  a.timestep = timestep;
  a.y = y;
  a.y_index = y_index;
  a.kv_off = kv_off;
  a.kv_len = kv_len;
  a.img_hw = img_hw;
  a.aspect = aspect;
  a.out = out;
  a.B = B;
  a.H = H;
  a.W = W;
  a.sumL = sum_l;
  a.max_len = max_l;
  a.kv_total = (long)kv_total;
  a.reuse_caption = reuse_caption;
  a.workspace = workspace;
  a.workspace_bytes = workspace_bytes;
  return dit_forward(h->d, a, (cudaStream_t)stream);
}

int ir_dit_patch_embed(ir_dit* h, const float* x, float* tokens, int B, int H, int W, void* stream) {
  if (!h || !x || !tokens) {
    set_last_error("ir_dit_patch_embed: null argument");
    return IR_ERR_INVALID;
  }
  return dit_patch_embed(h->d, x, tokens, B, H, W, (cudaStream_t)stream);
}

int ir_eps_to_x0(const float* x, const float* model_out, float* x0, int B, int C, int HW, float sqrt_abar,
                 float sqrt_one_minus_abar, void* stream) {
  return eps_to_x0_launch(x, model_out, x0, B, C, HW, sqrt_abar, sqrt_one_minus_abar, (cudaStream_t)stream);
}

int ir_lincomb3(const float* x, const float* m0, const float* m1, float* out, long long n, float ca, float c0, float c1,
                void* stream) {
  return lincomb3_launch(x, m0, m1, out, n, ca, c0, c1, (cudaStream_t)stream);
}

int ir_vae_create(const ir_vae_config* cfg, ir_vae** out) {
  if (!cfg || !out) {
    set_last_error("ir_vae_create: null argument");
    return IR_ERR_INVALID;
  }
  VaeConfig c;
  c.ch = cfg->ch;
  c.z_channels = cfg->z_channels;
  c.out_ch = cfg->out_ch;
  c.num_res_blocks = cfg->num_res_blocks;
  for (int i = 0; i < 4; ++i) c.ch_mult[i] = cfg->ch_mult[i];
  c.with_encoder = cfg->with_encoder;
  Vae* v = nullptr;
  int st = vae_create(c, &v);
  if (st != IR_OK) return st;
  *out = new ir_vae{v};
  return IR_OK;
}

void ir_vae_destroy(ir_vae* h) {
  if (!h) return;
  vae_destroy(h->v);
  delete h;
}

int ir_vae_num_params(const ir_vae* h) { return h ? (int)h->v->params.size() : 0; }

int ir_vae_param_info(const ir_vae* h, int i, char* name, int name_cap, long long* numel) {
  if (!h || i < 0 || i >= (int)h->v->params.size()) {
    set_last_error("ir_vae_param_info: index %d out of range", i);
    return IR_ERR_INVALID;
  }
  const VaeParam& e = h->v->params[i];
  if (name && name_cap > 0) {
    strncpy(name, e.name.c_str(), (size_t)name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (numel) *numel = e.numel;
  return IR_OK;
}

int ir_vae_load_param(ir_vae* h, const char* name, const float* src_dev, long long numel, void* stream) {
  if (!h || !name || !src_dev) {
    set_last_error("ir_vae_load_param: null argument");
    return IR_ERR_INVALID;
  }
  return vae_load_param(h->v, name, src_dev, (long)numel, (cudaStream_t)stream);
}

int ir_vae_set_graphs(ir_vae* h, int enable) {
  if (!h) {
    set_last_error("ir_vae_set_graphs: null handle");
    return IR_ERR_INVALID;
  }
  vae_set_graphs(h->v, enable != 0);
  return IR_OK;
}

size_t ir_vae_workspace_bytes(const ir_vae* h, int B, int h_lat, int w_lat) {
  return h ? vae_workspace_bytes(h->v, B, h_lat, w_lat) : 0;
}

int ir_vae_decode(ir_vae* h, const float* z, float* out, int B, int h_lat, int w_lat, float in_scale, float out_scale,
                  float out_shift, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h) {
    set_last_error("ir_vae_decode: null handle");
    return IR_ERR_INVALID;
  }
  return vae_decode(h->v, z, out, B, h_lat, w_lat, in_scale, out_scale, out_shift, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

size_t ir_vae_encode_workspace_bytes(const ir_vae* h, int B, int H, int W) {
  return h ? vae_encode_workspace_bytes(h->v, B, H, W) : 0;
}

int ir_vae_encode(ir_vae* h, const float* x, float* moments, int B, int H, int W, void* workspace, size_t workspace_bytes,
                  void* stream) {
  if (!h) {
    set_last_error("ir_vae_encode: null handle");
    return IR_ERR_INVALID;
  }
  return vae_encode(h->v, x, moments, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ SwinIR stage 1
struct ir_swinir {
  Swin* s;
};

int ir_swinir_create(ir_swinir** out) {
  if (!out) {
    set_last_error("ir_swinir_create: null argument");
    return IR_ERR_INVALID;
  }
  Swin* s = nullptr;
  int st = swin_create(SwinConfig{}, &s);
  if (st != IR_OK) return st;
  *out = new ir_swinir{s};
  return IR_OK;
}
void ir_swinir_destroy(ir_swinir* h) {
  if (!h) return;
  swin_destroy(h->s);
  delete h;
}
int ir_swinir_num_params(const ir_swinir* h) { return h ? (int)h->s->params.size() : 0; }
int ir_swinir_param_info(const ir_swinir* h, int i, char* name, int name_cap, long long* numel) {
  if (!h || i < 0 || i >= (int)h->s->params.size() || !name || name_cap <= 0) {
    set_last_error("ir_swinir_param_info: bad argument");
    return IR_ERR_INVALID;
  }
  const SwinParam& p = h->s->params[i];
  snprintf(name, (size_t)name_cap, "%s", p.name.c_str());
  if (numel) *numel = p.numel;
  return IR_OK;
}
int ir_swinir_load_param(ir_swinir* h, const char* name, const float* src_dev, long long numel, void* stream) {
  if (!h || !name || !src_dev) {
    set_last_error("ir_swinir_load_param: null argument");
    return IR_ERR_INVALID;
  }
  return swin_load_param(h->s, name, src_dev, numel, (cudaStream_t)stream);
}
size_t ir_swinir_workspace_bytes(const ir_swinir* h, int B, int H, int W) { return h ? swin_workspace_bytes(h->s, B, H, W) : 0; }
int ir_swinir_forward(ir_swinir* h, const float* x, float* out, int B, int H, int W, void* workspace, size_t workspace_bytes,
                      void* stream) {
  if (!h) {
    set_last_error("ir_swinir_forward: null handle");
    return IR_ERR_INVALID;
  }
  return swin_forward(h->s, x, out, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int ir_tile_gather(const float* src, float* dst, const int32_t* coords, int ntiles, int N, int C, int H, int W, int th,
                   int tw, int scale, void* stream) {
  return tile_gather_launch(src, dst, coords, ntiles, N, C, H, W, th, tw, scale, (cudaStream_t)stream);
}
int ir_tile_blend(const float* tiles, const int32_t* coords, int ntiles, float* out, int N, int C, int H, int W, int th,
                  int tw, int scale, void* stream) {
  return tile_blend_launch(tiles, coords, ntiles, out, N, C, H, W, th, tw, scale, (cudaStream_t)stream);
}
size_t ir_wavelet_workspace_bytes(int N, int C, int H, int W) { return wavelet_workspace_bytes(N, C, H, W); }
int ir_wavelet_reconstruction(const float* content, const float* style, float* out, int N, int C, int H, int W,
                              void* workspace, size_t workspace_bytes, void* stream) {
  return wavelet_reconstruction_launch(content, style, out, N, C, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}
int ir_adain(const float* content, const float* style, float* out, int N, int C, int HW, void* stream) {
  return adain_launch(content, style, out, N, C, HW, (cudaStream_t)stream);
}
int ir_to_uint8(const float* img, uint8_t* out, int N, int C, int H, int W, void* stream) {
  return to_uint8_launch(img, out, N, C, H, W, (cudaStream_t)stream);
}

int ir_gemm_bf16(const void* A, const void* W, const float* bias, int M, int N, int K, int batch, long long strideA,
                 long long strideW, long long strideO, int epilogue, float alpha, void* out_bf16, float* out_f32,
                 const float* resid_f32, const float* gate, long long gate_ld, int rows_per_gate, int force_bn,
                 void* stream) {
  GemmArgs g;
  g.A = (const bf16*)A;
  g.lda = K;
  g.strideA = strideA;
  g.W = (const bf16*)W;
  g.ldw = K;
  g.strideW = strideW;
  g.M = M;
  g.N = N;
  g.K = K;
  g.batch = batch;
  g.epi = epilogue;
  g.alpha = alpha;
  g.bias = bias;
  g.out_bf16 = (bf16*)out_bf16;
  g.ldo_b = N;
  g.stride_ob = strideO;
  g.out_f32 = out_f32;
  g.resid_f32 = resid_f32;
  g.ldo_f = N;
  g.stride_of = strideO;
  g.gate = gate;
  g.gate_ld = gate_ld;
  g.rows_per_gate = rows_per_gate;
  g.force_bn = force_bn;
  return gemm_launch(g, (cudaStream_t)stream);
}

int ir_conv3x3_bf16(const void* act, const void* weight, const float* bias, int n, int H, int W, int C, int Cout,
                    void* out_bf16, float* out_f32, const void* resid_bf16, const float* resid_f32, int force_bn,
                    void* stream) {
  GemmArgs g;
  g.A = (const bf16*)act;
  g.W = (const bf16*)weight;
  g.ldw = 9L * C;
  g.M = n * H * W;
  g.N = Cout;
  g.K = 9 * C;
  g.conv = 1;
  g.nimg = n;
  g.H = H;
  g.Wd = W;
  g.C = C;
  g.bias = bias;
  g.force_bn = force_bn;
  if (out_f32) {
    g.epi = EPI_F32;
    g.out_f32 = out_f32;
    g.resid_f32 = resid_f32;
    g.ldo_f = Cout;
    g.out_bf16 = (bf16*)out_bf16;
    g.ldo_b = Cout;
  } else {
    g.epi = EPI_BF16;
    g.out_bf16 = (bf16*)out_bf16;
    g.resid_bf16 = (const bf16*)resid_bf16;
    g.ldo_b = Cout;
  }
  return gemm_launch(g, (cudaStream_t)stream);
}

int ir_conv3x3_s2_bf16(const void* act, const void* weight, const float* bias, int n, int Ho, int Wo, int C, int Cout,
                       void* out_bf16, int force_bn, void* stream) {
  GemmArgs g;
  g.A = (const bf16*)act;
  g.W = (const bf16*)weight;
  g.ldw = 9L * C;
  g.M = n * Ho * Wo;
  g.N = Cout;
  g.K = 9 * C;
  g.conv = 1;
  g.conv_stride = 2;
  g.nimg = n;
  g.H = Ho;
  g.Wd = Wo;
  g.C = C;
  g.bias = bias;
  g.force_bn = force_bn;
  g.epi = EPI_BF16;
  g.out_bf16 = (bf16*)out_bf16;
  g.ldo_b = Cout;
  return gemm_launch(g, (cudaStream_t)stream);
}

int ir_conv1x1_bf16(const void* act, const void* weight, const float* bias, int n, int H, int W, int C, int Cout,
                    void* out_bf16, const void* resid_bf16, int force_bn, void* stream) {
  GemmArgs g;
  g.A = (const bf16*)act;
  g.W = (const bf16*)weight;
  g.ldw = C;
  g.M = n * H * W;
  g.N = Cout;
  g.K = C;
  g.conv = 1;
  g.conv_taps = 1;
  g.nimg = n;
  g.H = H;
  g.Wd = W;
  g.C = C;
  g.bias = bias;
  g.force_bn = force_bn;
  g.epi = EPI_BF16;
  g.out_bf16 = (bf16*)out_bf16;
  g.resid_bf16 = (const bf16*)resid_bf16;
  g.ldo_b = Cout;
  return gemm_launch(g, (cudaStream_t)stream);
}

int ir_gemm_attn_pass(const void* A, const void* W, int M, int N, int K, long long lda, long long ldw, int mode, float alpha,
                      const float* att_row, float* att_out, void* out_bf16, long long ldo, int force_bn, void* stream) {
  GemmArgs g;
  g.A = (const bf16*)A;
  g.lda = lda;
  g.W = (const bf16*)W;
  g.ldw = ldw;
  g.M = M;
  g.N = N;
  g.K = K;
  g.epi = EPI_ATTN;
  g.att_mode = mode;
  g.alpha = alpha;
  g.att_row = att_row;
  g.att_out = att_out;
  g.out_bf16 = (bf16*)out_bf16;
  g.ldo_b = ldo;
  g.force_bn = force_bn;
  return gemm_launch(g, (cudaStream_t)stream);
}

int ir_upsample_conv3x3_bf16(const void* act, const float* weight_oihw, const float* bias, int n, int H, int W, int C,
                             void* phase_w_ws, void* out_bf16, int force_bn, void* stream) {
  if (!act || !weight_oihw || !phase_w_ws || !out_bf16) {
    set_last_error("ir_upsample_conv3x3_bf16: null pointer");
    return IR_ERR_INVALID;
  }
  IR_TRY(pack_upconv_phases(weight_oihw, (bf16*)phase_w_ws, C, C, (cudaStream_t)stream));
  return upsample_conv_phases_launch((const bf16*)act, (const bf16*)phase_w_ws, bias, (bf16*)out_bf16, n, H, W, C, nullptr,
                                     force_bn, (cudaStream_t)stream);
}

int ir_attention_bf16(const void* q, const void* k, const void* v, void* out, long long ldq, long long ldk,
                      long long ldv, long long ldo, int B, int heads, int head_dim, int Tq, int Tk,
                      const int32_t* kv_off, const int32_t* kv_len, float scale, void* stream) {
  AttnArgs a;
  a.q = (const bf16*)q;
  a.k = (const bf16*)k;
  a.v = (const bf16*)v;
  a.out = (bf16*)out;
  a.ldq = ldq;
  a.ldk = ldk;
  a.ldv = ldv;
  a.ldo = ldo;
  a.B = B;
  a.heads = heads;
  a.head_dim = head_dim;
  a.Tq = Tq;
  a.Tk = Tk;
  a.kv_off = kv_off;
  a.kv_len = kv_len;
  a.scale = scale;
  return attention_launch(a, (cudaStream_t)stream);
}

size_t ir_cross_attention_vt_bytes(int heads, int sum_l) { return (size_t)xattention_vt_elems(heads, sum_l) * sizeof(bf16); }

int ir_cross_attention_tc_bf16(const void* q, const void* kv, void* vt_ws, void* out, long long ldq, long long ldkv,
                               long long ldo, int B, int heads, int head_dim, int T, int sum_l, const int32_t* kv_off,
                               const int32_t* kv_len, int max_l, float scale, void* stream) {
  if (!kv || !vt_ws) {
    set_last_error("ir_cross_attention_tc_bf16: null kv / vt workspace");
    return IR_ERR_INVALID;
  }
  IR_TRY(xattention_transpose_v((const bf16*)kv, (bf16*)vt_ws, 1, heads, sum_l, ldkv, (cudaStream_t)stream));
  XAttnTcArgs a;
  a.vt = (const bf16*)vt_ws;
  a.q = (const bf16*)q;
  a.kv = (const bf16*)kv;
  a.out = (bf16*)out;
  a.ldq = ldq;
  a.ldkv = ldkv;
  a.ldo = ldo;
  a.B = B;
  a.H = heads;
  a.head_dim = head_dim;
  a.T = T;
  a.sumL = sum_l;
  a.kv_off = kv_off;
  a.kv_len = kv_len;
  a.max_len = max_l;
  a.kv_total = (long)B * max_l;
  a.scale = scale;
  return xattention_tc_launch(a, (cudaStream_t)stream);
}

int ir_gemm_qkv_heads(const void* A, const void* W, const float* bias, int M, int K, int T, int Tp, int H, int hd,
                      void* q_heads, void* k_heads, void* vt_heads, int force_bn, void* stream) {
  GemmArgs g;
  g.A = (const bf16*)A;
  g.lda = K;
  g.W = (const bf16*)W;
  g.ldw = K;
  g.M = M;
  g.N = 3 * H * hd;
  g.K = K;
  g.epi = EPI_QKV;
  g.bias = bias;
  g.q_heads = (bf16*)q_heads;
  g.k_heads = (bf16*)k_heads;
  g.vt_heads = (bf16*)vt_heads;
  g.qkv_T = T;
  g.qkv_Tp = Tp;
  g.qkv_H = H;
  g.qkv_hd = hd;
  g.force_bn = force_bn;
  return gemm_launch(g, (cudaStream_t)stream);
}

int ir_attention_tc_bf16(const void* q_heads, const void* k_heads, const void* vt_heads, void* out, long long ldo, int B,
                         int H, int head_dim, int T, int Tp, float scale, void* stream) {
  AttnTcArgs a;
  a.q = (const bf16*)q_heads;
  a.k = (const bf16*)k_heads;
  a.vt = (const bf16*)vt_heads;
  a.out = (bf16*)out;
  a.ldo = ldo;
  a.B = B;
  a.H = H;
  a.head_dim = head_dim;
  a.T = T;
  a.Tp = Tp;
  a.scale = scale;
  return attention_tc_launch(a, (cudaStream_t)stream);
}

int ir_debug_gemm_trace(long long* device_buf, int slots) {
  gemm_set_trace(device_buf, slots);
#ifdef IR_DEBUG
  return 16;
#else
  return 0;   // the release library carries no trace code
#endif
}

int ir_debug_attention_trace(long long* device_buf) {
  attention_tc_set_trace(device_buf);
  return attention_tc_trace_len();
}

int ir_ln_modulate(const float* x, void* out_bf16, const float* shift, const float* scale, long long mod_stride,
                   int rows, int T, int D, void* stream) {
  return ln_modulate_launch(x, (bf16*)out_bf16, shift, scale, mod_stride, rows, T, D, (cudaStream_t)stream);
}

int ir_pos_embed(float* table, int gh, int gw, int D, int base_size, float pe_interpolation, void* stream) {
  return pos_embed_launch(table, gh, gw, D, base_size, pe_interpolation, (cudaStream_t)stream);
}

}  // extern "C"
