// Unet3D velocity-field engine + the C ABI (include/ftb.h).
//
// The engine mirrors the structure of the reference ctor/forward
// (src/flowtrain/models/unet_attn_3d.py:509-667, :673-719) as a static sequence of kernel
// launches over blocked bf16 activations carved out of a caller-provided workspace.
// Parameters keep the reference state_dict names; they are copied in as fp32 and repacked
// (conv weights -> bf16 UMMA tiles, RMSNorm gains pre-multiplied by sqrt(C), pre-attention norm
// folded into to_qkv) lazily before the first forward after a change.
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ftb.h"
#include "ops.h"

namespace ftb {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(const char* file, int line, const std::string& msg) {
  const char* base = file;
  for (const char* c = file; *c; ++c)
    if (*c == '/') base = c + 1;
  g_err = std::string(base) + ":" + std::to_string(line) + ": " + msg;
  return -1;
}
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// ---- launch counter and optional per-launch event timing of the conv kernel
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }

namespace {
struct ProfRec {
  cudaEvent_t e0, e1;
  double flops, bytes;
  int kind;
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t pool_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace
bool prof_enabled() { return g_prof_on; }
int prof_begin(cudaStream_t st, double flops, double bytes, int kind) {
  if (!g_prof_on) return -1;
  ProfRec r{pool_event(), pool_event(), flops, bytes, kind};
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void prof_end(int idx, cudaStream_t st) {
  if (idx >= 0 && idx < (int)g_prof.size()) cudaEventRecord(g_prof[idx].e1, st);
}
void prof_set(bool on) { g_prof_on = on; }
// sums per kind (0: conv k>=3, 1: conv 1x1); returns number of records consumed
int prof_collect(double* flops, double* bytes, double* ms, int* launches, int nkinds) {
  for (int k = 0; k < nkinds; ++k) { flops[k] = 0; bytes[k] = 0; ms[k] = 0; launches[k] = 0; }
  int n = 0;
  for (ProfRec& r : g_prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess &&
        r.kind >= 0 && r.kind < nkinds) {
      flops[r.kind] += r.flops;
      bytes[r.kind] += r.bytes;
      ms[r.kind] += t;
      launches[r.kind] += 1;
      ++n;
    }
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
  return n;
}

namespace eng {

struct Param {
  std::string name;
  std::vector<int> shape;
  int64_t numel = 0;
  float* dev = nullptr;  // fp32 copy owned by the handle
  bool set = false;
};

struct ConvLayer {
  std::string wname, bname;  // bname empty = no bias
  int cout = 0, cin = 0, k = 1;
  int cin_pad = 0, n_tile = 0, ntiles = 1;
  bool unfold_w = false;     // stem convs: W taps unfolded into the channels (K extent k*cin instead of k x pad16(cin))
  std::string in_scale;      // name of a gain vector folded into the input channels ("" = none)
  float in_scale_mul = 1.f;
  bf16* packed = nullptr;
  float* bias = nullptr;     // padded to ntiles*n_tile
  float* scale_tmp = nullptr;
};

struct GainVec {
  std::string gname;
  int c = 0;
  float* gs = nullptr;  // g * sqrt(C)
};

}  // namespace eng
}  // namespace ftb

using namespace ftb;
using namespace ftb::eng;

namespace ftb_engine_detail {
struct DgradPack {
  int ci0 = 0, cin_sub = 0, n_tile = 0;
  ftb::bf16* packed = nullptr;
};
struct TrainState;
struct TrainCtx;
}  // namespace ftb_engine_detail

struct ftb_unet {
  ftb_unet_cfg cfg;
  std::vector<int> dims;
  std::vector<std::pair<int, int>> in_out;
  int time_dim = 0;
  std::vector<Param> params;
  std::map<std::string, int> pindex;
  std::map<std::string, ConvLayer> convs;
  std::map<std::string, GainVec> gains;
  struct KShift { float* dev = nullptr; bool ok = false; bool q_ok = false; };
  std::map<std::string, KShift> kshift;   // LinearAttention layers: softmax shift of the fused k/v-context path
  std::vector<std::string> film_blocks;  // FiLM-table rows: prefix of the (SiLU, Linear) time MLP's Linear
  std::map<std::string, std::string> film_gain;   // block -> gain vector folded into its scale half ("" = none)
  std::map<std::string, int> film_off;
  int film_rows = 0;
  const float** d_film_w = nullptr;
  const float** d_film_b = nullptr;
  const float** d_film_gs = nullptr;
  int* d_film_off = nullptr;
  float drop_p = 0.f;            // training: nn.Dropout probability of Block1 (0 = off)
  unsigned long long drop_seed = 0;
  bool dirty = true;
  bool kshift_stale = true;      // the fused-attention shift vectors lag the weights (training skips them)
  bool dgrad_dirty = true;       // transposed packs for the data gradients lag the weights
  ftb::PackJob* d_jobs = nullptr;        // device table: forward packs of every conv
  int n_jobs = 0;
  const float* jobs_base = nullptr;      // parameter storage the table was built for
  ftb::PackJob* d_djobs = nullptr;       // device table: transposed packs (data gradients)
  int n_djobs = 0, cap_djobs = 0;
  const float* djobs_base = nullptr;
  std::map<std::string, std::vector<ftb_engine_detail::DgradPack>> dgrad;
  std::map<std::string, std::pair<ftb::bf16*, ftb::bf16*>> f32packs;   // fp32 mode: (hi, lo) weight packs per conv
  ftb::PackJob* d_f32jobs = nullptr;
  int n_f32jobs = 0;
  const float* f32jobs_base = nullptr;
  bool f32_stale = true;
  std::map<std::string, std::vector<long long>> taps32;   // fp32 mode: name -> (pointer, B, C, D, H, W)
  std::shared_ptr<ftb_engine_detail::TrainState> train;
  std::shared_ptr<ftb_engine_detail::TrainCtx> train_ctx;
  bool on_device = false;
  std::map<std::string, Act> taps;
  bool keep_taps = true;
  int launches = 0;
  std::vector<void*> owned;

  ~ftb_unet() {
    for (void* p : owned) cudaFree(p);
  }
};

namespace ftb_engine_detail {

template <typename T>
int dev_alloc(ftb_unet* U, T** p, size_t n) {
  void* q = nullptr;
  FTB_CUDA(cudaMalloc(&q, n * sizeof(T) > 0 ? n * sizeof(T) : 16));
  U->owned.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

void add_param(ftb_unet* U, const std::string& name, std::vector<int> shape) {
  Param p;
  p.name = name;
  p.shape = shape;
  p.numel = 1;
  for (int s : shape) p.numel *= s;
  U->pindex[name] = (int)U->params.size();
  U->params.push_back(p);
}

void add_conv(ftb_unet* U, const std::string& prefix, int cout, int cin, int k, bool bias,
              const std::string& in_scale = "", int n_tile = 0) {
  add_param(U, prefix + ".weight", {cout, cin, k, k, k});
  if (bias) add_param(U, prefix + ".bias", {cout});
  ConvLayer c;
  c.wname = prefix + ".weight";
  c.bname = bias ? prefix + ".bias" : "";
  c.cout = cout; c.cin = cin; c.k = k;
  c.cin_pad = round_up(cin, 16);
  if (k >= 5 && k * cin <= 128 && round_up(k * cin, 16) < k * c.cin_pad && getenv("FTB_NO_UNFOLD") == nullptr) {
    c.unfold_w = true;
    c.cin_pad = round_up(k * cin, 16);
  }
  if (n_tile == 0) n_tile = round_up(cout, 16);
  c.n_tile = n_tile;
  c.ntiles = cdiv(cout, n_tile);
  c.in_scale = in_scale;
  c.in_scale_mul = sqrtf((float)cin);
  U->convs[prefix] = c;
}

void add_gain(ftb_unet* U, const std::string& gname, int c, bool as_param = true) {
  if (as_param) add_param(U, gname, {1, c, 1, 1, 1});
  GainVec g;
  g.gname = gname;
  g.c = c;
  U->gains[gname] = g;
}

// the reference names the block's time MLP `mlp` in unet_attn_3d.py:255 and `time_mlp` in
// unet_attn_3d_cond_v3.py:337
std::string resnet_mlp(const ftb_unet* U, const std::string& p) {
  return p + (U->cfg.conditional ? ".time_mlp.1" : ".mlp.1");
}

// N tile of a conv at the coarsest level (downs[n-1], mid, ups[0]: 4^3 voxels per sample at 64^3): there a launch has
// only 8-32 work items and every CTA streams the whole weight tensor through shared memory, so a wide layer
// (>= 128 output channels) is cut into <= 48-channel N tiles that run as blockIdx.y of the same launch; the channel
// norm then runs as its own pass (normact_fwd).  0 = one tile.
int deep_n_tile(int cout) {
  if (cout < 128 || getenv("FTB_NO_NSPLIT") != nullptr) return 0;
  for (int n = 48; n >= 16; n -= 16)
    if (cout % n == 0) return n;
  return 0;
}

void add_resnet(ftb_unet* U, const std::string& p, int cin, int cout, bool deep = false) {
  add_param(U, resnet_mlp(U, p) + ".weight", {2 * cout, U->time_dim});
  add_param(U, resnet_mlp(U, p) + ".bias", {2 * cout});
  add_conv(U, p + ".block1.proj", cout, cin, 3, true, "", deep ? deep_n_tile(cout) : 0);
  add_gain(U, p + ".block1.norm.g", cout);
  add_conv(U, p + ".block2.proj", cout, cout, 3, true, "", deep ? deep_n_tile(cout) : 0);
  add_gain(U, p + ".block2.norm.g", cout);
  if (cin != cout) add_conv(U, p + ".res_conv", cout, cin, 1, true);
}

// EmbedATb (unet_attn_3d_cond_v3.py:120-129) + MixATb (:156-173) of one stage
void add_embed_mix(ftb_unet* U, const std::string& p, int d, bool deep = false) {
  add_conv(U, p + ".0.conv1", d, U->cfg.data_channels, 5, true, "", deep ? deep_n_tile(d) : 0);
  add_conv(U, p + ".0.conv2", d, d, 5, true, "", deep ? deep_n_tile(d) : 0);
  add_param(U, p + ".1.time_mlp.1.weight", {4 * d, U->time_dim});
  add_param(U, p + ".1.time_mlp.1.bias", {4 * d});
  add_conv(U, p + ".1.conv1", d, 2 * d, 3, true);
  add_gain(U, p + ".1.norm.g", d);
  add_conv(U, p + ".1.conv2", d, d, 3, true, "", deep ? deep_n_tile(d) : 0);
}

void add_attn(ftb_unet* U, const std::string& p, int dim, bool full) {
  const int heads = U->cfg.attn_heads, dh = U->cfg.attn_dim_head, hd = heads * dh;
  const int nm = U->cfg.num_mem_kv;
  if (full) add_param(U, p + ".mem_kv", {2, heads, nm, dh});
  else add_param(U, p + ".mem_kv", {2, heads, dh, nm});
  add_param(U, p + ".norm.g", {1, dim, 1, 1, 1});
  // pre-norm gain * sqrt(C) is folded into the to_qkv input channels; 3 N tiles (q | k | v)
  add_conv(U, p + ".to_qkv", 3 * hd, dim, 1, false, p + ".norm.g", hd);
  if (full) {
    add_conv(U, p + ".to_out", dim, hd, 1, true);
  } else {
    // LinearAttention's out projection is folded into per-sample weights (attention.cu), so
    // its weight stays fp32; only the bias/gain are used by the conv epilogue.
    U->kshift[p] = ftb_unet::KShift{};
    add_conv(U, p + ".to_out.0", dim, hd, 1, true);   // packed copy is used by the training path only
    add_gain(U, p + ".to_out.1.g", dim);
  }
}

int build_plan(ftb_unet* U) {
  const ftb_unet_cfg& c = U->cfg;
  FTB_CHECK(c.n_stages >= 1 && c.n_stages <= FTB_MAX_STAGES, "n_stages out of range");
  FTB_CHECK(c.dim % 16 == 0 && c.dim >= 16, "dim must be a multiple of 16");
  FTB_CHECK(c.attn_dim_head == 16 || c.attn_dim_head == 32, "attn_dim_head must be 16 or 32");
  FTB_CHECK((c.attn_heads * c.attn_dim_head) % 16 == 0 && c.attn_heads * c.attn_dim_head <= 256,
            "heads*dim_head must be a multiple of 16, at most 256");
  FTB_CHECK(c.data_channels >= 1 && c.data_channels <= 256, "data_channels out of range");
  U->dims.clear();
  U->dims.push_back(c.dim);
  for (int i = 0; i < c.n_stages; ++i) {
    FTB_CHECK(c.dim_mults[i] >= 1 && c.dim * c.dim_mults[i] <= 256, "dim*mult must be in [16,256]");
    U->dims.push_back(c.dim * c.dim_mults[i]);
  }
  for (int i = 0; i < c.n_stages; ++i) U->in_out.push_back({U->dims[i], U->dims[i + 1]});
  U->time_dim = c.dim * 4;
  const int n = c.n_stages;

  const bool cond = c.conditional != 0;
  // module indices inside a stage: [resnet, resnet, attn, resample] (unet_attn_3d.py:600-611) or
  // [EmbedATb, MixATb, resnet, resnet, attn, resample] (unet_attn_3d_cond_v3.py:702-716)
  const int o = cond ? 2 : 0;
  auto sub = [&](const std::string& p, int k) { return p + "." + std::to_string(k); };
  if (cond) {
    add_conv(U, "init_conv_x", c.dim, c.data_channels, 7, true);
    add_conv(U, "init_conv_ATb", c.data_channels, c.data_channels, 7, true);
  } else {
    add_conv(U, "init_conv", c.dim, c.data_channels, 7, true);
  }
  add_param(U, "time_mlp.0.freqs", {c.time_resolution});
  add_param(U, "time_mlp.0.phases", {c.time_resolution});
  add_param(U, "time_mlp.1.weight", {U->time_dim, c.time_resolution});
  add_param(U, "time_mlp.1.bias", {U->time_dim});
  add_param(U, "time_mlp.3.weight", {U->time_dim, U->time_dim});
  add_param(U, "time_mlp.3.bias", {U->time_dim});
  for (int i = 0; i < n; ++i) {
    const int din = U->in_out[i].first, dout = U->in_out[i].second;
    const std::string p = "downs." + std::to_string(i);
    if (cond) add_embed_mix(U, p, din, i == n - 1 && n > 1);
    add_resnet(U, sub(p, o), din, din, i == n - 1 && n > 1);
    add_resnet(U, sub(p, o + 1), din, din, i == n - 1 && n > 1);
    add_attn(U, sub(p, o + 2), din, c.full_attn[i] != 0);
    if (i >= n - 1) add_conv(U, sub(p, o + 3), dout, din, 3, true, "", n > 1 ? deep_n_tile(dout) : 0);
    else add_conv(U, sub(p, o + 3) + ".conv", dout, din, 1, true);
  }
  for (int i = 0; i < n; ++i) {
    const int din = U->in_out[n - 1 - i].first, dout = U->in_out[n - 1 - i].second;
    const std::string p = "ups." + std::to_string(i);
    if (cond) add_embed_mix(U, p, dout, i == 0 && n > 1);
    add_resnet(U, sub(p, o), dout + din, dout, i == 0 && n > 1);
    add_resnet(U, sub(p, o + 1), dout + din, dout, i == 0 && n > 1);
    add_attn(U, sub(p, o + 2), dout, c.full_attn[n - 1 - i] != 0);
    if (i == n - 1) add_conv(U, sub(p, o + 3), din, dout, 3, true);
    else add_conv(U, sub(p, o + 3) + ".conv", din, dout, 3, true);
  }
  const int mid = U->dims.back();
  add_resnet(U, "mid_block1", mid, mid, n > 1);
  add_attn(U, "mid_attn", mid, true);
  add_resnet(U, "mid_block2", mid, mid, n > 1);
  add_resnet(U, "final_res_block", c.dim * 2, c.dim);
  add_conv(U, "final_conv", c.data_channels, c.dim, 1, true);

  // FiLM table in execution order: every (SiLU, Linear) time MLP of the network in one launch
  auto film_resnet = [&](const std::string& b) {
    const std::string lin = resnet_mlp(U, b);
    U->film_blocks.push_back(lin);
    U->film_gain[lin] = b + ".block1.norm.g";
  };
  auto film_mix = [&](const std::string& p) {
    U->film_blocks.push_back(p + ".1.time_mlp.1");
    U->film_gain[p + ".1.time_mlp.1"] = "";
  };
  for (int i = 0; i < n; ++i) {
    const std::string p = "downs." + std::to_string(i);
    if (cond) film_mix(p);
    film_resnet(sub(p, o));
    film_resnet(sub(p, o + 1));
  }
  film_resnet("mid_block1");
  film_resnet("mid_block2");
  for (int i = 0; i < n; ++i) {
    const std::string p = "ups." + std::to_string(i);
    if (cond) film_mix(p);
    film_resnet(sub(p, o));
    film_resnet(sub(p, o + 1));
  }
  film_resnet("final_res_block");
  int off = 0;
  for (const std::string& b : U->film_blocks) {
    U->film_off[b] = off;
    off += U->params[U->pindex[b + ".bias"]].shape[0];
  }
  U->film_rows = off;

  return 0;
}

// device storage is created on first use so that the plan (names, shapes) can be queried on a
// machine without a GPU
int ensure_device(ftb_unet* U) {
  if (U->on_device) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    FTB_FAIL("no CUDA device: the B200 path has no CPU fallback");
  for (Param& p : U->params) FTB_TRY(dev_alloc(U, &p.dev, (size_t)p.numel));
  for (auto& kv : U->convs) {
    ConvLayer& cl = kv.second;
    const size_t elems = (size_t)cl.ntiles * cl.k * cl.k * (cl.unfold_w ? 1 : cl.k) * cl.cin_pad * cl.n_tile;
    FTB_TRY(dev_alloc(U, &cl.packed, elems));
    FTB_TRY(dev_alloc(U, &cl.bias, (size_t)cl.ntiles * cl.n_tile));
    FTB_CUDA(cudaMemset(cl.bias, 0, (size_t)cl.ntiles * cl.n_tile * sizeof(float)));
    if (!cl.in_scale.empty()) FTB_TRY(dev_alloc(U, &cl.scale_tmp, (size_t)cl.cin));
  }
  for (auto& kv : U->gains) FTB_TRY(dev_alloc(U, &kv.second.gs, (size_t)kv.second.c));
  for (auto& kv : U->kshift) FTB_TRY(dev_alloc(U, &kv.second.dev, (size_t)U->cfg.attn_heads * U->cfg.attn_dim_head));
  const int nb = (int)U->film_blocks.size();
  FTB_TRY(dev_alloc(U, &U->d_film_w, (size_t)nb));
  FTB_TRY(dev_alloc(U, &U->d_film_b, (size_t)nb));
  FTB_TRY(dev_alloc(U, &U->d_film_off, (size_t)nb + 1));
  FTB_TRY(dev_alloc(U, &U->d_film_gs, (size_t)nb));
  std::vector<const float*> hw(nb), hb(nb), hg(nb);
  std::vector<int> ho(nb + 1);
  for (int i = 0; i < nb; ++i) {
    hw[i] = U->params[U->pindex[U->film_blocks[i] + ".weight"]].dev;
    hb[i] = U->params[U->pindex[U->film_blocks[i] + ".bias"]].dev;
    ho[i] = U->film_off[U->film_blocks[i]];
    const std::string& gname = U->film_gain.at(U->film_blocks[i]);
    hg[i] = gname.empty() ? nullptr : U->gains.at(gname).gs;
  }
  ho[nb] = U->film_rows;
  FTB_CUDA(cudaMemcpy(U->d_film_w, hw.data(), nb * sizeof(float*), cudaMemcpyHostToDevice));
  FTB_CUDA(cudaMemcpy(U->d_film_b, hb.data(), nb * sizeof(float*), cudaMemcpyHostToDevice));
  FTB_CUDA(cudaMemcpy(U->d_film_off, ho.data(), (nb + 1) * sizeof(int), cudaMemcpyHostToDevice));
  FTB_CUDA(cudaMemcpy(U->d_film_gs, hg.data(), nb * sizeof(float*), cudaMemcpyHostToDevice));
  U->on_device = true;
  return 0;
}

__global__ void scale_vec_kernel(const float* g, float mul, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = g[i] * mul;
}

__global__ void test_affine_kernel(const float* g, const float* scale, const float* shift, float sqrt_c,
                                   int B, int C, int stride, float* mul, float* add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  float m = g ? g[c] * sqrt_c : 1.f;
  if (scale) m *= scale[i] + 1.f;
  mul[(size_t)b * stride + c] = m;
  add[(size_t)b * stride + c] = shift ? shift[i] : 0.f;
}

int finalize_kshift(ftb_unet* U, cudaStream_t st);

int finalize(ftb_unet* U, cudaStream_t st, bool for_train = false) {
  if (!U->dirty) return (for_train || !U->kshift_stale) ? 0 : finalize_kshift(U, st);
  for (const Param& p : U->params)
    FTB_CHECK(p.set, "parameter '" + p.name + "' was never set");
  for (auto& kv : U->gains) {
    GainVec& g = kv.second;
    const float* src = U->params[U->pindex[g.gname]].dev;
    scale_vec_kernel<<<cdiv(g.c, 128), 128, 0, st>>>(src, sqrtf((float)g.c), g.gs, g.c);
  }
  // weight packs: one table-driven launch for all convs (the table holds parameter pointers: rebuilt when they move)
  const float* base0 = U->params.empty() ? nullptr : U->params[0].dev;
  if (!U->d_jobs || U->jobs_base != base0) {
    std::vector<PackJob> jobs;
    for (auto& kv : U->convs) {
      ConvLayer& cl = kv.second;
      PackJob jb{};
      jb.w = U->params[U->pindex[cl.wname]].dev;
      jb.in_scale = cl.in_scale.empty() ? nullptr : cl.scale_tmp;
      jb.dst = cl.packed;
      jb.cout = cl.cout; jb.cin_real = cl.cin; jb.ksize = cl.k; jb.cin_pad = cl.cin_pad; jb.n = cl.n_tile;
      jb.ntiles = cl.ntiles; jb.unfold_w = cl.unfold_w ? 1 : 0;
      jobs.push_back(jb);
    }
    if (!U->d_jobs) FTB_TRY(dev_alloc(U, &U->d_jobs, jobs.size()));
    FTB_CUDA(cudaMemcpyAsync(U->d_jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
    FTB_CUDA(cudaStreamSynchronize(st));   // `jobs` is pageable host memory
    U->n_jobs = (int)jobs.size();
    U->jobs_base = base0;
  }
  for (auto& kv : U->convs) {
    ConvLayer& cl = kv.second;
    if (!cl.in_scale.empty()) {
      const float* g = U->params[U->pindex[cl.in_scale]].dev;
      scale_vec_kernel<<<cdiv(cl.cin, 128), 128, 0, st>>>(g, cl.in_scale_mul, cl.scale_tmp, cl.cin);
    }
    if (!cl.bname.empty())
      FTB_CUDA(cudaMemcpyAsync(cl.bias, U->params[U->pindex[cl.bname]].dev, cl.cout * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
  }
  FTB_TRY(pack_conv_weights_batched(U->d_jobs, U->n_jobs, st));
  FTB_CUDA(cudaGetLastError());
  U->dirty = false;
  U->dgrad_dirty = true;
  U->kshift_stale = true;
  U->f32_stale = true;
  return for_train ? 0 : finalize_kshift(U, st);
}

// LinearAttention: shift[d] = 1.02*||W_k[d,:] (x) g*sqrt(C)||_2 bounds |k[d,n]| for every voxel, which
// lets the fused k/v-context kernel skip the max pass; only trusted while it is far from underflow.
// (Synchronises the stream once per layer, so the training step, which runs the unfused path, skips it.)
int finalize_kshift(ftb_unet* U, cudaStream_t st) {
  const int hd = U->cfg.attn_heads * U->cfg.attn_dim_head;
  std::vector<float> hshift(hd);
  for (auto& kv : U->kshift) {
    const ConvLayer& cl = U->convs.at(kv.first + ".to_qkv");
    const float* w = U->params[U->pindex[cl.wname]].dev;
    // the same bound for the q rows: |q[d,n]| <= 60 lets the fused q/out kernel drop the max pass of its softmax
    FTB_TRY(linattn_kshift(w, cl.scale_tmp, hd, cl.cin, kv.second.dev, st, 0));
    FTB_CUDA(cudaMemcpyAsync(hshift.data(), kv.second.dev, hd * sizeof(float), cudaMemcpyDeviceToHost, st));
    FTB_CUDA(cudaStreamSynchronize(st));
    float mq = 0.f;
    for (float v : hshift) mq = v > mq || !(v == v) ? v : mq;
    kv.second.q_ok = mq == mq && mq <= 60.f && getenv("FTB_LINATTN_EXACT") == nullptr;
    FTB_TRY(linattn_kshift(w, cl.scale_tmp, hd, cl.cin, kv.second.dev, st, hd));
    FTB_CUDA(cudaMemcpyAsync(hshift.data(), kv.second.dev, hd * sizeof(float), cudaMemcpyDeviceToHost, st));
    FTB_CUDA(cudaStreamSynchronize(st));
    float mx = 0.f;
    for (float v : hshift) mx = v > mx || !(v == v) ? v : mx;
    kv.second.ok = mx == mx && mx <= 60.f && hd == 128 && cl.cin_pad <= 128 && getenv("FTB_LINATTN_EXACT") == nullptr;
  }
  FTB_CUDA(cudaGetLastError());
  U->kshift_stale = false;
  return 0;
}

// ------------------------------------------------------------------ forward
struct Fwd {
  ftb_unet* U;
  cudaStream_t st;
  char* base;
  size_t off = 0, cap = 0;
  bool dry;
  int B;
  float* film = nullptr;
  const float* atb = nullptr;   // conditional model: ATb [atb_B, C, X, Y, Z] fp32 (atb_B = 1 or B)
  int atb_B = 0;
  bool reuse_atb = false;       // the workspace still holds the ATb-only branch of the previous call
  bool skip = false;            // allocate only (cached ATb branch)

  void* raw(size_t bytes) {
    off = round_up_sz(off, 256);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  Act act(int C, int D, int H, int W, int batch = 0) {
    Act a;
    a.B = batch > 0 ? batch : B; a.C = round_up(C, 16); a.D = D; a.H = H; a.W = W;
    a.p = reinterpret_cast<bf16*>(raw(a.bytes()));
    return a;
  }
  float* f32(size_t n) { return reinterpret_cast<float*>(raw(n * sizeof(float))); }
  const float* pdev(const std::string& n) { return U->params[U->pindex[n]].dev; }
  void tap(const std::string& name, const Act& a) {
    if (U->keep_taps) U->taps[name] = a;
  }

  int conv(const std::string& name, const ConvSrc& s0, const ConvSrc& s1, ConvEpilogue e, Act& out,
           int out_cgoff = 0) {
    const ConvLayer& cl = U->convs.at(name);
    ConvWeights w;
    w.w = cl.packed; w.ksize = cl.k; w.cin = cl.cin_pad; w.n = cl.n_tile; w.ntiles = cl.ntiles;
    w.ksize_w = cl.unfold_w ? 1 : 0;
    w.cin_real = cl.unfold_w ? cl.k * cl.cin : cl.cin; w.cout_real = cl.cout;
    if (!cl.bname.empty()) e.bias = cl.bias;
    if (dry || skip) return 0;
    U->launches += 1;
    return conv_dispatch(s0, s1, w, e, out, out_cgoff, st);
  }

  // ResnetBlock.forward (:265-278); input may be the channel concat of (x0, x1)
  // `sumsq` (optional, [B][voxels]) receives ||out voxel||^2 for a following attention's pre-norm
  // conv -> (RMSNorm * mul + add -> SiLU -> + resid): fused into the conv epilogue, or, for a layer cut into several
  // N tiles (deep_n_tile), a bias-only conv followed by the norm/activation pass
  int conv_norm_act(const std::string& name, const ConvSrc& s0, const ConvSrc& s1, ConvEpilogue e, Act& out) {
    const ConvLayer& cl = U->convs.at(name);
    if (cl.ntiles == 1) return conv(name, s0, s1, e, out);
    Act u = act(cl.cout, out.D, out.H, out.W);
    FTB_TRY(conv(name, s0, s1, ConvEpilogue{}, u));
    if (dry) return 0;
    U->launches += 1;
    return normact_fwd(u, e.norm, e.mul_stride ? nullptr : e.mul, e.mul_stride ? e.mul : nullptr, e.add, e.mul_stride,
                       e.silu, e.resid, out, st);
  }

  int resnet(const std::string& p, const Act& x0, const Act* x1, int cout, Act* out, float** sumsq = nullptr) {
    const Act& a = x0;
    ConvSrc s0{&x0, 0, x0.cg()}, s1{};
    if (x1) s1 = ConvSrc{x1, 0, x1->cg()};
    const int cin = x0.C + (x1 ? x1->C : 0);
    const float* film_p = film ? film + U->film_off.at(resnet_mlp(U, p)) : nullptr;
    Act h1 = act(cout, a.D, a.H, a.W);
    ConvEpilogue e1;
    // Block1: RMSNorm gain and FiLM (scale+1) arrive pre-multiplied from film_mlps
    e1.norm = true;
    e1.mul = film_p; e1.mul_stride = U->film_rows;
    e1.add = film_p ? film_p + cout : nullptr; e1.add_stride = U->film_rows;
    e1.silu = true;
    FTB_TRY(conv_norm_act(p + ".block1.proj", s0, s1, e1, h1));
    tap(p + ".block1", h1);
    Act res;
    const Act* resp = &x0;
    if (cin != cout) {
      res = act(cout, a.D, a.H, a.W);
      FTB_TRY(conv(p + ".res_conv", s0, s1, ConvEpilogue{}, res));
      resp = &res;
    }
    *out = act(cout, a.D, a.H, a.W);
    ConvEpilogue e2;
    e2.norm = true;
    e2.mul = U->gains.at(p + ".block2.norm.g").gs;
    e2.silu = true;
    e2.resid = resp;
    // ||out voxel||^2 for a following attention's pre-norm comes from the fused epilogue only
    if (sumsq && U->convs.at(p + ".block2.proj").ntiles > 1) *sumsq = nullptr;
    e2.sumsq_out = sumsq ? *sumsq : nullptr;
    FTB_TRY(conv_norm_act(p + ".block2.proj", ConvSrc{&h1, 0, h1.cg()}, ConvSrc{}, e2, *out));
    tap(p, *out);
    return 0;
  }

  // x + attn(x)  (:695, :702, :712)
  int attention(const std::string& p, const Act& x, const float* x_sumsq, bool full, Act* out) {
    const ftb_unet_cfg& c = U->cfg;
    const int heads = c.attn_heads, dh = c.attn_dim_head, hd = heads * dh;
    // (the dry sizing pass reserves the larger, unfused footprint)
    const bool fused = !dry && !full && x_sumsq != nullptr && U->kshift.count(p) && U->kshift.at(p).ok;
    // fused path: q, k and v are never materialised (kvctx_kernel, qout_kernel)
    Act qkv;
    if (!fused) qkv = act(3 * hd, x.D, x.H, x.W);
    ConvEpilogue eq;
    eq.prenorm = true;
    eq.prenorm_ss = x_sumsq;
    if (!full) { eq.q_softmax_heads = heads; eq.q_dim_head = dh; eq.q_scale = 1.f / sqrtf((float)dh); }
    const ConvLayer& cq = U->convs.at(p + ".to_qkv");
    if (!fused) FTB_TRY(conv(p + ".to_qkv", ConvSrc{&x, 0, x.cg()}, ConvSrc{}, eq, qkv));
    *out = act(x.C, x.D, x.H, x.W);
    if (full) {
      Act ao = act(hd, x.D, x.H, x.W);
      if (!dry) {
        FTB_TRY(full_attention(qkv, heads, dh, pdev(p + ".mem_kv"), c.num_mem_kv, ao, st));
        U->launches += 1;
      }
      ConvEpilogue eo;
      eo.resid = &x;
      FTB_TRY(conv(p + ".to_out", ConvSrc{&ao, 0, ao.cg()}, ConvSrc{}, eo, *out));
    } else {
      const size_t vox = x.voxels();
      // voxel slices per sample for the context GEMM: about two CTAs per SM over the batch, each
      // with at least a few 128-voxel tiles
      static const int kv_mult = getenv("FTB_KV_SPLIT") ? atoi(getenv("FTB_KV_SPLIT")) : 2;
      int nsplit = cdiv(kv_mult * num_sms(), B);
      const int max_split = (int)((vox + 511) / 512);
      nsplit = nsplit > max_split ? max_split : nsplit;
      nsplit = nsplit < 1 ? 1 : (nsplit > 256 ? 256 : nsplit);
      float* kmax = f32((size_t)B * hd);
      float* part = f32((size_t)B * heads * nsplit * (dh * dh + dh));
      bf16* mpack = reinterpret_cast<bf16*>(raw((size_t)B * x.C * hd * sizeof(bf16)));
      if (!dry) {
        if (fused) {
          const size_t tile = (size_t)cq.cin_pad * cq.n_tile;
          const float* shift = U->kshift.at(p).dev;
          FTB_TRY(linattn_kv_context(x, 0, x.cg(), x_sumsq, cq.packed + tile, cq.packed + 2 * tile, shift, heads, dh,
                                     nsplit, part, st));
          FTB_TRY(linattn_combine(part, nsplit, shift, 0, B, heads, dh, pdev(p + ".mem_kv"), c.num_mem_kv,
                                  pdev(p + ".to_out.0.weight"), x.C, 1.f, mpack, nullptr, st));
          U->launches += 2;
        } else {
          FTB_TRY(linattn_kmax(qkv, heads, dh, nsplit, kmax, st));
          FTB_TRY(linattn_context_partial(qkv, heads, dh, nsplit, kmax, part, st));
          FTB_TRY(linattn_combine(part, nsplit, kmax, hd, B, heads, dh, pdev(p + ".mem_kv"), c.num_mem_kv,
                                  pdev(p + ".to_out.0.weight"), x.C, 1.f, mpack, nullptr, st));
          U->launches += 4;
        }
        if (fused) {
          FTB_TRY(linattn_q_out(x, x_sumsq, cq.packed, mpack, (long long)x.C * hd, pdev(p + ".to_out.0.bias"),
                                U->gains.at(p + ".to_out.1.g").gs, heads, dh, *out, st, U->kshift.at(p).q_ok));
          U->launches += 1;
        } else {
          ConvWeights w;
          w.w = mpack; w.ksize = 1; w.cin = hd; w.n = x.C; w.ntiles = 1;
          w.batch_stride = (long long)x.C * hd;
          ConvEpilogue eo;
          eo.bias = pdev(p + ".to_out.0.bias");
          eo.norm = true;
          eo.mul = U->gains.at(p + ".to_out.1.g").gs;
          eo.resid = &x;
          FTB_TRY(conv_dispatch(ConvSrc{&qkv, 0, hd / 8}, ConvSrc{}, w, eo, *out, 0, st));
          U->launches += 1;
        }
      }
    }
    tap(p, *out);
    return 0;
  }

  // NCDHW fp32 network input -> blocked bf16 operand of the stem conv `name` (W-unfolded when the
  // stem was planned that way)
  int pack_input(const std::string& name, const float* x, int batch, int X, int Y, int Z, Act* out) {
    const ConvLayer& cl = U->convs.at(name);
    *out = act(cl.unfold_w ? cl.k * cl.cin : cl.cin, X, Y, Z, batch);
    if (dry || skip) return 0;
    U->launches += 1;
    if (cl.unfold_w) return pack_unfold_w(x, batch, cl.cin, X, Y, Z, cl.k, *out, st);
    return pack_ncdhw_to_blocked(x, batch, cl.cin, X, Y, Z, *out, st);
  }

  // EmbedATb.forward (unet_attn_3d_cond_v3.py:131-139) on the opened ATb; depends on ATb only, so a
  // sampler computes it once per trajectory (reuse_atb) instead of once per evaluation.
  int embed_atb(const std::string& p, const Act& opened, int C, int D, int H, int W, Act* out) {
    Act src = opened;
    if (opened.D != D || opened.H != H || opened.W != W) {
      src = act(opened.C, D, H, W, opened.B);
      if (!dry && !skip) {
        FTB_TRY(trilinear_resample(opened, src, st));
        U->launches += 1;
      }
    }
    Act e1 = act(C, D, H, W, opened.B);
    ConvEpilogue a1;
    a1.silu = true;
    FTB_TRY(conv(p + ".conv1", ConvSrc{&src, 0, src.cg()}, ConvSrc{}, a1, e1));
    *out = act(C, D, H, W, opened.B);
    FTB_TRY(conv(p + ".conv2", ConvSrc{&e1, 0, e1.cg()}, ConvSrc{}, ConvEpilogue{}, *out));
    return 0;
  }

  // MixATb.forward (unet_attn_3d_cond_v3.py:175-190)
  int mix_atb(const std::string& p, const Act& x, const Act& emb, Act* out) {
    const int C = x.C;
    Act cat = act(2 * C, x.D, x.H, x.W);
    if (!dry) {
      FTB_TRY(film_concat(x, emb, film + U->film_off.at(p + ".time_mlp.1"), U->film_rows, C, cat, st));
      U->launches += 1;
    }
    Act h1 = act(C, x.D, x.H, x.W);
    ConvEpilogue e1;
    e1.norm = true;
    e1.mul = U->gains.at(p + ".norm.g").gs;
    e1.silu = true;
    FTB_TRY(conv(p + ".conv1", ConvSrc{&cat, 0, cat.cg()}, ConvSrc{}, e1, h1));
    *out = act(C, x.D, x.H, x.W);
    ConvEpilogue e2;
    e2.resid = &x;
    FTB_TRY(conv(p + ".conv2", ConvSrc{&h1, 0, h1.cg()}, ConvSrc{}, e2, *out));
    tap(p, *out);
    return 0;
  }

  int resample(const Act& in, int D, int H, int W, Act* out) {
    *out = act(in.C, D, H, W);
    if (dry) return 0;
    U->launches += 1;
    return trilinear_resample(in, *out, st);
  }

  int run(const float* x, const float* t, float* y, int X, int Y, int Z) {
    const ftb_unet_cfg& c = U->cfg;
    const int n = c.n_stages;
    const bool cond = c.conditional != 0;
    const int o = cond ? 2 : 0;
    auto sub = [&](const std::string& p, int k) { return p + "." + std::to_string(k); };
    U->taps.clear();
    U->launches = 0;
    // ---- ATb-only branch first, so that its place in the arena does not depend on anything else:
    // init_conv_ATb (:778) and the EmbedATb of every down and up stage (:793, :812)
    std::vector<Act> emb_down(n), emb_up(n);
    if (cond) {
      skip = reuse_atb;
      Act ain;
      FTB_TRY(pack_input("init_conv_ATb", atb, atb_B, X, Y, Z, &ain));
      Act opened = act(c.data_channels, X, Y, Z, atb_B);
      FTB_TRY(conv("init_conv_ATb", ConvSrc{&ain, 0, ain.cg()}, ConvSrc{}, ConvEpilogue{}, opened));
      if (!skip) tap("init_conv_ATb", opened);
      for (int i = 0; i < n; ++i) {
        const int f = 1 << i;
        FTB_TRY(embed_atb("downs." + std::to_string(i) + ".0", opened, U->in_out[i].first, X / f, Y / f, Z / f,
                          &emb_down[i]));
        if (!skip) tap("downs." + std::to_string(i) + ".0", emb_down[i]);
      }
      for (int i = 0; i < n; ++i) {
        const int f = 1 << (n - 1 - i);
        FTB_TRY(embed_atb("ups." + std::to_string(i) + ".0", opened, U->in_out[n - 1 - i].second, X / f, Y / f,
                          Z / f, &emb_up[i]));
        if (!skip) tap("ups." + std::to_string(i) + ".0", emb_up[i]);
      }
      skip = false;
    }
    // time path
    float* temb = f32((size_t)B * U->time_dim);
    float* temb_silu = f32((size_t)B * U->time_dim);
    film = f32((size_t)B * U->film_rows);
    const std::string init_name = cond ? "init_conv_x" : "init_conv";
    Act xin;
    FTB_TRY(pack_input(init_name, x, B, X, Y, Z, &xin));
    if (!dry) {
      TimeMlpParams tp{pdev("time_mlp.0.freqs"), pdev("time_mlp.0.phases"), pdev("time_mlp.1.weight"),
                       pdev("time_mlp.1.bias"), pdev("time_mlp.3.weight"), pdev("time_mlp.3.bias"),
                       c.time_resolution, U->time_dim};
      FTB_TRY(time_embed(tp, t, B, temb, temb_silu, st));
      FilmTable ft{U->d_film_w, U->d_film_b, U->d_film_gs, U->d_film_off, (int)U->film_blocks.size(),
                   U->film_rows, U->time_dim};
      FTB_TRY(film_mlps(ft, temb_silu, B, film, st));
      U->launches += 2;
    }
    Act r = act(c.dim, X, Y, Z);
    FTB_TRY(conv(init_name, ConvSrc{&xin, 0, xin.cg()}, ConvSrc{}, ConvEpilogue{}, r));
    tap(init_name, r);
    Act cur = r;
    std::vector<Act> skips;
    for (int i = 0; i < n; ++i) {
      const std::string p = "downs." + std::to_string(i);
      const int din = U->in_out[i].first, dout = U->in_out[i].second;
      Act a1, a2, a3, a4;
      if (cond) {
        Act mixed;
        FTB_TRY(mix_atb(sub(p, 1), cur, emb_down[i], &mixed));
        cur = mixed;
      }
      FTB_TRY(resnet(sub(p, o), cur, nullptr, din, &a1));
      skips.push_back(a1);
      float* ss = f32((size_t)B * a1.voxels());
      FTB_TRY(resnet(sub(p, o + 1), a1, nullptr, din, &a2, &ss));
      FTB_TRY(attention(sub(p, o + 2), a2, ss, c.full_attn[i] != 0, &a3));
      skips.push_back(a3);
      if (i >= n - 1) {
        a4 = act(dout, a3.D, a3.H, a3.W);
        FTB_TRY(conv(sub(p, o + 3), ConvSrc{&a3, 0, a3.cg()}, ConvSrc{}, ConvEpilogue{}, a4));
      } else {
        Act ds;
        FTB_TRY(resample(a3, a3.D / 2, a3.H / 2, a3.W / 2, &ds));
        a4 = act(dout, ds.D, ds.H, ds.W);
        FTB_TRY(conv(sub(p, o + 3) + ".conv", ConvSrc{&ds, 0, ds.cg()}, ConvSrc{}, ConvEpilogue{}, a4));
      }
      tap(sub(p, o + 3), a4);
      cur = a4;
    }
    {
      Act m1, m2, m3;
      const int mid = U->dims.back();
      float* ss = f32((size_t)B * cur.voxels());
      FTB_TRY(resnet("mid_block1", cur, nullptr, mid, &m1, &ss));
      FTB_TRY(attention("mid_attn", m1, ss, true, &m2));
      FTB_TRY(resnet("mid_block2", m2, nullptr, mid, &m3));
      cur = m3;
    }
    for (int i = 0; i < n; ++i) {
      const std::string p = "ups." + std::to_string(i);
      const int din = U->in_out[n - 1 - i].first, dout = U->in_out[n - 1 - i].second;
      Act a1, a2, a3, a4;
      if (cond) {
        Act mixed;
        FTB_TRY(mix_atb(sub(p, 1), cur, emb_up[i], &mixed));
        cur = mixed;
      }
      Act s = skips.back(); skips.pop_back();
      FTB_TRY(resnet(sub(p, o), cur, &s, dout, &a1));
      s = skips.back(); skips.pop_back();
      float* ss = f32((size_t)B * a1.voxels());
      FTB_TRY(resnet(sub(p, o + 1), a1, &s, dout, &a2, &ss));
      FTB_TRY(attention(sub(p, o + 2), a2, ss, c.full_attn[n - 1 - i] != 0, &a3));
      if (i == n - 1) {
        a4 = act(din, a3.D, a3.H, a3.W);
        FTB_TRY(conv(sub(p, o + 3), ConvSrc{&a3, 0, a3.cg()}, ConvSrc{}, ConvEpilogue{}, a4));
      } else {
        Act us;
        FTB_TRY(resample(a3, a3.D * 2, a3.H * 2, a3.W * 2, &us));
        a4 = act(din, us.D, us.H, us.W);
        FTB_TRY(conv(sub(p, o + 3) + ".conv", ConvSrc{&us, 0, us.cg()}, ConvSrc{}, ConvEpilogue{}, a4));
      }
      tap(sub(p, o + 3), a4);
      cur = a4;
    }
    Act fin;
    FTB_TRY(resnet("final_res_block", cur, &r, c.dim, &fin));
    ConvEpilogue ef;
    ef.out_f32 = y;
    ef.out_f32_c = c.data_channels;
    Act dummy = fin;  // spatial dims only; the epilogue writes NCDHW fp32 to y
    FTB_TRY(conv("final_conv", ConvSrc{&fin, 0, fin.cg()}, ConvSrc{}, ef, dummy));
    return 0;
  }
};

}  // namespace ftb_engine_detail
#include "engine_train.cuh"
#include "engine_f32.cuh"
using namespace ftb_engine_detail;

// ====================================================================== C ABI
extern "C" {

const char* ftb_last_error(void) { return g_err.c_str(); }
int ftb_version(void) { return 100; }
int ftb_device_sm_count(void) { return num_sms(); }
int64_t ftb_launch_count(void) { return launch_count(); }
int ftb_profile_enable(int on) {
  prof_set(on != 0);
  return 0;
}
int ftb_profile_collect(double* flops, double* bytes, double* ms, int* launches, int nkinds) {
  FTB_CHECK(flops && bytes && ms && launches && nkinds >= 1, "null argument");
  return prof_collect(flops, bytes, ms, launches, nkinds) >= 0 ? 0 : -1;
}

int ftb_unet3d_create(const ftb_unet_cfg* cfg, ftb_unet** out) {
  FTB_CHECK(cfg && out, "null argument");
  std::unique_ptr<ftb_unet> U(new ftb_unet());
  U->cfg = *cfg;
  FTB_TRY(build_plan(U.get()));
  *out = U.release();
  return 0;
}

int ftb_unet3d_destroy(ftb_unet* h) {
  delete h;
  return 0;
}

int ftb_unet3d_num_params(const ftb_unet* h) { return h ? (int)h->params.size() : 0; }
const char* ftb_unet3d_param_name(const ftb_unet* h, int i) {
  return (h && i >= 0 && i < (int)h->params.size()) ? h->params[i].name.c_str() : nullptr;
}
int64_t ftb_unet3d_param_numel(const ftb_unet* h, int i) {
  return (h && i >= 0 && i < (int)h->params.size()) ? h->params[i].numel : -1;
}
int ftb_unet3d_param_shape(const ftb_unet* h, int i, int* dims, int max_dims) {
  if (!h || i < 0 || i >= (int)h->params.size() || !dims) return -1;
  const std::vector<int>& s = h->params[i].shape;
  if ((int)s.size() > max_dims) return -1;
  for (size_t k = 0; k < s.size(); ++k) dims[k] = s[k];
  return (int)s.size();
}

int ftb_unet3d_set_param(ftb_unet* h, const char* name, const float* data, int64_t numel, void* stream) {
  FTB_CHECK(h && name && data, "null argument");
  auto it = h->pindex.find(name);
  FTB_CHECK(it != h->pindex.end(), std::string("unknown parameter '") + name + "'");
  Param& p = h->params[it->second];
  FTB_CHECK(p.numel == numel, std::string("parameter '") + name + "': expected " +
                                  std::to_string(p.numel) + " elements, got " + std::to_string(numel));
  FTB_TRY(ensure_device(h));
  FTB_CUDA(cudaMemcpyAsync(p.dev, data, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  p.set = true;
  h->dirty = true;
  return 0;
}

static int check_dims(const ftb_unet* h, int B, int X, int Y, int Z) {
  FTB_CHECK(h, "null handle");
  FTB_CHECK(B >= 1 && X >= 1 && Y >= 1 && Z >= 1, "non-positive dims");
  const int f = 1 << (h->cfg.n_stages - 1);
  FTB_CHECK(X % f == 0 && Y % f == 0 && Z % f == 0,
            "input dimensions need to be divisible by " + std::to_string(f));
  return 0;
}

static size_t workspace_bytes_impl(ftb_unet* h, int B, int atb_B, int X, int Y, int Z) {
  if (check_dims(h, B, X, Y, Z) != 0) return 0;
  Fwd f{h, nullptr, nullptr, 0, 0, true, B};
  f.atb_B = atb_B;
  const bool kt = h->keep_taps;
  h->keep_taps = false;
  const int r = f.run(nullptr, nullptr, nullptr, X, Y, Z);
  h->keep_taps = kt;
  return r == 0 ? round_up_sz(f.off, 256) + 256 : 0;
}

size_t ftb_unet3d_workspace_bytes(ftb_unet* h, int B, int X, int Y, int Z) {
  return workspace_bytes_impl(h, B, h && h->cfg.conditional ? B : 0, X, Y, Z);
}
size_t ftb_unet3d_cond_workspace_bytes(ftb_unet* h, int B, int atb_B, int X, int Y, int Z) {
  if (!h || !h->cfg.conditional || !(atb_B == 1 || atb_B == B)) return 0;
  return workspace_bytes_impl(h, B, atb_B, X, Y, Z);
}

// A sampling / fp32 forward on the handle ends the life of a recorded training tape: the host keeps ONE resident
// workspace per module, so the tape's activations are about to be freed or overwritten (a later backward would
// otherwise pass the base-pointer check on a recycled block and read clobbered activations).
static void invalidate_tape(ftb_unet* h) {
  if (h->train) h->train->valid = false;
}

int ftb_unet3d_forward(ftb_unet* h, const float* x, const float* t, float* out, int B, int X, int Y,
                       int Z, void* workspace, size_t workspace_bytes, void* stream) {
  FTB_TRY(check_dims(h, B, X, Y, Z));
  FTB_CHECK(!h->cfg.conditional, "conditional model: call ftb_unet3d_cond_forward (ATb is required)");
  FTB_CHECK(x && t && out && workspace, "null argument");
  FTB_CHECK(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const size_t need = ftb_unet3d_workspace_bytes(h, B, X, Y, Z);
  FTB_CHECK(need > 0 && workspace_bytes >= need,
            "workspace too small: need " + std::to_string(need) + " bytes");
  cudaStream_t st = (cudaStream_t)stream;
  FTB_TRY(ensure_device(h));
  FTB_TRY(finalize(h, st));
  invalidate_tape(h);
  Fwd f{h, st, reinterpret_cast<char*>(workspace), 0, workspace_bytes, false, B};
  return f.run(x, t, out, X, Y, Z);
}

size_t ftb_unet3d_f32_workspace_bytes(ftb_unet* h, int B, int X, int Y, int Z) {
  if (check_dims(h, B, X, Y, Z) != 0) return 0;
  FwdF32 f{h, nullptr, nullptr, true, B};
  return f.run(nullptr, nullptr, nullptr, X, Y, Z) == 0 ? round_up_sz(f.peak, 256) + 256 : 0;
}

static int forward_f32_impl(ftb_unet* h, const float* x, const float* atb, const float* t, float* out, int B, int X,
                            int Y, int Z, void* workspace, size_t workspace_bytes, void* stream) {
  FTB_TRY(check_dims(h, B, X, Y, Z));
  FTB_CHECK(x && t && out && workspace, "null argument");
  FTB_CHECK(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const size_t need = ftb_unet3d_f32_workspace_bytes(h, B, X, Y, Z);
  FTB_CHECK(need > 0 && workspace_bytes >= need, "workspace too small: need " + std::to_string(need) + " bytes");
  cudaStream_t st = (cudaStream_t)stream;
  FTB_TRY(ensure_device(h));
  FTB_TRY(finalize(h, st, true));
  FTB_TRY(finalize_f32(h, st));
  invalidate_tape(h);
  FwdF32 f{h, st, reinterpret_cast<char*>(workspace), false, B};
  return f.run(x, t, out, X, Y, Z, atb);
}

int ftb_unet3d_forward_f32(ftb_unet* h, const float* x, const float* t, float* out, int B, int X, int Y, int Z,
                           void* workspace, size_t workspace_bytes, void* stream) {
  FTB_CHECK(h, "null handle");
  FTB_CHECK(!h->cfg.conditional, "conditional model: call ftb_unet3d_cond_forward_f32 (ATb is required)");
  return forward_f32_impl(h, x, nullptr, t, out, B, X, Y, Z, workspace, workspace_bytes, stream);
}

int ftb_unet3d_cond_forward_f32(ftb_unet* h, const float* x, const float* atb, const float* t, float* out, int B,
                                int X, int Y, int Z, void* workspace, size_t workspace_bytes, void* stream) {
  FTB_CHECK(h && atb, "null argument");
  FTB_CHECK(h->cfg.conditional, "unconditional model: call ftb_unet3d_forward_f32");
  return forward_f32_impl(h, x, atb, t, out, B, X, Y, Z, workspace, workspace_bytes, stream);
}

int ftb_unet3d_cond_forward(ftb_unet* h, const float* x, const float* atb, int atb_B, const float* t, float* out,
                            int B, int X, int Y, int Z, void* workspace, size_t workspace_bytes, int reuse_atb,
                            void* stream) {
  FTB_TRY(check_dims(h, B, X, Y, Z));
  FTB_CHECK(h->cfg.conditional, "unconditional model: call ftb_unet3d_forward");
  FTB_CHECK(x && atb && t && out && workspace, "null argument");
  FTB_CHECK(atb_B == 1 || atb_B == B, "ATb batch must be 1 (shared by the batch) or B");
  FTB_CHECK(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  const size_t need = ftb_unet3d_cond_workspace_bytes(h, B, atb_B, X, Y, Z);
  FTB_CHECK(need > 0 && workspace_bytes >= need,
            "workspace too small: need " + std::to_string(need) + " bytes");
  cudaStream_t st = (cudaStream_t)stream;
  FTB_TRY(ensure_device(h));
  const bool was_dirty = h->dirty;
  FTB_TRY(finalize(h, st));
  invalidate_tape(h);
  Fwd f{h, st, reinterpret_cast<char*>(workspace), 0, workspace_bytes, false, B};
  f.atb = atb;
  f.atb_B = atb_B;
  f.reuse_atb = reuse_atb != 0 && !was_dirty;   // new weights invalidate the cached branch
  return f.run(x, t, out, X, Y, Z);
}

int ftb_unet3d_tap_channels(ftb_unet* h, const char* name, int* C, int* X, int* Y, int* Z) {
  FTB_CHECK(h && name, "null argument");
  auto it = h->taps.find(name);
  FTB_CHECK(it != h->taps.end(), std::string("no tap named '") + name + "'");
  if (C) *C = it->second.C;
  if (X) *X = it->second.D;
  if (Y) *Y = it->second.H;
  if (Z) *Z = it->second.W;
  return 0;
}

int ftb_unet3d_get_tap(ftb_unet* h, const char* name, float* out, void* stream) {
  FTB_CHECK(h && name && out, "null argument");
  auto it = h->taps.find(name);
  FTB_CHECK(it != h->taps.end(), std::string("no tap named '") + name + "'");
  return unpack_blocked_to_ncdhw(it->second, 0, it->second.C, out, (cudaStream_t)stream);
}

/* fp32 mode: copy a named fp32 intermediate of the last ftb_unet3d_forward_f32 (dims: B, C, X, Y, Z) */
int ftb_unet3d_get_tap_f32(ftb_unet* h, const char* name, float* out, int* dims, void* stream) {
  FTB_CHECK(h && name, "null argument");
  auto it = h->taps32.find(name);
  FTB_CHECK(it != h->taps32.end(), std::string("no fp32 tap named '") + name + "'");
  const std::vector<long long>& v = it->second;
  if (dims) for (int i = 0; i < 5; ++i) dims[i] = (int)v[1 + i];
  if (out)
    FTB_CUDA(cudaMemcpyAsync(out, reinterpret_cast<const float*>((uintptr_t)v[0]),
                             (size_t)(v[1] * v[2] * v[3] * v[4] * v[5]) * sizeof(float), cudaMemcpyDeviceToDevice,
                             (cudaStream_t)stream));
  return 0;
}

int ftb_unet3d_last_launches(const ftb_unet* h) { return h ? h->launches : 0; }

int ftb_interp_xt_bt(int kind, int one_sided, float gamma_a, const float* x0, const float* x1,
                     const float* z, const float* t, float* xt, float* bt, int B, int64_t n, void* stream) {
  FTB_CHECK(x0 && x1 && t && xt, "null argument");
  return interp_xt_bt(kind, one_sided, gamma_a, x0, x1, z, t, xt, bt, B, n, (cudaStream_t)stream);
}
int ftb_ode_axpy(float* out, const float* x, const float* k, double h, int64_t n,
                 const unsigned char* frozen, int64_t inner, void* stream) {
  FTB_CHECK(out && x && k, "null argument");
  return axpy_out(out, x, k, (float)h, n, frozen, inner, (cudaStream_t)stream);
}
int ftb_ode_heun_combine(float* out, const float* x, const float* k1, const float* k2, double h,
                         int64_t n, void* stream) {
  FTB_CHECK(out && x && k1 && k2, "null argument");
  return heun_combine(out, x, k1, k2, h, n, (cudaStream_t)stream);
}
int ftb_ode_rk4_combine(float* out, const float* x, const float* k1, const float* k2, const float* k3,
                        const float* k4, double h, int64_t n, void* stream) {
  FTB_CHECK(out && x && k1 && k2 && k3 && k4, "null argument");
  return rk4_combine(out, x, k1, k2, k3, k4, h, n, (cudaStream_t)stream);
}
int ftb_denoise_drift(float* out, const float* x, const float* eta, const float* noise, float alpha,
                      float beta, float alpha_dot, float beta_dot, float eps, int use_sde, int64_t n,
                      void* stream) {
  FTB_CHECK(out && x && eta, "null argument");
  return denoise_drift(out, x, eta, noise, alpha, beta, alpha_dot, beta_dot, eps, use_sde, n,
                       (cudaStream_t)stream);
}
int ftb_decode(const float* x, const float* en, int64_t* out, int B, int E, int ncat, int64_t n, void* stream) {
  FTB_CHECK(x && en && out, "null argument");
  return decode_argmax(x, en, reinterpret_cast<long long*>(out), B, E, ncat, n, (cudaStream_t)stream);
}
int ftb_ode_lincomb(float* out, const float* y0, const float* const* k, const double* coef, int nk, int64_t n,
                    void* stream) {
  FTB_CHECK(out && y0, "null argument");
  return ode_lincomb(out, y0, k, coef, nk, n, (cudaStream_t)stream);
}
int ftb_ode_error_ratio(const float* y0, const float* y1, const float* const* k, const double* coef, int nk, float rtol,
                        float atol, int64_t n, double* acc, void* stream) {
  FTB_CHECK(y0 && y1 && acc, "null argument");
  return ode_error_ratio(y0, y1, k, coef, nk, rtol, atol, n, acc, (cudaStream_t)stream);
}
int ftb_ode_scaled_sumsq(const float* a1, const float* a2, const float* y, float rtol, float atol, int64_t n,
                         double* acc, void* stream) {
  FTB_CHECK(a1 && y && acc, "null argument");
  return ode_scaled_sumsq(a1, a2, y, rtol, atol, n, acc, (cudaStream_t)stream);
}
int ftb_ode_dense_eval(float* out, const float* y0, const float* y1, const float* ymid, const float* f0,
                       const float* f1, double dt, double x, int64_t n, void* stream) {
  FTB_CHECK(out && y0 && y1 && ymid && f0 && f1, "null argument");
  return ode_dense_eval(out, y0, y1, ymid, f0, f1, dt, x, n, (cudaStream_t)stream);
}
int ftb_ode_ctl_init(double* ctl, double t0, void* stream) {
  FTB_CHECK(ctl, "null argument");
  return ode_ctl_init(ctl, t0, (cudaStream_t)stream);
}
int ftb_ode_ctl_first_step(double* ctl, int phase, int64_t n, int order, void* stream) {
  FTB_CHECK(ctl && n > 0 && order >= 1 && (phase == 0 || phase == 1), "bad argument");
  return ode_ctl_first_step(ctl, phase, n, order, (cudaStream_t)stream);
}
int ftb_ode_ctl_stage_time(float* tbuf, const double* ctl, double alpha, int B, void* stream) {
  FTB_CHECK(tbuf && ctl && B > 0, "bad argument");
  return ode_ctl_stage_time(tbuf, ctl, alpha, B, (cudaStream_t)stream);
}
int ftb_ode_lincomb_dev(float* out, const float* y0, const float* const* k, const double* coef, int nk, int64_t n,
                        const double* ctl, void* stream) {
  FTB_CHECK(out && y0 && ctl, "null argument");
  return ode_lincomb_dev(out, y0, k, coef, nk, n, ctl, (cudaStream_t)stream);
}
int ftb_ode_error_ratio_dev(const float* y0, const float* y1, const float* const* k, const double* coef, int nk,
                            float rtol, float atol, int64_t n, double* ctl, void* stream) {
  FTB_CHECK(y0 && y1 && ctl, "null argument");
  return ode_error_ratio_dev(y0, y1, k, coef, nk, rtol, atol, n, ctl, (cudaStream_t)stream);
}
int ftb_ode_ctl_step(double* ctl, const double* grid, int n_out, int64_t n, int order, int64_t max_steps, void* stream) {
  FTB_CHECK(ctl && grid && n_out >= 1 && n > 0, "bad argument");
  return ode_ctl_step(ctl, grid, n_out, n, order, max_steps, (cudaStream_t)stream);
}
int ftb_ode_advance(float* y0, float* f0, const float* y1, const float* f1, const float* const* k, const double* c_mid,
                    int nk, const double* ctl, const double* grid, int n_out, float* traj, float* last, int64_t n,
                    void* stream) {
  FTB_CHECK(y0 && f0 && y1 && f1 && ctl && grid, "null argument");
  return ode_advance(y0, f0, y1, f1, k, c_mid, nk, ctl, grid, n_out, traj, last, n, (cudaStream_t)stream);
}
int ftb_denoise_drift_dev(float* out, const float* x, const float* eta, const float* noise, const float* coef,
                          int use_sde, int64_t n, void* stream) {
  FTB_CHECK(out && x && eta && coef, "null argument");
  return denoise_drift_dev(out, x, eta, noise, coef, use_sde, n, (cudaStream_t)stream);
}
int ftb_cond_frontend(const int64_t* cats, const int32_t* bores, const int32_t* n_bores, int max_bores, const float* w,
                      int B, int E, int ncat, int shift, int X, int Y, int Z, int surface, uint8_t* mask, float* x1,
                      float* atb, void* stream) {
  FTB_CHECK(cats && (mask || x1 || atb), "null argument");
  FTB_CHECK(w || (!x1 && !atb), "embedding matrix missing");
  FTB_CHECK(max_bores == 0 || (bores && n_bores), "borehole table missing");
  FTB_CHECK(B >= 1 && X >= 1 && Y >= 1 && Z >= 1, "non-positive dims");
  return cond_frontend(reinterpret_cast<const long long*>(cats), bores, n_bores, max_bores, w, B, E, ncat, shift, X, Y, Z,
                       surface, mask, x1, atb, (cudaStream_t)stream);
}
int ftb_cond_loss_accumulate(const float* vt, const float* vhat, const float* xt, const float* x1_clean,
                             const float* x1_noisy, const uint8_t* mask, const float* t, int B, int E, int64_t n,
                             double* acc6, void* stream) {
  FTB_CHECK(vt && vhat && xt && x1_clean && x1_noisy && mask && t && acc6, "null argument");
  return cond_loss_partial(vt, vhat, xt, x1_clean, x1_noisy, mask, t, B, E, n, acc6, (cudaStream_t)stream);
}
int ftb_cond_loss_grad(const float* vt, const float* vhat, const float* xt, const float* x1_clean, const uint8_t* mask,
                       const float* t, int B, int E, int64_t n, const double* acc6, float lambda_reconstruct, float scale,
                       float* dout, void* stream) {
  FTB_CHECK(vt && vhat && xt && x1_clean && mask && t && acc6 && dout, "null argument");
  return cond_loss_grad(vt, vhat, xt, x1_clean, mask, t, B, E, n, acc6, lambda_reconstruct, scale, dout,
                        (cudaStream_t)stream);
}
int ftb_decode_vote(const float* x, const float* en, int S, int E, int ncat, int64_t n, int64_t* decoded,
                    int32_t* counts, void* stream) {
  FTB_CHECK(x && en && counts, "null argument");
  return decode_vote(x, en, S, E, ncat, n, reinterpret_cast<long long*>(decoded), counts, (cudaStream_t)stream);
}
int ftb_vote_finalize(const int32_t* counts, int S, int ncat, int64_t n, int shift, float* probs, float* entropy,
                      int64_t* most_probable, float* entropy_masked, void* stream) {
  FTB_CHECK(counts && (probs || entropy || most_probable || entropy_masked), "null argument");
  return vote_finalize(counts, S, ncat, n, shift, probs, entropy, reinterpret_cast<long long*>(most_probable),
                       entropy_masked, (cudaStream_t)stream);
}
int ftb_decode_logits(const float* x, const float* en, float* logits, int B, int E, int ncat, int64_t n, void* stream) {
  FTB_CHECK(x && en && logits, "null argument");
  return decode_logits(x, en, logits, B, E, ncat, n, (cudaStream_t)stream);
}
int ftb_embed(const int64_t* cats, const float* w, float* out, int B, int E, int ncat, int64_t n,
              int shift, void* stream) {
  FTB_CHECK(cats && w && out, "null argument");
  return embed_lookup(reinterpret_cast<const long long*>(cats), w, out, B, E, ncat, n, shift,
                      (cudaStream_t)stream);
}
int ftb_ema_update(float* shadow, const float* param, int64_t n, double decay, void* stream) {
  FTB_CHECK(shadow && param, "null argument");
  return ema_update(shadow, param, n, decay, (cudaStream_t)stream);
}
int ftb_mse_ratio_accumulate(const float* v, const float* vhat, int64_t n, double* acc2, void* stream) {
  FTB_CHECK(v && vhat && acc2, "null argument");
  return mse_ratio_partial(v, vhat, n, acc2, (cudaStream_t)stream);
}

}  // extern "C"


// ---------------------------------------------------------------------- training step
namespace ftb_engine_detail {
static void ensure_offsets(ftb_unet* h, TrainState* T) {
  if (!T->poff.empty()) return;
  int64_t off = 0;
  for (const Param& p : h->params) {
    T->poff.push_back(off);
    off += p.numel;
  }
  T->ptotal = off;
}
static size_t max_wt_elems(const ftb_unet* h) {
  size_t m = 16;
  for (const auto& kv : h->convs) {
    const ConvLayer& cl = kv.second;
    const size_t e = (size_t)cl.cout * cl.cin * cl.k * cl.k * cl.k;
    if (e > m) m = e;
  }
  return m;
}
// runs (or, dry, only sizes) forward + backward bookkeeping; returns bytes through *total
static int train_backward_impl(ftb_unet* h, TrainState* T, TrainCtx& c, const float* dout, float* grads) {
  c.gmap.clear();
  c.dout = dout;
  c.grads = grads;
  c.dfilm = c.f32((size_t)c.B * h->film_rows);
  c.dts = c.f32((size_t)c.B * h->time_dim);
  c.wt_tmp = c.f32(max_wt_elems(h));
  c.wg_part_bytes = conv_wgrad_partial_bytes();
  c.wg_part = c.f32(c.wg_part_bytes / sizeof(float));
  FTB_TRY(c.zero(c.dfilm, (size_t)c.B * h->film_rows * sizeof(float)));
  FTB_TRY(c.zero(c.dts, (size_t)c.B * h->time_dim * sizeof(float)));
  if (!c.dry && h->dgrad_dirty) FTB_TRY(pack_dgrad(h, c.wt_tmp, c.st));
  for (size_t i = T->tape.size(); i-- > 0;) FTB_TRY(T->tape[i](c));
  return 0;
}
}  // namespace ftb_engine_detail

extern "C" {

int64_t ftb_unet3d_param_offset(ftb_unet* h, int i) {
  if (!h || i < 0 || i > (int)h->params.size()) return -1;
  int64_t off = 0;
  for (int k = 0; k < i; ++k) off += h->params[k].numel;
  return off;
}

int ftb_unet3d_bind_params(ftb_unet* h, float* flat, void* stream) {
  FTB_CHECK(h && flat, "null argument");
  FTB_TRY(ensure_device(h));
  int64_t off = 0;
  for (Param& p : h->params) {
    p.dev = flat + off;
    p.set = true;
    off += p.numel;
  }
  // the FiLM table holds parameter pointers
  const int nb = (int)h->film_blocks.size();
  std::vector<const float*> hw(nb), hb(nb);
  for (int i = 0; i < nb; ++i) {
    hw[i] = h->params[h->pindex[h->film_blocks[i] + ".weight"]].dev;
    hb[i] = h->params[h->pindex[h->film_blocks[i] + ".bias"]].dev;
  }
  FTB_CUDA(cudaMemcpyAsync(h->d_film_w, hw.data(), nb * sizeof(float*), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  FTB_CUDA(cudaMemcpyAsync(h->d_film_b, hb.data(), nb * sizeof(float*), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  FTB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  h->dirty = true;
  return 0;
}

int ftb_unet3d_set_dropout(ftb_unet* h, float p, uint64_t seed) {
  FTB_CHECK(h, "null handle");
  FTB_CHECK(p >= 0.f && p < 1.f, "dropout probability must be in [0, 1)");
  h->drop_p = p;
  h->drop_seed = seed;
  return 0;
}

int ftb_unet3d_mark_dirty(ftb_unet* h) {
  FTB_CHECK(h, "null handle");
  h->dirty = true;
  return 0;
}

size_t ftb_unet3d_train_workspace_bytes(ftb_unet* h, int B, int X, int Y, int Z) {
  if (check_dims(h, B, X, Y, Z) != 0) return 0;
  TrainState T;
  ensure_offsets(h, &T);
  TrainCtx c{h, &T, nullptr, reinterpret_cast<char*>(uintptr_t(1) << 40), 0, true, B};
  TrainFwd f{h, &T, c};
  if (f.run(nullptr, nullptr, nullptr, X, Y, Z) != 0) return 0;
  if (train_backward_impl(h, &T, c, nullptr, reinterpret_cast<float*>(uintptr_t(1) << 41)) != 0) return 0;
  return round_up_sz(c.off, 256) + 256;
}

static int forward_train_impl(ftb_unet* h, const float* x, const float* atb, const float* t, float* out, int B, int X,
                              int Y, int Z, void* workspace, size_t workspace_bytes, void* stream) {
  FTB_TRY(check_dims(h, B, X, Y, Z));
  FTB_CHECK(x && t && out && workspace, "null argument");
  FTB_CHECK(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  FTB_TRY(ensure_device(h));
  FTB_TRY(finalize(h, st, true));
  if (!h->train) h->train = std::make_shared<TrainState>();
  TrainState* T = h->train.get();
  ensure_offsets(h, T);
  T->valid = false;
  if (T->need_bytes == 0 || T->B != B || T->X != X || T->Y != Y || T->Z != Z)   // sized once per shape (a dry tape run)
    T->need_bytes = ftb_unet3d_train_workspace_bytes(h, B, X, Y, Z);
  FTB_CHECK(T->need_bytes > 0 && workspace_bytes >= T->need_bytes,
            "training workspace too small: need " + std::to_string(T->need_bytes) + " bytes");
  h->train_ctx.reset(new TrainCtx{h, T, st, reinterpret_cast<char*>(workspace), 0, false, B});
  TrainFwd f{h, T, *h->train_ctx};
  FTB_TRY(f.run(x, t, out, X, Y, Z, atb));
  FTB_CHECK(h->train_ctx->off <= workspace_bytes, "training workspace too small for the forward");
  T->fwd_bytes = h->train_ctx->off;
  T->B = B; T->X = X; T->Y = Y; T->Z = Z;
  T->valid = true;
  return 0;
}

int ftb_unet3d_forward_train(ftb_unet* h, const float* x, const float* t, float* out, int B, int X, int Y, int Z,
                             void* workspace, size_t workspace_bytes, void* stream) {
  FTB_CHECK(h, "null handle");
  FTB_CHECK(!h->cfg.conditional, "conditional model: call ftb_unet3d_cond_forward_train (ATb is required)");
  return forward_train_impl(h, x, nullptr, t, out, B, X, Y, Z, workspace, workspace_bytes, stream);
}

int ftb_unet3d_cond_forward_train(ftb_unet* h, const float* x, const float* atb, const float* t, float* out, int B,
                                  int X, int Y, int Z, void* workspace, size_t workspace_bytes, void* stream) {
  FTB_CHECK(h && atb, "null argument");
  FTB_CHECK(h->cfg.conditional, "unconditional model: call ftb_unet3d_forward_train");
  return forward_train_impl(h, x, atb, t, out, B, X, Y, Z, workspace, workspace_bytes, stream);
}

int ftb_unet3d_backward(ftb_unet* h, const float* dout, float* grads, void* workspace, size_t workspace_bytes,
                        void (*bucket_cb)(void*, int64_t, int64_t), void* cb_user, void* stream) {
  FTB_CHECK(h && dout && grads && workspace, "null argument");
  FTB_CHECK(h->train && h->train->valid && h->train_ctx, "backward: no matching forward_train");
  TrainState* T = h->train.get();
  TrainCtx& c = *h->train_ctx;
  FTB_CHECK(c.base == reinterpret_cast<char*>(workspace), "backward: workspace differs from the forward's");
  FTB_CHECK(workspace_bytes >= T->need_bytes, "training workspace too small: need " + std::to_string(T->need_bytes) + " bytes");
  c.st = (cudaStream_t)stream;
  c.off = T->fwd_bytes;
  c.cb = bucket_cb;
  c.cb_user = cb_user;
  T->valid = false;   // one backward per forward
  return train_backward_impl(h, T, c, dout, grads);
}

int ftb_mse_ratio_grad(const float* v, const float* vhat, int64_t n, const double* acc2, float scale, float* dout,
                       void* stream) {
  FTB_CHECK(v && vhat && acc2 && dout, "null argument");
  return mse_ratio_grad(v, vhat, n, acc2, scale, dout, (cudaStream_t)stream);
}
int ftb_grad_sumsq(const float* g, int64_t n, double* acc, void* stream) {
  FTB_CHECK(g && acc, "null argument");
  return grad_sumsq(g, n, acc, (cudaStream_t)stream);
}
int ftb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int decoupled, int step, const double* sumsq, float grad_scale,
                  float max_norm, void* stream) {
  FTB_CHECK(p && g && m && v && step >= 1, "null argument / step < 1");
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, decoupled, step, sumsq, grad_scale, max_norm,
                   (cudaStream_t)stream);
}

}  // extern "C"

// ---------------------------------------------------------------- test hooks
namespace ftb_engine_detail {
struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch() {
    for (void* p : ptrs) cudaFree(p);
  }
  template <typename T>
  T* get(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T) + 16) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return reinterpret_cast<T*>(p);
  }
};
}  // namespace

extern "C" {

int ftb_test_conv3d(const float* x, int c1, const float* x2, int c2, const float* w, const float* bias,
                    int cout, int ksize, const float* g, const float* scale, const float* shift,
                    const float* resid, int flags, float* out, int B, int X, int Y, int Z, int impl,
                    void* stream) {
  FTB_CHECK(x && w && out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch S;
  auto mk = [&](int C) {
    Act a;
    a.B = B; a.C = round_up(C, 16); a.D = X; a.H = Y; a.W = Z;
    a.p = S.get<bf16>(a.elems());
    return a;
  };
  // like the engine, a wide-kernel / small-Cin conv runs W-unfolded (impl 2 forces the cubic kernel)
  const bool unfold = !x2 && ksize >= 5 && ksize * c1 <= 128 && round_up(ksize * c1, 16) < ksize * round_up(c1, 16) &&
                      impl != 2 && getenv("FTB_NO_UNFOLD") == nullptr;
  Act a0 = mk(unfold ? ksize * c1 : c1), a1, ar, ao = mk(cout);
  FTB_CHECK(a0.p && ao.p, "scratch allocation failed");
  if (unfold) FTB_TRY(pack_unfold_w(x, B, c1, X, Y, Z, ksize, a0, st));
  else FTB_TRY(pack_ncdhw_to_blocked(x, B, c1, X, Y, Z, a0, st));
  if (x2) {
    FTB_CHECK(c1 % 16 == 0, "concat test needs c1 % 16 == 0");
    a1 = mk(c2);
    FTB_CHECK(a1.p, "scratch allocation failed");
    FTB_TRY(pack_ncdhw_to_blocked(x2, B, c2, X, Y, Z, a1, st));
  }
  const int cin = c1 + (x2 ? c2 : 0);
  const int cin_pad = a0.C + (x2 ? a1.C : 0);
  const int n_tile = round_up(cout, 16) > 256 ? 128 : round_up(cout, 16);
  const int ntiles = cdiv(cout, n_tile);
  const int taps = ksize * ksize * (unfold ? 1 : ksize);
  bf16* packed = S.get<bf16>((size_t)ntiles * taps * cin_pad * n_tile);
  float* bias_p = S.get<float>((size_t)ntiles * n_tile);
  float* gs = S.get<float>((size_t)n_tile);
  FTB_CHECK(packed && bias_p && gs, "scratch allocation failed");
  FTB_TRY(pack_conv_weights(w, cout, cin, ksize, cin_pad, n_tile, ntiles, nullptr, packed, st, unfold));
  FTB_CUDA(cudaMemsetAsync(bias_p, 0, (size_t)ntiles * n_tile * sizeof(float), st));
  if (bias) FTB_CUDA(cudaMemcpyAsync(bias_p, bias, cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ConvWeights cw;
  cw.w = packed; cw.ksize = ksize; cw.cin = cin_pad; cw.n = n_tile; cw.ntiles = ntiles;
  cw.ksize_w = unfold ? 1 : 0;
  ConvEpilogue e;
  if (bias) e.bias = bias_p;
  float *mul_p = nullptr, *add_p = nullptr;
  const int pstride = n_tile * ntiles;
  if (g) FTB_CHECK(ntiles == 1, "norm needs a single N tile");
  if (g || (scale && shift)) {
    // mul[b][c] = g[c]*sqrt(C) * (scale[b][c] + 1), add[b][c] = shift[b][c]
    mul_p = S.get<float>((size_t)B * pstride);
    add_p = S.get<float>((size_t)B * pstride);
    FTB_CHECK(mul_p && add_p, "scratch allocation failed");
    FTB_CUDA(cudaMemsetAsync(mul_p, 0, (size_t)B * pstride * sizeof(float), st));
    FTB_CUDA(cudaMemsetAsync(add_p, 0, (size_t)B * pstride * sizeof(float), st));
    test_affine_kernel<<<cdiv(B * cout, 128), 128, 0, st>>>(g, scale, shift, sqrtf((float)cout), B, cout,
                                                            pstride, mul_p, add_p);
    e.norm = g != nullptr;
    e.mul = mul_p; e.add = add_p; e.mul_stride = pstride; e.add_stride = pstride;
  }
  e.silu = (flags & 1) != 0;
  e.prenorm = (flags & 2) != 0;
  if (resid) {
    ar = mk(cout);
    FTB_CHECK(ar.p, "scratch allocation failed");
    FTB_TRY(pack_ncdhw_to_blocked(resid, B, cout, X, Y, Z, ar, st));
    e.resid = &ar;
  }
  ConvSrc s0{&a0, 0, a0.cg()}, s1{};
  if (x2) s1 = ConvSrc{&a1, 0, a1.cg()};
  if (impl == 1) FTB_TRY(conv_naive(s0, s1, cw, e, ao, 0, st));   // impl 2: tcgen05 kernel, cubic (no W-unfold)
  else FTB_TRY(conv_igemm(s0, s1, cw, e, ao, 0, st));
  FTB_TRY(unpack_blocked_to_ncdhw(ao, 0, cout, out, st));
  FTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// dw [cout][c1+c2][k^3] = conv3d weight gradient from x (|| x2) and dy (NCDHW fp32, rounded to bf16 inside)
int ftb_test_conv_wgrad(const float* x, int c1, const float* x2, int c2, const float* dy, int cout, int ksize,
                        float* dw, int B, int X, int Y, int Z, int unfold, void* stream) {
  FTB_CHECK(x && dy && dw, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch S;
  auto mk = [&](int C) {
    Act a;
    a.B = B; a.C = round_up(C, 16); a.D = X; a.H = Y; a.W = Z;
    a.p = S.get<bf16>(a.elems());
    return a;
  };
  Act a0 = mk(unfold ? ksize * c1 : c1), a1, ady = mk(cout);
  FTB_CHECK(a0.p && ady.p, "scratch allocation failed");
  if (unfold) FTB_TRY(pack_unfold_w(x, B, c1, X, Y, Z, ksize, a0, st));
  else FTB_TRY(pack_ncdhw_to_blocked(x, B, c1, X, Y, Z, a0, st));
  FTB_TRY(pack_ncdhw_to_blocked(dy, B, cout, X, Y, Z, ady, st));
  const int cin = c1 + (x2 ? c2 : 0);
  FTB_CUDA(cudaMemsetAsync(dw, 0, (size_t)cout * cin * ksize * ksize * ksize * sizeof(float), st));
  const size_t pbytes = conv_wgrad_partial_bytes();
  float* part = S.get<float>(pbytes / sizeof(float));
  FTB_CHECK(part, "scratch allocation failed");
  FTB_TRY(conv_wgrad(a0, 0, a0.cg(), ady, 0, cout, ksize, unfold ? c1 : 0, dw, cin, 0, c1, 0, st, part, pbytes));
  if (x2) {
    a1 = mk(c2);
    FTB_CHECK(a1.p, "scratch allocation failed");
    FTB_TRY(pack_ncdhw_to_blocked(x2, B, c2, X, Y, Z, a1, st));
    FTB_TRY(conv_wgrad(a1, 0, a1.cg(), ady, 0, cout, ksize, 0, dw, cin, c1, c2, 0, st, part, pbytes));
  }
  FTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// dx [B][cin][...] (+ acc) = conv3d input gradient of dy through w [cout][cin][k^3]: the forward kernel on
// flipped, transposed weights
int ftb_test_conv_dgrad(const float* dy, const float* w, int cout, int cin, int ksize, const float* acc, float* dx,
                        int B, int X, int Y, int Z, void* stream) {
  FTB_CHECK(dy && w && dx, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch S;
  auto mk = [&](int C) {
    Act a;
    a.B = B; a.C = round_up(C, 16); a.D = X; a.H = Y; a.W = Z;
    a.p = S.get<bf16>(a.elems());
    return a;
  };
  Act ady = mk(cout), adx = mk(cin);
  FTB_CHECK(ady.p && adx.p, "scratch allocation failed");
  FTB_TRY(pack_ncdhw_to_blocked(dy, B, cout, X, Y, Z, ady, st));
  if (acc) FTB_TRY(pack_ncdhw_to_blocked(acc, B, cin, X, Y, Z, adx, st));
  const int k3 = ksize * ksize * ksize, n_tile = round_up(cin, 16), kpad = round_up(cout, 16);
  float* wt = S.get<float>((size_t)cin * cout * k3);
  bf16* packed = S.get<bf16>((size_t)k3 * kpad * n_tile);
  FTB_CHECK(wt && packed, "scratch allocation failed");
  FTB_TRY(transpose_flip(w, cout, cin, ksize, 0, cin, nullptr, wt, st));
  FTB_TRY(pack_conv_weights(wt, cin, cout, ksize, kpad, n_tile, 1, nullptr, packed, st));
  ConvWeights cw;
  cw.w = packed; cw.ksize = ksize; cw.cin = kpad; cw.n = n_tile; cw.ntiles = 1;
  ConvEpilogue e;
  if (acc) e.resid = &adx;
  FTB_TRY(conv_igemm(ConvSrc{&ady, 0, ady.cg()}, ConvSrc{}, cw, e, adx, 0, st));
  FTB_TRY(unpack_blocked_to_ncdhw(adx, 0, cin, dx, st));
  FTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int ftb_test_conv_debug(long long* host, int n) { return conv_debug_read(host, n); }

int ftb_test_trilinear(const float* x, int B, int C, int X, int Y, int Z, int Xo, int Yo, int Zo,
                       float* out, void* stream) {
  FTB_CHECK(x && out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch S;
  Act a, o;
  a.B = o.B = B; a.C = o.C = round_up(C, 16);
  a.D = X; a.H = Y; a.W = Z; o.D = Xo; o.H = Yo; o.W = Zo;
  a.p = S.get<bf16>(a.elems());
  o.p = S.get<bf16>(o.elems());
  FTB_CHECK(a.p && o.p, "scratch allocation failed");
  FTB_TRY(pack_ncdhw_to_blocked(x, B, C, X, Y, Z, a, st));
  FTB_TRY(trilinear_resample(a, o, st));
  FTB_TRY(unpack_blocked_to_ncdhw(o, 0, C, out, st));
  FTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

/* adjoint of the trilinear resample: dout [B,C,Xo,Yo,Zo] -> din [B,C,X,Y,Z] (+= acc when given) */
int ftb_test_trilinear_bwd(const float* dout, int B, int C, int X, int Y, int Z, int Xo, int Yo, int Zo,
                           const float* acc, float* din, void* stream) {
  FTB_CHECK(dout && din, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Scratch S;
  Act a, o;
  a.B = o.B = B; a.C = o.C = round_up(C, 16);
  a.D = X; a.H = Y; a.W = Z; o.D = Xo; o.H = Yo; o.W = Zo;
  a.p = S.get<bf16>(a.elems());
  o.p = S.get<bf16>(o.elems());
  FTB_CHECK(a.p && o.p, "scratch allocation failed");
  FTB_TRY(pack_ncdhw_to_blocked(dout, B, C, Xo, Yo, Zo, o, st));
  if (acc) FTB_TRY(pack_ncdhw_to_blocked(acc, B, C, X, Y, Z, a, st));
  FTB_TRY(trilinear_resample_bwd(o, a, acc != nullptr, st));
  FTB_TRY(unpack_blocked_to_ncdhw(a, 0, C, din, st));
  FTB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
