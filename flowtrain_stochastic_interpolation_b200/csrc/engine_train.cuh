// Training path of the Unet3D engine: forward with saved activations + backward (included by engine.cu).
//
// The training step (reference Geo3DStochInterp.training_step,
// project/geodata-3d-unconditional/model_train_inference.py:417-457, autograd through
// Unet3D.forward, src/flowtrain/models/unet_attn_3d.py:673-719) needs the intermediates the fused
// inference kernels never write.  The train-mode forward therefore runs every conv with a plain
// (bias-only) epilogue and the RMSNorm / FiLM / SiLU / residual as a separate pass, keeps every
// tensor in the (bump-allocated) workspace, and records one backward closure per op on a tape.
// backward() replays the tape in reverse: data gradients are the same implicit-GEMM conv kernel
// with flipped, transposed weights (conv_igemm.cu), weight gradients the voxel-contraction GEMM
// of wgrad.cu, everything else the bandwidth kernels of train_ops.cu.  Gradients of activations
// are blocked bf16 like the activations; parameter gradients are fp32, accumulated (+=) into the
// caller's flat buffer in state_dict order.
#pragma once
#include <functional>

namespace ftb_engine_detail {

struct TrainCtx;
typedef std::function<int(TrainCtx&)> BwdFn;

struct TrainState {
  std::vector<BwdFn> tape;
  std::vector<int64_t> poff;                              // flat offset of each parameter
  int64_t ptotal = 0;
  size_t fwd_bytes = 0;      // workspace consumed by the forward (the backward continues after it)
  size_t need_bytes = 0;     // forward + backward footprint of the current shape
  int B = 0, X = 0, Y = 0, Z = 0;
  bool valid = false;        // a forward_train has run and its tape matches the workspace
  const float* t_dev = nullptr;
};

typedef void (*ftb_bucket_cb)(void* user, int64_t offset, int64_t count);

struct TrainCtx {
  ftb_unet* U;
  TrainState* T;
  cudaStream_t st;
  char* base;
  size_t off = 0;
  bool dry;
  int B;
  float* grads = nullptr;          // flat parameter gradients (state_dict order)
  const float* dout = nullptr;     // dL/d(output) NCDHW fp32
  float* film_raw = nullptr;       // [B][film_rows]: (scale+1 | shift) per block
  float* film_fold = nullptr;      // [B][film_rows]: ((scale+1)*g*sqrt(C) | shift), the fused conv epilogue's operands
  float* dfilm = nullptr;          // [B][film_rows]
  float* temb_silu = nullptr;      // [B][time_dim]
  float* dts = nullptr;            // [B][time_dim] gradient w.r.t. silu(temb)
  float* wt_tmp = nullptr;         // transposed fp32 weights scratch
  float* wg_part = nullptr;        // per-CTA partial weight gradients of the running conv_wgrad launch
  size_t wg_part_bytes = 0;
  ftb_bucket_cb cb = nullptr;
  void* cb_user = nullptr;
  struct GradSlot { Act g; bool init = false; };
  std::map<const void*, GradSlot> gmap;

  void* raw(size_t bytes) {
    off = round_up_sz(off, 256);
    void* p = base + off;   // dry: base is a fake non-null address that is never dereferenced
    off += bytes;
    return p;
  }
  Act act(int C, int D, int H, int W, int batch = 0) {
    Act a;
    a.B = batch > 0 ? batch : B; a.C = round_up(C, 16); a.D = D; a.H = H; a.W = W;
    a.p = reinterpret_cast<bf16*>(raw(a.bytes()));
    return a;
  }
  Act like(const Act& x) { return act(x.C, x.D, x.H, x.W, x.B); }
  float* f32(size_t n) { return reinterpret_cast<float*>(raw(n * sizeof(float))); }
  int zero(void* p, size_t bytes) {
    if (!dry) FTB_CUDA(cudaMemsetAsync(p, 0, bytes, st));
    return 0;
  }
  const float* pdev(const std::string& n) { return U->params[U->pindex.at(n)].dev; }
  float* gptr(const std::string& n) { return grads + T->poff[U->pindex.at(n)]; }
  GradSlot& grad(const Act& x) {
    auto it = gmap.find(x.p);
    if (it == gmap.end()) {
      GradSlot s;
      s.g = like(x);
      it = gmap.emplace(x.p, s).first;
    }
    return it->second;
  }
  // G(x) (+)= src.  `src` is always the gradient of the calling closure's own output, which nobody reads after that
  // closure: when it is the FIRST contribution to G(x) the slot simply takes over that buffer (no copy pass); later
  // contributions add into it in place, after (stream order) the kernels that read it as `src`.
  int accumulate(const Act& x, const Act& src) {
    GradSlot& s = grad(x);
    if (!s.init && alias_ok) {
      s.g = src;
    } else if (!dry) {
      FTB_TRY(act_accum(s.g, src, s.init, st));
    }
    s.init = true;
    return 0;
  }
  bool alias_ok = getenv("FTB_TRAIN_NO_ALIAS") == nullptr;
  int need(const Act& x, Act* g, const char* what) {
    auto it = gmap.find(x.p);
    FTB_CHECK(it != gmap.end() && it->second.init, std::string("backward: no gradient reached ") + what);
    *g = it->second.g;
    return 0;
  }
};

#define TRUN(expr) do { if (!c.dry) FTB_TRY(expr); } while (0)

// ------------------------------------------------------------------ forward (train mode)
struct TrainFwd {
  ftb_unet* U;
  TrainState* T;
  TrainCtx& c;       // allocator shared with the backward
  bool fuse = getenv("FTB_TRAIN_NOFUSE") == nullptr;   // norm/FiLM/SiLU in the conv epilogue (pre-norm side output)

  ConvWeights weights(const ConvLayer& cl) {
    ConvWeights w;
    w.w = cl.packed; w.ksize = cl.k; w.cin = cl.cin_pad; w.n = cl.n_tile; w.ntiles = cl.ntiles;
    w.ksize_w = cl.unfold_w ? 1 : 0;
    w.cin_real = cl.unfold_w ? cl.k * cl.cin : cl.cin; w.cout_real = cl.cout;
    return w;
  }

  // transposed / flipped packs for the data gradient of `name`, one per source (channel split)
  int ensure_dgrad(const std::string& name, const std::vector<int>& split) {
    const ConvLayer& cl = U->convs.at(name);
    if (c.dry) return 0;   // sizing pass: nothing to allocate
    auto& v = U->dgrad[name];
    if (v.empty()) {
      int ci0 = 0;
      for (int cs : split) {
        DgradPack d;
        d.ci0 = ci0; d.cin_sub = cs; d.n_tile = round_up(cs, 16);
        FTB_CHECK(d.n_tile <= 256, "dgrad: more than 256 channels in one source");
        const size_t elems = (size_t)cl.k * cl.k * cl.k * round_up(cl.cout, 16) * d.n_tile;
        FTB_TRY(dev_alloc(U, &d.packed, elems));
        v.push_back(d);
        ci0 += cs;
      }
      U->dgrad_dirty = true;
    }
    return 0;
  }

  // out = conv(x0 [|| x1]) + bias (+ resid) (q softmax on the first N tile when qsm); out_f32: NCDHW fp32 output
  int conv(const std::string& name, const Act& x0, const Act* x1, int c0_real, int c1_real, const Act* resid, bool qsm,
           float* out_f32, Act& out, bool leaf, bool f32out = false, bool bias_by_consumer = false,
           const ConvEpilogue* fused = nullptr, Act* fused_out = nullptr) {
    const ConvLayer& cl = U->convs.at(name);
    ConvWeights w = weights(cl);
    ConvEpilogue e;
    if (!cl.bname.empty()) e.bias = cl.bias;
    e.resid = resid;
    if (qsm) {
      e.q_softmax_heads = U->cfg.attn_heads; e.q_dim_head = U->cfg.attn_dim_head;
      e.q_scale = 1.f / sqrtf((float)U->cfg.attn_dim_head);
    }
    if (f32out) { e.out_f32 = out_f32; e.out_f32_c = cl.cout; }
    ConvSrc s0{&x0, 0, x0.cg()}, s1{};
    if (x1) s1 = ConvSrc{x1, 0, x1->cg()};
    if (!leaf) {
      std::vector<int> split{c0_real};
      if (x1) split.push_back(c1_real);
      FTB_TRY(ensure_dgrad(name, split));
    }
    if (fused) {
      // norm / FiLM / SiLU / dropout / residual fused into this conv's epilogue (as on the inference path); the
      // pre-norm tensor `out` that the backward needs is a side output of the same launch
      ConvEpilogue ef = *fused;
      ef.bias = e.bias;
      ef.pre_out = &out;
      TRUN(conv_dispatch(s0, s1, w, ef, *fused_out, 0, c.st));
    } else {
      TRUN(conv_dispatch(s0, s1, w, e, out, 0, c.st));
    }
    // ---- backward closure
    const Act X0 = x0, X1 = x1 ? *x1 : Act(), R = resid ? *resid : Act(), O = out;
    const bool has1 = x1 != nullptr, hasr = resid != nullptr;
    T->tape.push_back([=](TrainCtx& c) -> int {
      ftb_unet* U = c.U;
      const ConvLayer& cl = U->convs.at(name);
      Act dY;
      if (f32out) {
        dY = c.act(cl.cout, O.D, O.H, O.W);
        TRUN(pack_ncdhw_to_blocked(c.dout, c.B, cl.cout, O.D, O.H, O.W, dY, c.st));
      } else {
        FTB_TRY(c.need(O, &dY, name.c_str()));
      }
      if (hasr) FTB_TRY(c.accumulate(R, dY));
      // (when a norm/activation pass consumes this conv's output, its backward already summed dY per channel)
      if (!cl.bname.empty() && !bias_by_consumer) TRUN(bias_grad(dY, 0, cl.cout, c.gptr(cl.bname), c.st));
      // weight gradient
      const int cin_tot = cl.cin;
      float* dw = c.gptr(cl.wname);
      float* dw_target = dw;
      const size_t wn = (size_t)cl.cout * cl.cin * cl.k * cl.k * cl.k;
      if (!cl.in_scale.empty()) {
        dw_target = c.f32(wn);
        FTB_TRY(c.zero(dw_target, wn * sizeof(float)));
      }
      if (cl.unfold_w) {
        TRUN(conv_wgrad(X0, 0, X0.cg(), dY, 0, cl.cout, cl.k, cl.cin, dw_target, cin_tot, 0, cl.cin, 0, c.st, c.wg_part, c.wg_part_bytes));
      } else {
        TRUN(conv_wgrad(X0, 0, X0.cg(), dY, 0, cl.cout, cl.k, 0, dw_target, cin_tot, 0, c0_real, 0, c.st, c.wg_part, c.wg_part_bytes));
        if (has1) TRUN(conv_wgrad(X1, 0, X1.cg(), dY, 0, cl.cout, cl.k, 0, dw_target, cin_tot, c0_real, c1_real, 0, c.st, c.wg_part, c.wg_part_bytes));
      }
      if (!cl.in_scale.empty())
        TRUN(fold_gain_bwd(dw_target, c.pdev(cl.wname), cl.scale_tmp, cl.cout, cl.cin, cl.in_scale_mul, dw,
                           c.gptr(cl.in_scale), c.st));
      // data gradient: the same implicit-GEMM kernel with flipped, transposed weights
      if (!leaf) {
        for (int s = 0; s < (has1 ? 2 : 1); ++s) {
          const Act& Xs = s == 0 ? X0 : X1;
          TrainCtx::GradSlot& gs = c.grad(Xs);
          if (!c.dry) {
            const DgradPack& dp = c.U->dgrad.at(name).at(s);
            ConvWeights w;
            w.w = dp.packed; w.ksize = cl.k; w.cin = round_up(cl.cout, 16); w.n = dp.n_tile; w.ntiles = 1;
            w.cin_real = cl.cout; w.cout_real = dp.cin_sub;
            ConvEpilogue e;
            if (gs.init) e.resid = &gs.g;
            FTB_TRY(conv_dispatch(ConvSrc{&dY, 0, dY.cg()}, ConvSrc{}, w, e, gs.g, 0, c.st));
          }
          gs.init = true;
        }
      }
      return 0;
    });
    return 0;
  }

  // explicit rows of the FiLM table (MixATb: the (scale | shift) halves of x and of the ATb embedding); the caller
  // pushes the backward of the block's Linear itself
  struct FilmSlice { int s1 = -1, sh = -1; };

  // out = act(norm(u) * gain * s1 + sh) + resid
  int normact(const Act& u, bool norm, const std::string& gain_name, const std::string& film_block, bool silu,
              const Act* resid, Act& out, const std::string& bias_of_u = "", bool forward_done = false,
              const FilmSlice* fs = nullptr) {
    const int C = u.C;
    const float* gain = gain_name.empty() ? nullptr : U->gains.at(gain_name).gs;
    const int foff = fs ? fs->s1 : (film_block.empty() ? -1 : U->film_off.at(film_block));
    const int shoff = fs ? fs->sh : foff + C;
    const bool own_linear = foff >= 0 && fs == nullptr;
    const float* s1 = foff >= 0 ? c.film_raw + foff : nullptr;
    const float* sh = foff >= 0 ? c.film_raw + shoff : nullptr;
    const int fstride = U->film_rows;
    // nn.Dropout sits at the end of Block.forward (:244) and only block1 gets p > 0 (:261): the FiLM'd block
    const float dp = own_linear ? U->drop_p : 0.f;
    const unsigned long long dkey = U->drop_seed * 0x9E3779B97F4A7C15ull + ((unsigned long long)(foff + 1) << 40);
    if (!forward_done) TRUN(normact_fwd(u, norm, gain, s1, sh, fstride, silu, resid, out, c.st, dp, dkey));
    const Act Uu = u, O = out, R = resid ? *resid : Act();
    const bool hasr = resid != nullptr;
    T->tape.push_back([=](TrainCtx& c) -> int {
      Act dO;
      FTB_TRY(c.need(O, &dO, gain_name.empty() ? "normact" : gain_name.c_str()));
      if (hasr) FTB_TRY(c.accumulate(R, dO));
      float* Rb = c.f32((size_t)c.B * C);
      FTB_TRY(c.zero(Rb, (size_t)c.B * C * sizeof(float)));
      TrainCtx::GradSlot& gu = c.grad(Uu);
      float* S = foff >= 0 ? c.dfilm + shoff : nullptr;
      float* dbias = bias_of_u.empty() ? nullptr : c.gptr(bias_of_u);   // bias gradient of the conv that produced u
      if (gu.init) {   // the input also feeds a residual path: add to its gradient
        Act tmp = c.like(Uu);
        TRUN(normact_bwd(dO, Uu, norm, gain, s1, sh, fstride, silu, tmp, Rb, S, fstride, dbias, c.st, dp, dkey));
        TRUN(act_accum(gu.g, tmp, true, c.st));
      } else {
        TRUN(normact_bwd(dO, Uu, norm, gain, s1, sh, fstride, silu, gu.g, Rb, S, fstride, dbias, c.st, dp, dkey));
      }
      gu.init = true;
      float* ds1 = foff >= 0 ? c.dfilm + foff : nullptr;
      float* dg = gain_name.empty() ? nullptr : c.gptr(gain_name);
      if (ds1 || dg) TRUN(normact_finish(Rb, c.B, C, gain, s1, fstride, sqrtf((float)C), ds1, dg, c.st));
      if (own_linear) {
        // this block's FiLM Linear (SiLU -> Linear(time_dim -> 2C)): dW, db now, d silu(temb) accumulated
        const int td = c.U->time_dim;
        TRUN(linear_bwd(c.dfilm + foff, fstride, c.temb_silu, c.pdev(film_block + ".weight"), c.B, 2 * C, td,
                        c.gptr(film_block + ".weight"), c.gptr(film_block + ".bias"), c.dts, true, c.st));
      }
      return 0;
    });
    return 0;
  }

  int resample(const Act& in, int D, int H, int W, Act* out) {
    *out = c.act(in.C, D, H, W);
    TRUN(trilinear_resample(in, *out, c.st));
    const Act I = in, O = *out;
    T->tape.push_back([=](TrainCtx& c) -> int {
      Act dO;
      FTB_TRY(c.need(O, &dO, "resample"));
      TrainCtx::GradSlot& gi = c.grad(I);
      TRUN(trilinear_resample_bwd(dO, gi.g, gi.init, c.st));
      gi.init = true;
      return 0;
    });
    return 0;
  }

  // epilogue equivalent of normact(norm, gain, film, silu, resid) for the fused train-mode conv
  ConvEpilogue fused_epilogue(const std::string& gain_name, const std::string& film_block, int C, bool silu,
                              const Act* resid) {
    ConvEpilogue e;
    e.norm = true;
    if (!film_block.empty()) {
      const int foff = U->film_off.at(film_block);
      e.mul = c.film_fold + foff; e.mul_stride = U->film_rows;           // (scale + 1) * g * sqrt(C)
      e.add = c.film_fold + foff + C; e.add_stride = U->film_rows;
      e.drop_p = U->drop_p;
      e.drop_key = U->drop_seed * 0x9E3779B97F4A7C15ull + ((unsigned long long)(foff + 1) << 40);
    } else {
      e.mul = U->gains.at(gain_name).gs;
    }
    e.silu = silu;
    e.resid = resid;
    return e;
  }
  // (a layer cut into several N tiles cannot norm over all channels in its epilogue)
  bool fuse_ok(const std::string& conv_name, int cout) const {
    return fuse && round_up(cout, 16) <= 256 && U->convs.at(conv_name).ntiles == 1;
  }

  // ResnetBlock (:265-278)
  int resnet(const std::string& p, const Act& x0, const Act* x1, int c0, int c1, int cout, Act* out) {
    const std::string film = resnet_mlp(U, p);
    Act u1 = c.act(cout, x0.D, x0.H, x0.W), h1 = c.like(u1);
    const bool fz = fuse_ok(p + ".block1.proj", cout);
    ConvEpilogue f1 = fused_epilogue(p + ".block1.norm.g", film, cout, true, nullptr);
    FTB_TRY(conv(p + ".block1.proj", x0, x1, c0, c1, nullptr, false, nullptr, u1, false, false, true, fz ? &f1 : nullptr, &h1));
    FTB_TRY(normact(u1, true, p + ".block1.norm.g", film, true, nullptr, h1, p + ".block1.proj.bias", fz));
    Act res = x0;
    const int cin = c0 + (x1 ? c1 : 0);
    if (cin != cout) {
      res = c.like(u1);
      FTB_TRY(conv(p + ".res_conv", x0, x1, c0, c1, nullptr, false, nullptr, res, false));
    }
    Act u2 = c.like(u1);
    *out = c.like(u1);
    ConvEpilogue f2 = fused_epilogue(p + ".block2.norm.g", "", cout, true, &res);
    FTB_TRY(conv(p + ".block2.proj", h1, nullptr, cout, 0, nullptr, false, nullptr, u2, false, false, true, fz ? &f2 : nullptr, out));
    FTB_TRY(normact(u2, true, p + ".block2.norm.g", "", true, &res, *out, p + ".block2.proj.bias", fz));
    return 0;
  }

  // x + attn(x)
  int attention(const std::string& p, const Act& x, bool full, Act* out) {
    const ftb_unet_cfg& cf = U->cfg;
    const int heads = cf.attn_heads, dh = cf.attn_dim_head, hd = heads * dh, nm = cf.num_mem_kv;
    const int C = x.C;
    Act xh = c.like(x);
    FTB_TRY(normact(x, true, "", "", false, nullptr, xh));
    Act qkv = c.act(3 * hd, x.D, x.H, x.W);
    FTB_TRY(conv(p + ".to_qkv", xh, nullptr, C, 0, nullptr, !full, nullptr, qkv, false));
    *out = c.like(x);
    if (full) {
      Act ao = c.act(hd, x.D, x.H, x.W);
      TRUN(full_attention(qkv, heads, dh, c.pdev(p + ".mem_kv"), nm, ao, c.st));
      const Act Q = qkv, AO = ao;
      T->tape.push_back([=](TrainCtx& c) -> int {
        Act dao;
        FTB_TRY(c.need(AO, &dao, "attention output"));
        const int n = (int)Q.voxels();
        float* scratch = c.f32((size_t)2 * c.B * heads * n * (n + nm));
        TrainCtx::GradSlot& gq = c.grad(Q);
        FTB_CHECK(!gq.init, "attention backward: qkv already has a gradient");
        TRUN(full_attention_bwd(Q, AO, dao, heads, dh, c.pdev(p + ".mem_kv"), nm, scratch, gq.g, c.gptr(p + ".mem_kv"), c.st));
        gq.init = true;
        return 0;
      });
      FTB_TRY(conv(p + ".to_out", ao, nullptr, hd, 0, &x, false, nullptr, *out, false));
      return 0;
    }
    // ---- LinearAttention (:308-341), unfused: k softmax statistics -> context -> o = ctx^T q -> to_out -> norm
    FTB_CHECK(hd % 16 == 0 && hd <= 256, "linear attention: heads*dim_head");
    const size_t vox = x.voxels();
    int nsplit = cdiv(2 * num_sms(), c.B);
    const int max_split = (int)((vox + 511) / 512);
    nsplit = nsplit > max_split ? max_split : nsplit;
    nsplit = nsplit < 1 ? 1 : (nsplit > 256 ? 256 : nsplit);
    float* kmax = c.f32((size_t)c.B * hd);
    float* part = c.f32((size_t)c.B * heads * nsplit * (dh * dh + dh));
    float* ctx = c.f32((size_t)c.B * heads * dh * dh);
    float* kstat = c.f32((size_t)c.B * hd * 2);
    bf16* mpack = reinterpret_cast<bf16*>(c.raw((size_t)c.B * C * hd * sizeof(bf16)));
    float* dense = c.f32((size_t)c.B * hd * hd);
    bf16* wf = reinterpret_cast<bf16*>(c.raw((size_t)c.B * hd * hd * sizeof(bf16)));
    TRUN(linattn_kmax(qkv, heads, dh, nsplit, kmax, c.st));
    TRUN(linattn_context_partial(qkv, heads, dh, nsplit, kmax, part, c.st));
    TRUN(linattn_combine(part, nsplit, kmax, hd, c.B, heads, dh, c.pdev(p + ".mem_kv"), nm, c.pdev(p + ".to_out.0.weight"),
                         C, 1.f, mpack, ctx, c.st, kstat));
    TRUN(ksoftmax_apply(qkv, hd, kstat, c.st));   // k third now holds softmax_n(k)
    // o[(h,e)] = sum_d ctx[h][d][e] q~[(h,d)]: per-sample 1x1 conv with block-diagonal weights
    TRUN(blockdiag(ctx, c.B, heads, dh, true, 1.f, dense, c.st));
    TRUN(pack_conv_weights(dense, c.B * hd, hd, 1, hd, hd, c.B, nullptr, wf, c.st));
    Act o = c.act(hd, x.D, x.H, x.W);
    {
      ConvWeights w;
      w.w = wf; w.ksize = 1; w.cin = hd; w.n = hd; w.ntiles = 1; w.batch_stride = (long long)hd * hd;
      TRUN(conv_dispatch(ConvSrc{&qkv, 0, hd / 8}, ConvSrc{}, w, ConvEpilogue{}, o, 0, c.st));
    }
    const Act Q = qkv, O = o;
    T->tape.push_back([=](TrainCtx& c) -> int {
      Act dO;
      FTB_TRY(c.need(O, &dO, "linear attention o"));
      TrainCtx::GradSlot& gq = c.grad(Q);
      FTB_CHECK(!gq.init, "linear attention backward: qkv already has a gradient");
      float* full = c.f32((size_t)c.B * hd * hd);
      float* dctx = c.f32((size_t)c.B * heads * dh * dh);
      float* ssum = c.f32((size_t)c.B * hd);
      float* dn = c.f32((size_t)c.B * hd * hd);
      bf16* wb = reinterpret_cast<bf16*>(c.raw((size_t)3 * c.B * hd * hd * sizeof(bf16)));
      const size_t ws = (size_t)c.B * hd * hd;
      FTB_TRY(c.zero(full, ws * sizeof(float)));
      // dctx[b][(h,d)][(h',e)] = sum_n q~[(h,d),n] do[(h',e),n]  (voxel-contraction GEMM, one slab per sample)
      TRUN(conv_wgrad(dO, 0, hd / 8, Q, 0, hd, 1, 0, full, hd, 0, hd, (long long)hd * hd, c.st, c.wg_part, c.wg_part_bytes));
      TRUN(dctx_extract(full, ctx, c.B, heads, dh, dctx, ssum, c.st));
      ConvWeights w;
      w.ksize = 1; w.cin = hd; w.n = hd; w.ntiles = 1; w.batch_stride = (long long)hd * hd;
      // dq~[(h,d)] = sum_e ctx[h][d][e] do[(h,e)]
      TRUN(blockdiag(ctx, c.B, heads, dh, false, 1.f, dn, c.st));
      TRUN(pack_conv_weights(dn, c.B * hd, hd, 1, hd, hd, c.B, nullptr, wb, c.st));
      w.w = wb;
      TRUN(conv_dispatch(ConvSrc{&dO, 0, hd / 8}, ConvSrc{}, w, ConvEpilogue{}, gq.g, 0, c.st));
      TRUN(qsoftmax_bwd(gq.g, Q, heads, dh, c.st));
      // dk~[(h,d)] = sum_e dctx[h][d][e] v[(h,e)]
      TRUN(blockdiag(dctx, c.B, heads, dh, false, 1.f, dn, c.st));
      TRUN(pack_conv_weights(dn, c.B * hd, hd, 1, hd, hd, c.B, nullptr, wb + ws, c.st));
      w.w = wb + ws;
      TRUN(conv_dispatch(ConvSrc{&Q, 2 * hd / 8, hd / 8}, ConvSrc{}, w, ConvEpilogue{}, gq.g, hd / 8, c.st));
      TRUN(ksoftmax_bwd(gq.g, Q, hd, ssum, c.st));
      // dv[(h,e)] = sum_d dctx[h][d][e] k~[(h,d)]
      TRUN(blockdiag(dctx, c.B, heads, dh, true, 1.f, dn, c.st));
      TRUN(pack_conv_weights(dn, c.B * hd, hd, 1, hd, hd, c.B, nullptr, wb + 2 * ws, c.st));
      w.w = wb + 2 * ws;
      TRUN(conv_dispatch(ConvSrc{&Q, hd / 8, hd / 8}, ConvSrc{}, w, ConvEpilogue{}, gq.g, 2 * hd / 8, c.st));
      TRUN(linattn_mem_bwd(c.pdev(p + ".mem_kv"), nm, kstat, dctx, ssum, c.B, heads, dh, c.gptr(p + ".mem_kv"), c.st));
      gq.init = true;
      return 0;
    });
    Act uo = c.like(x);
    const bool fz = fuse_ok(p + ".to_out.0", C);
    ConvEpilogue fo = fused_epilogue(p + ".to_out.1.g", "", C, false, &x);
    FTB_TRY(conv(p + ".to_out.0", o, nullptr, hd, 0, nullptr, false, nullptr, uo, false, false, true, fz ? &fo : nullptr, out));
    FTB_TRY(normact(uo, true, p + ".to_out.1.g", "", false, &x, *out, p + ".to_out.0.bias", fz));
    return 0;
  }

  // EmbedATb.forward (unet_attn_3d_cond_v3.py:131-139): trilinear(opened ATb) -> conv 5^3 -> SiLU -> conv 5^3
  int embed_atb(const std::string& p, const Act& opened, int cin_real, int C, int D, int H, int W, Act* out) {
    Act src = opened;
    if (opened.D != D || opened.H != H || opened.W != W) FTB_TRY(resample(opened, D, H, W, &src));
    Act u1 = c.act(C, D, H, W), e1 = c.like(u1);
    FTB_TRY(conv(p + ".conv1", src, nullptr, cin_real, 0, nullptr, false, nullptr, u1, false));
    FTB_TRY(normact(u1, false, "", "", true, nullptr, e1));
    *out = c.like(u1);
    FTB_TRY(conv(p + ".conv2", e1, nullptr, C, 0, nullptr, false, nullptr, *out, false));
    return 0;
  }

  // MixATb.forward (unet_attn_3d_cond_v3.py:175-190): FiLM(cat(x, emb)) -> conv 3^3 -> RMSNorm -> SiLU -> conv 3^3, + x.
  // The FiLM of the concat is applied to its two halves separately (rows [0,C) | [C,2C) of scale and shift) and the
  // conv reads them as its two sources, so the concatenated tensor is never formed.
  int mix_atb(const std::string& p, const Act& x, const Act& emb, Act* out) {
    const int C = x.C;
    const std::string lin = p + ".time_mlp.1";
    const int foff = U->film_off.at(lin);
    // (pushed first: in the backward it runs after both FiLM halves have written their rows of dfilm)
    T->tape.push_back([=](TrainCtx& c) -> int {
      TRUN(linear_bwd(c.dfilm + foff, c.U->film_rows, c.temb_silu, c.pdev(lin + ".weight"), c.B, 4 * C, c.U->time_dim,
                      c.gptr(lin + ".weight"), c.gptr(lin + ".bias"), c.dts, true, c.st));
      return 0;
    });
    Act xf = c.like(x), ef = c.like(emb);
    const FilmSlice fx{foff, foff + 2 * C}, fe{foff + C, foff + 3 * C};
    FTB_TRY(normact(x, false, "", "", false, nullptr, xf, "", false, &fx));
    FTB_TRY(normact(emb, false, "", "", false, nullptr, ef, "", false, &fe));
    Act u1 = c.like(x), h1 = c.like(x);
    const bool fz = fuse_ok(p + ".conv1", C);
    ConvEpilogue f1 = fused_epilogue(p + ".norm.g", "", C, true, nullptr);
    FTB_TRY(conv(p + ".conv1", xf, &ef, C, C, nullptr, false, nullptr, u1, false, false, true, fz ? &f1 : nullptr, &h1));
    FTB_TRY(normact(u1, true, p + ".norm.g", "", true, nullptr, h1, p + ".conv1.bias", fz));
    *out = c.like(x);
    FTB_TRY(conv(p + ".conv2", h1, nullptr, C, 0, &x, false, nullptr, *out, false));
    return 0;
  }

  // a contiguous range of parameters [first, last] is complete once the backward passes this point
  void marker(const std::string& first, const std::string& last) {
    const int i0 = U->pindex.at(first), i1 = U->pindex.at(last);
    const int64_t lo = T->poff[i0], hi = T->poff[i1] + U->params[i1].numel;
    T->tape.push_back([=](TrainCtx& c) -> int {
      if (c.cb && !c.dry) c.cb(c.cb_user, lo, hi - lo);
      return 0;
    });
  }

  int run(const float* x, const float* t, float* y, int X, int Y, int Z, const float* atb = nullptr) {
    const ftb_unet_cfg& cf = U->cfg;
    const bool cond = cf.conditional != 0;
    const int o = cond ? 2 : 0;   // module index of the first ResnetBlock inside a stage (after EmbedATb, MixATb)
    const int n = cf.n_stages;
    auto sub = [&](const std::string& p, int k) { return p + "." + std::to_string(k); };
    auto last_param = [&](const std::string& prefix) {   // last parameter whose name starts with prefix
      std::string r;
      for (const Param& p : U->params)
        if (p.name.compare(0, prefix.size(), prefix) == 0) r = p.name;
      return r;
    };
    auto first_param = [&](const std::string& prefix) {
      for (const Param& p : U->params)
        if (p.name.compare(0, prefix.size(), prefix) == 0) return p.name;
      return std::string();
    };
    T->tape.clear();
    const int td = U->time_dim, tr = cf.time_resolution;
    // ---- time path
    float* temb = c.f32((size_t)c.B * td);
    c.temb_silu = c.f32((size_t)c.B * td);
    float* tsave = c.f32((size_t)c.B * (tr + 2 * td));
    c.film_raw = c.f32((size_t)c.B * U->film_rows);
    c.film_fold = c.f32((size_t)c.B * U->film_rows);
    float* tcopy = c.f32((size_t)c.B);
    if (!c.dry) FTB_CUDA(cudaMemcpyAsync(tcopy, t, c.B * sizeof(float), cudaMemcpyDeviceToDevice, c.st));
    T->t_dev = tcopy;
    marker(U->params[0].name, "time_mlp.3.bias");
    {
      TimeMlpParams tp{c.pdev("time_mlp.0.freqs"), c.pdev("time_mlp.0.phases"), c.pdev("time_mlp.1.weight"),
                       c.pdev("time_mlp.1.bias"), c.pdev("time_mlp.3.weight"), c.pdev("time_mlp.3.bias"), tr, td};
      TRUN(time_embed(tp, tcopy, c.B, temb, c.temb_silu, c.st, tsave));
      FilmTable ft{U->d_film_w, U->d_film_b, nullptr, U->d_film_off, (int)U->film_blocks.size(), U->film_rows, td};
      TRUN(film_mlps(ft, c.temb_silu, c.B, c.film_raw, c.st));
      FilmTable ftf = ft;
      ftf.gs = U->d_film_gs;
      TRUN(film_mlps(ftf, c.temb_silu, c.B, c.film_fold, c.st));
      T->tape.push_back([=](TrainCtx& c) -> int {
        // d silu(temb) -> time_mlp.3 -> GELU -> time_mlp.1 -> Fourier features
        const int B = c.B;
        float* dy = c.f32((size_t)B * tr);
        float* dh1 = c.f32((size_t)B * td);
        const float* ysave = tsave;
        TRUN(act_bwd(c.dts, temb, (size_t)B * td, 0, c.st));   // dtemb = dts * silu'(temb)
        // temb = W2 h1 + b2: the kernels take x as [B][cols] contiguous, so stage h1 / y rows
        float* h1c = c.f32((size_t)B * td);
        float* prec = c.f32((size_t)B * td);
        float* yc = c.f32((size_t)B * tr);
        if (!c.dry) {
          const size_t pitch = (size_t)(tr + 2 * td) * sizeof(float);
          FTB_CUDA(cudaMemcpy2DAsync(yc, tr * sizeof(float), ysave, pitch, tr * sizeof(float), B, cudaMemcpyDeviceToDevice, c.st));
          FTB_CUDA(cudaMemcpy2DAsync(prec, td * sizeof(float), ysave + tr, pitch, td * sizeof(float), B, cudaMemcpyDeviceToDevice, c.st));
          FTB_CUDA(cudaMemcpy2DAsync(h1c, td * sizeof(float), ysave + tr + td, pitch, td * sizeof(float), B, cudaMemcpyDeviceToDevice, c.st));
        }
        TRUN(linear_bwd(c.dts, td, h1c, c.pdev("time_mlp.3.weight"), B, td, td, c.gptr("time_mlp.3.weight"),
                        c.gptr("time_mlp.3.bias"), dh1, false, c.st));
        TRUN(act_bwd(dh1, prec, (size_t)B * td, 1, c.st));
        TRUN(linear_bwd(dh1, td, yc, c.pdev("time_mlp.1.weight"), B, td, tr, c.gptr("time_mlp.1.weight"),
                        c.gptr("time_mlp.1.bias"), dy, false, c.st));
        TRUN(fourier_bwd(dy, tcopy, c.pdev("time_mlp.0.freqs"), c.pdev("time_mlp.0.phases"), B, tr,
                         c.gptr("time_mlp.0.freqs"), c.gptr("time_mlp.0.phases"), c.st));
        return 0;
      });
    }
    // ---- stem(s): NCDHW fp32 -> blocked bf16 (W-unfolded when the conv was planned that way) -> 7^3 conv
    auto stem_conv = [&](const std::string& name, const float* src, int cout, Act* out) -> int {
      const ConvLayer& stem = U->convs.at(name);
      Act xin = c.act(stem.unfold_w ? stem.k * stem.cin : stem.cin, X, Y, Z);
      if (stem.unfold_w) TRUN(pack_unfold_w(src, c.B, stem.cin, X, Y, Z, stem.k, xin, c.st));
      else TRUN(pack_ncdhw_to_blocked(src, c.B, stem.cin, X, Y, Z, xin, c.st));
      *out = c.act(cout, X, Y, Z);
      return conv(name, xin, nullptr, stem.cin, 0, nullptr, false, nullptr, *out, true);
    };
    Act opened;   // init_conv_ATb(ATb) (:778): its gradient collects from the EmbedATb of every stage
    if (cond) FTB_TRY(stem_conv("init_conv_ATb", atb, cf.data_channels, &opened));
    Act r;
    FTB_TRY(stem_conv(cond ? "init_conv_x" : "init_conv", x, cf.dim, &r));
    Act cur = r;
    std::vector<Act> skips;
    for (int i = 0; i < n; ++i) {
      const std::string p = "downs." + std::to_string(i);
      const int din = U->in_out[i].first, dout = U->in_out[i].second;
      marker(first_param(p + "."), last_param(p + "."));
      Act a1, a2, a3, a4;
      if (cond) {
        Act emb, mixed;
        FTB_TRY(embed_atb(sub(p, 0), opened, cf.data_channels, din, cur.D, cur.H, cur.W, &emb));
        FTB_TRY(mix_atb(sub(p, 1), cur, emb, &mixed));
        cur = mixed;
      }
      FTB_TRY(resnet(sub(p, o), cur, nullptr, din, 0, din, &a1));
      skips.push_back(a1);
      FTB_TRY(resnet(sub(p, o + 1), a1, nullptr, din, 0, din, &a2));
      FTB_TRY(attention(sub(p, o + 2), a2, cf.full_attn[i] != 0, &a3));
      skips.push_back(a3);
      if (i >= n - 1) {
        a4 = c.act(dout, a3.D, a3.H, a3.W);
        FTB_TRY(conv(sub(p, o + 3), a3, nullptr, din, 0, nullptr, false, nullptr, a4, false));
      } else {
        Act ds;
        FTB_TRY(resample(a3, a3.D / 2, a3.H / 2, a3.W / 2, &ds));
        a4 = c.act(dout, ds.D, ds.H, ds.W);
        FTB_TRY(conv(sub(p, o + 3) + ".conv", ds, nullptr, din, 0, nullptr, false, nullptr, a4, false));
      }
      cur = a4;
    }
    {
      const int mid = U->dims.back();
      marker(first_param("mid_block1."), last_param("mid_block2."));
      Act m1, m2, m3;
      FTB_TRY(resnet("mid_block1", cur, nullptr, mid, 0, mid, &m1));
      FTB_TRY(attention("mid_attn", m1, true, &m2));
      FTB_TRY(resnet("mid_block2", m2, nullptr, mid, 0, mid, &m3));
      cur = m3;
    }
    for (int i = 0; i < n; ++i) {
      const std::string p = "ups." + std::to_string(i);
      const int din = U->in_out[n - 1 - i].first, dout = U->in_out[n - 1 - i].second;
      marker(first_param(p + "."), last_param(p + "."));
      Act a1, a2, a3, a4;
      if (cond) {
        Act emb, mixed;
        FTB_TRY(embed_atb(sub(p, 0), opened, cf.data_channels, dout, cur.D, cur.H, cur.W, &emb));
        FTB_TRY(mix_atb(sub(p, 1), cur, emb, &mixed));
        cur = mixed;
      }
      Act s = skips.back(); skips.pop_back();
      FTB_TRY(resnet(sub(p, o), cur, &s, dout, din, dout, &a1));
      s = skips.back(); skips.pop_back();
      FTB_TRY(resnet(sub(p, o + 1), a1, &s, dout, din, dout, &a2));
      FTB_TRY(attention(sub(p, o + 2), a2, cf.full_attn[n - 1 - i] != 0, &a3));
      if (i == n - 1) {
        a4 = c.act(din, a3.D, a3.H, a3.W);
        FTB_TRY(conv(sub(p, o + 3), a3, nullptr, dout, 0, nullptr, false, nullptr, a4, false));
      } else {
        Act us;
        FTB_TRY(resample(a3, a3.D * 2, a3.H * 2, a3.W * 2, &us));
        a4 = c.act(din, us.D, us.H, us.W);
        FTB_TRY(conv(sub(p, o + 3) + ".conv", us, nullptr, dout, 0, nullptr, false, nullptr, a4, false));
      }
      cur = a4;
    }
    marker(first_param("final_res_block."), "final_conv.bias");
    Act fin;
    FTB_TRY(resnet("final_res_block", cur, &r, cf.dim, cf.dim, cf.dim, &fin));
    Act dummy = fin;
    FTB_TRY(conv("final_conv", fin, nullptr, cf.dim, 0, nullptr, false, y, dummy, false, true));
    return 0;
  }
};

// (re)pack the transposed weights of every conv that has a data gradient: one table-driven launch
inline int pack_dgrad(ftb_unet* U, float* /*wt_tmp*/, cudaStream_t st) {
  int n = 0;
  for (auto& kv : U->dgrad) n += (int)kv.second.size();
  const float* base0 = U->params.empty() ? nullptr : U->params[0].dev;
  if (n != U->n_djobs || U->djobs_base != base0) {
    std::vector<PackJob> jobs;
    for (auto& kv : U->dgrad) {
      const ConvLayer& cl = U->convs.at(kv.first);
      for (DgradPack& d : kv.second) {
        PackJob jb{};
        jb.w = U->params[U->pindex.at(cl.wname)].dev;
        jb.in_scale = cl.in_scale.empty() ? nullptr : cl.scale_tmp;   // pre-norm gain folded into the forward pack
        jb.dst = d.packed;
        jb.cout = d.cin_sub; jb.cin_real = cl.cout; jb.ksize = cl.k; jb.cin_pad = round_up(cl.cout, 16);
        jb.n = d.n_tile; jb.ntiles = 1; jb.unfold_w = 0;
        jb.transposed = 1; jb.ci0 = d.ci0; jb.src_cin = cl.cin;
        jobs.push_back(jb);
      }
    }
    if (n > U->cap_djobs) {
      FTB_TRY(dev_alloc(U, &U->d_djobs, (size_t)n + 64));
      U->cap_djobs = n + 64;
    }
    FTB_CUDA(cudaMemcpyAsync(U->d_djobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
    FTB_CUDA(cudaStreamSynchronize(st));
    U->n_djobs = n;
    U->djobs_base = base0;
  }
  FTB_TRY(pack_conv_weights_batched(U->d_djobs, U->n_djobs, st));
  U->dgrad_dirty = false;
  return 0;
}

}  // namespace ftb_engine_detail
