// fp32 mode of Unet3D.forward (unet_attn_3d.py:673-719): the accuracy path behind BASELINE's "fp32 velocity field
// within 1e-4 relative L2".  Activations stay NCDHW fp32 (the reference's own layout); every Conv3d runs on the
// tcgen05 implicit-GEMM kernel as three bf16 products with fp32 accumulation,
//     W x ~= W_hi x_hi + W_hi x_lo + W_lo x_hi      (hi = bf16(.), lo = bf16(. - hi)),
// written / accumulated straight into the fp32 output by the conv epilogue; norms, FiLM, SiLU, residuals, both
// attentions and the trilinear resamples are plain fp32 kernels (f32_ops.cu).  Included by engine.cu.
#pragma once

namespace ftb_engine_detail {

// Weight packs of every conv (no folded gains, no W-unfolding), rebuilt when the parameters change.  With CP = pad16(Cin):
//   3*CP <= 512: ONE launch, sources (x_hi | x_lo) and x_hi again, weights [W_hi | W_hi | W_lo] along K — the three
//                products share the TMEM accumulator and the fp32 output is written once;
//   2*CP <= 512: (x_hi | x_lo) x [W_hi | W_hi], then x_hi x W_lo accumulated in place by the epilogue;
//   else       : three launches of CP channels each (also when the K-extended input would need more channel-chunk
//                passes than the conv planner accepts, e.g. the 5^3 192 -> 192 EmbedATb conv).
inline int f32_mode(int cp, int k, int n) {
  const int cg = cp / 8, lim = conv_max_chunks();
  if (3 * cp <= 512 && conv_chunk_count(2 * cg, k, n) + conv_chunk_count(cg, k, n) <= lim) return 3;
  if (2 * cp <= 512 && conv_chunk_count(2 * cg, k, n) <= lim) return 2;
  return 1;
}

int finalize_f32(ftb_unet* U, cudaStream_t st) {
  const float* base0 = U->params.empty() ? nullptr : U->params[0].dev;
  if (!U->d_f32jobs || U->f32jobs_base != base0) {
    std::vector<PackJob> jobs;
    for (auto& kv : U->convs) {
      const ConvLayer& cl = kv.second;
      const int cp = round_up(cl.cin, 16), mode = f32_mode(cp, cl.k, cl.n_tile);
      const size_t elems = (size_t)cl.ntiles * cl.k * cl.k * cl.k * cp * cl.n_tile;
      std::pair<bf16*, bf16*>& pk = U->f32packs[kv.first];
      if (!pk.first) {
        FTB_TRY(dev_alloc(U, &pk.first, elems * (mode == 1 ? 1 : mode)));
        if (mode != 3) FTB_TRY(dev_alloc(U, &pk.second, elems));
      }
      for (int which = 0; which < (mode == 3 ? 1 : 2); ++which) {
        PackJob jb{};
        jb.w = U->params[U->pindex[cl.wname]].dev;
        jb.in_scale = nullptr;
        jb.dst = which == 0 ? pk.first : pk.second;
        jb.cout = cl.cout; jb.cin_real = cl.cin; jb.ksize = cl.k; jb.n = cl.n_tile;
        jb.ntiles = cl.ntiles; jb.unfold_w = 0;
        if (which == 1) { jb.cin_pad = cp; jb.part = 2; }               // W_lo
        else if (mode == 3) { jb.cin_pad = 3 * cp; jb.part = 3; }       // [W_hi | W_hi | W_lo]
        else if (mode == 2) { jb.cin_pad = 2 * cp; jb.part = 4; }       // [W_hi | W_hi]
        else { jb.cin_pad = cp; jb.part = 1; }                          // W_hi
        jobs.push_back(jb);
      }
    }
    if (!U->d_f32jobs) FTB_TRY(dev_alloc(U, &U->d_f32jobs, jobs.size()));
    FTB_CUDA(cudaMemcpyAsync(U->d_f32jobs, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
    FTB_CUDA(cudaStreamSynchronize(st));   // `jobs` is pageable host memory
    U->n_f32jobs = (int)jobs.size();
    U->f32jobs_base = base0;
    U->f32_stale = true;
  }
  if (!U->f32_stale) return 0;
  FTB_TRY(pack_conv_weights_batched(U->d_f32jobs, U->n_f32jobs, st));
  U->f32_stale = false;
  return 0;
}

struct T32 {
  float* p = nullptr;
  int C = 0, D = 0, H = 0, W = 0;
  size_t vox() const { return (size_t)D * H * W; }
};

struct FwdF32 {
  ftb_unet* U;
  cudaStream_t st;
  char* base;
  bool dry;
  int B;
  size_t off = 0, peak = 0;
  float* film = nullptr;
  // FTB_F32_KEEP=1 (debugging): temporaries are never released, so every tap survives the forward
  bool keep = getenv("FTB_F32_KEEP") != nullptr;

  void release(size_t mark) {
    if (keep) return;
    off = mark;
    // taps that lived in the released region are about to be overwritten
    for (auto it = U->taps32.begin(); it != U->taps32.end();)
      it = ((uintptr_t)it->second[0] >= (uintptr_t)(base + mark)) ? U->taps32.erase(it) : std::next(it);
  }
  void tap(const std::string& name, const T32& t) {
    if (!dry && U->keep_taps) U->taps32[name] = std::vector<long long>{(long long)(uintptr_t)t.p, B, t.C, t.D, t.H, t.W};
  }

  void* raw(size_t bytes) {
    off = round_up_sz(off, 256);
    void* p = base ? base + off : nullptr;
    off += bytes;
    if (off > peak) peak = off;
    return p;
  }
  float* f32(size_t n) { return reinterpret_cast<float*>(raw(n * sizeof(float))); }
  T32 t32(int C, int D, int H, int W) {
    T32 t;
    t.C = C; t.D = D; t.H = H; t.W = W;
    t.p = f32((size_t)B * C * t.vox());
    return t;
  }
  T32 like(const T32& a, int C) { return t32(C, a.D, a.H, a.W); }
  const float* pdev(const std::string& n) { return U->params[U->pindex[n]].dev; }

  // out = conv3d(a || b) + bias, three bf16 products accumulated in fp32
  int conv(const std::string& name, const T32& a, const T32* b, const T32& out) {
    const ConvLayer& cl = U->convs.at(name);
    const int cin = a.C + (b ? b->C : 0);
    FTB_CHECK(cin == cl.cin && out.C == cl.cout, "fp32 conv '" + name + "': channel mismatch");
    const int cp = round_up(cin, 16);
    const size_t mark = off;
    Act sp;
    sp.B = B; sp.C = 2 * cp; sp.D = a.D; sp.H = a.H; sp.W = a.W;
    sp.p = reinterpret_cast<bf16*>(raw(sp.bytes()));
    if (!dry) {
      FTB_TRY(f32_pack_split(a.p, a.C, b ? b->p : nullptr, b ? b->C : 0, B, a.vox(), sp, st));
      const std::pair<bf16*, bf16*>& pk = U->f32packs.at(name);
      ConvWeights w;
      w.ksize = cl.k; w.cin = cp; w.n = cl.n_tile; w.ntiles = cl.ntiles;
      w.cin_real = cl.cin; w.cout_real = cl.cout;
      ConvEpilogue e;
      e.out_f32 = out.p;
      e.out_f32_c = cl.cout;
      Act dummy = sp;   // spatial dims only: the epilogue writes NCDHW fp32
      const ConvSrc hi{&sp, 0, cp / 8}, lo{&sp, cp / 8, cp / 8}, hilo{&sp, 0, 2 * cp / 8};
      const int mode = f32_mode(cp, cl.k, cl.n_tile);
      e.bias = cl.bname.empty() ? nullptr : cl.bias;
      w.w = pk.first;
      if (mode == 3) {
        w.cin = 3 * cp;
        FTB_TRY(conv_igemm(hilo, hi, w, e, dummy, 0, st));           // W_hi x_hi + W_hi x_lo + W_lo x_hi + bias
      } else if (mode == 2) {
        w.cin = 2 * cp;
        FTB_TRY(conv_igemm(hilo, ConvSrc{}, w, e, dummy, 0, st));    // W_hi (x_hi + x_lo) + bias
      } else {
        FTB_TRY(conv_igemm(hi, ConvSrc{}, w, e, dummy, 0, st));      // W_hi x_hi + bias
        e.bias = nullptr;
        e.out_f32_accum = true;
        FTB_TRY(conv_igemm(lo, ConvSrc{}, w, e, dummy, 0, st));      // += W_hi x_lo
      }
      if (mode != 3) {
        e.bias = nullptr;
        e.out_f32_accum = true;
        w.cin = cp;
        w.w = pk.second;
        FTB_TRY(conv_igemm(hi, ConvSrc{}, w, e, dummy, 0, st));      // += W_lo x_hi
      }
      U->launches += 1 + (mode == 3 ? 1 : (mode == 2 ? 2 : 3));
    }
    off = mark;   // the split operand is dead once the three launches are enqueued (stream order)
    return 0;
  }

  int normact(const T32& u, bool norm, const float* gain, const float* s1, const float* sh, int fstride, bool silu,
              const float* resid, const T32& out) {
    if (dry) return 0;
    U->launches += 1;
    return f32_normact(u.p, B, u.C, u.vox(), norm, gain, s1, sh, fstride, silu, resid, out.p, st);
  }

  // ResnetBlock.forward (:265-278) on (x0 || x1); `out` is allocated by the caller
  int resnet(const std::string& p, const T32& x0, const T32* x1, const T32& out) {
    const int cout = out.C, cin = x0.C + (x1 ? x1->C : 0);
    const float* film_p = film ? film + U->film_off.at(resnet_mlp(U, p)) : nullptr;
    const size_t mark = off;
    T32 h = like(x0, cout);
    FTB_TRY(conv(p + ".block1.proj", x0, x1, h));
    // Block1 (:232-244): RMSNorm gain and FiLM (scale + 1) arrive pre-multiplied from film_mlps
    FTB_TRY(normact(h, true, nullptr, film_p, film_p + cout, U->film_rows, true, nullptr, h));
    tap(p + ".block1", h);
    const float* resp = x0.p;
    if (cin != cout) {
      T32 res = like(x0, cout);
      FTB_TRY(conv(p + ".res_conv", x0, x1, res));
      resp = res.p;
    } else {
      FTB_CHECK(x1 == nullptr, "fp32 resnet: identity residual of a concat input");
    }
    T32 h2 = like(x0, cout);
    FTB_TRY(conv(p + ".block2.proj", h, nullptr, h2));
    FTB_TRY(normact(h2, true, U->gains.at(p + ".block2.norm.g").gs, nullptr, nullptr, 0, true, resp, out));
    tap(p, out);
    release(mark);
    return 0;
  }

  // x + attn(x)  (:695, :702, :712); LinearAttention :308-341, Attention :357-373 / :436-465
  int attention(const std::string& p, const T32& x, bool full, const T32& out) {
    const ftb_unet_cfg& c = U->cfg;
    const int heads = c.attn_heads, dh = c.attn_dim_head, hd = heads * dh;
    const ConvLayer& cq = U->convs.at(p + ".to_qkv");
    const size_t mark = off;
    T32 xn = like(x, x.C);
    FTB_TRY(normact(x, true, cq.scale_tmp, nullptr, nullptr, 0, false, nullptr, xn));   // RMSNorm: g * sqrt(C)
    T32 qkv = like(x, 3 * hd);
    FTB_TRY(conv(p + ".to_qkv", xn, nullptr, qkv));
    T32 ao = like(x, hd);
    T32 o = like(x, x.C);
    if (full) {
      FTB_CHECK(x.vox() <= (size_t)1 << 20, "fp32 attention: too many tokens");
      if (!dry) {
        FTB_TRY(f32_full_attention(qkv.p, B, heads, dh, (int)x.vox(), pdev(p + ".mem_kv"), c.num_mem_kv, ao.p, st));
        U->launches += 1;
      }
      FTB_TRY(conv(p + ".to_out", ao, nullptr, o));
      FTB_TRY(normact(o, false, nullptr, nullptr, nullptr, 0, false, x.p, out));
    } else {
      float* scratch = f32(f32_linattn_scratch(B, heads, dh, x.vox()));
      if (!dry) {
        FTB_TRY(f32_linear_attention(qkv.p, B, heads, dh, x.vox(), pdev(p + ".mem_kv"), c.num_mem_kv, scratch, ao.p, st));
        U->launches += 5;
      }
      FTB_TRY(conv(p + ".to_out.0", ao, nullptr, o));
      FTB_TRY(normact(o, true, U->gains.at(p + ".to_out.1.g").gs, nullptr, nullptr, 0, false, x.p, out));
    }
    tap(p + ".attn_out", ao);
    tap(p, out);
    release(mark);
    return 0;
  }

  // EmbedATb.forward (unet_attn_3d_cond_v3.py:131-139); `out` allocated by the caller
  int embed_atb(const std::string& p, const T32& opened, const T32& out) {
    const size_t mark = off;
    T32 src = opened;
    if (opened.D != out.D || opened.H != out.H || opened.W != out.W) {
      src = like(out, opened.C);
      FTB_TRY(resample(opened, src));
    }
    T32 e1 = like(out, out.C);
    FTB_TRY(conv(p + ".conv1", src, nullptr, e1));
    FTB_TRY(normact(e1, false, nullptr, nullptr, nullptr, 0, true, nullptr, e1));   // SiLU
    FTB_TRY(conv(p + ".conv2", e1, nullptr, out));
    release(mark);
    return 0;
  }

  // MixATb.forward (unet_attn_3d_cond_v3.py:175-190): FiLM on cat(x, emb) (applied to the two halves), conv, norm, SiLU,
  // conv, + x
  int mix_atb(const std::string& p, const T32& x, const T32& emb, const T32& out) {
    const int C = x.C;
    const float* fp = film + U->film_off.at(p + ".time_mlp.1");   // (scale + 1)[2C] | shift[2C]
    const size_t mark = off;
    T32 xf = like(x, C), ef = like(x, C), h = like(x, C), h2 = like(x, C);
    FTB_TRY(normact(x, false, nullptr, fp, fp + 2 * C, U->film_rows, false, nullptr, xf));
    FTB_TRY(normact(emb, false, nullptr, fp + C, fp + 3 * C, U->film_rows, false, nullptr, ef));
    FTB_TRY(conv(p + ".conv1", xf, &ef, h));
    FTB_TRY(normact(h, true, U->gains.at(p + ".norm.g").gs, nullptr, nullptr, 0, true, nullptr, h));
    FTB_TRY(conv(p + ".conv2", h, nullptr, h2));
    FTB_TRY(normact(h2, false, nullptr, nullptr, nullptr, 0, false, x.p, out));
    tap(p, out);
    release(mark);
    return 0;
  }

  int resample(const T32& in, const T32& out) {
    if (dry) return 0;
    U->launches += 1;
    return f32_trilinear(in.p, B, in.C, in.D, in.H, in.W, out.D, out.H, out.W, out.p, st);
  }

  int run(const float* x, const float* t, float* y, int X, int Y, int Z, const float* atb = nullptr) {
    const ftb_unet_cfg& c = U->cfg;
    const int n = c.n_stages;
    const bool cond = c.conditional != 0;
    const int o = cond ? 2 : 0;
    auto sub = [&](const std::string& p, int k) { return p + "." + std::to_string(k); };
    U->taps32.clear();
    U->launches = 0;
    float* temb = f32((size_t)B * U->time_dim);
    float* temb_silu = f32((size_t)B * U->time_dim);
    film = f32((size_t)B * U->film_rows);
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
    T32 xin;
    xin.p = const_cast<float*>(x); xin.C = c.data_channels; xin.D = X; xin.H = Y; xin.W = Z;
    T32 opened;
    if (cond) {   // init_conv_ATb (:778); one conditioning volume per sample in this mode
      T32 ain;
      ain.p = const_cast<float*>(atb); ain.C = c.data_channels; ain.D = X; ain.H = Y; ain.W = Z;
      opened = t32(c.data_channels, X, Y, Z);
      FTB_TRY(conv("init_conv_ATb", ain, nullptr, opened));
      tap("init_conv_ATb", opened);
    }
    const std::string init_name = cond ? "init_conv_x" : "init_conv";
    T32 r = t32(c.dim, X, Y, Z);
    FTB_TRY(conv(init_name, xin, nullptr, r));
    tap(init_name, r);
    T32 cur = r;
    std::vector<T32> skips;
    for (int i = 0; i < n; ++i) {
      const std::string p = "downs." + std::to_string(i);
      const int din = U->in_out[i].first, dout = U->in_out[i].second;
      T32 a1 = like(cur, din);
      T32 a3 = like(cur, din);          // second skip; temporaries and the stage output live above the two skips
      const size_t mark = off;
      if (cond) {
        T32 emb = like(cur, din), mixed = like(cur, din);
        FTB_TRY(embed_atb(sub(p, 0), opened, emb));
        FTB_TRY(mix_atb(sub(p, 1), cur, emb, mixed));
        cur = mixed;
      }
      FTB_TRY(resnet(sub(p, o), cur, nullptr, a1));
      skips.push_back(a1);
      T32 a2 = like(cur, din);
      FTB_TRY(resnet(sub(p, o + 1), a1, nullptr, a2));
      FTB_TRY(attention(sub(p, o + 2), a2, c.full_attn[i] != 0, a3));
      skips.push_back(a3);
      release(mark);
      T32 a4;
      if (i >= n - 1) {
        a4 = like(a3, dout);
        FTB_TRY(conv(sub(p, o + 3), a3, nullptr, a4));
      } else {
        a4 = t32(dout, a3.D / 2, a3.H / 2, a3.W / 2);
        const size_t m2 = off;
        T32 ds = t32(din, a3.D / 2, a3.H / 2, a3.W / 2);
        FTB_TRY(resample(a3, ds));
        FTB_TRY(conv(sub(p, o + 3) + ".conv", ds, nullptr, a4));
        release(m2);
      }
      tap(sub(p, o + 3), a4);
      cur = a4;
    }
    {
      const int mid = U->dims.back();
      T32 m3 = like(cur, mid);
      const size_t mark = off;
      T32 m1 = like(cur, mid), m2 = like(cur, mid);
      FTB_TRY(resnet("mid_block1", cur, nullptr, m1));
      FTB_TRY(attention("mid_attn", m1, true, m2));
      FTB_TRY(resnet("mid_block2", m2, nullptr, m3));
      release(mark);
      cur = m3;
    }
    for (int i = 0; i < n; ++i) {
      const std::string p = "ups." + std::to_string(i);
      const int din = U->in_out[n - 1 - i].first, dout = U->in_out[n - 1 - i].second;
      // stage output first (it outlives the stage's temporaries)
      T32 a4 = (i == n - 1) ? like(cur, din) : t32(din, cur.D * 2, cur.H * 2, cur.W * 2);
      const size_t mark = off;
      if (cond) {
        T32 emb = like(cur, dout), mixed = like(cur, dout);
        FTB_TRY(embed_atb(sub(p, 0), opened, emb));
        FTB_TRY(mix_atb(sub(p, 1), cur, emb, mixed));
        cur = mixed;
      }
      T32 a1 = like(cur, dout), a2 = like(cur, dout), a3 = like(cur, dout);
      T32 s = skips.back(); skips.pop_back();
      FTB_TRY(resnet(sub(p, o), cur, &s, a1));
      s = skips.back(); skips.pop_back();
      FTB_TRY(resnet(sub(p, o + 1), a1, &s, a2));
      FTB_TRY(attention(sub(p, o + 2), a2, c.full_attn[n - 1 - i] != 0, a3));
      if (i == n - 1) {
        FTB_TRY(conv(sub(p, o + 3), a3, nullptr, a4));
      } else {
        T32 us = t32(dout, a4.D, a4.H, a4.W);
        FTB_TRY(resample(a3, us));
        FTB_TRY(conv(sub(p, o + 3) + ".conv", us, nullptr, a4));
      }
      tap(sub(p, o + 3), a4);
      release(mark);
      cur = a4;
    }
    T32 fin = like(cur, c.dim);
    FTB_TRY(resnet("final_res_block", cur, &r, fin));
    T32 yo;
    yo.p = y; yo.C = c.data_channels; yo.D = X; yo.H = Y; yo.W = Z;
    FTB_TRY(conv("final_conv", fin, nullptr, yo));
    return 0;
  }
};

}  // namespace ftb_engine_detail
