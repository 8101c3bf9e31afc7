// Host-side launch API of the flowtrain-b200 kernels (internal; the public C ABI is include/ftb.h).
#pragma once
#include "ftb_common.cuh"

namespace ftb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- conv (implicit GEMM)
// Packed weights: [ntile][kh][kw][kstep][j][N/8][2][8][8] bf16, j = K-1-kd — per (kh,kw,kstep) a
// (K*N) x 16 B-operand tile in the no-swizzle K-major core-matrix layout (LBO = 128 B,
// SBO = 256 B) whose K row blocks are the depth taps in descending kd (depth-tap stacking).
struct ConvWeights {
  const bf16* w = nullptr;
  int ksize = 1;        // 1, 3, 5, 7 (stride 1, "same" zero padding): extent along D and H
  int ksize_w = 0;      // extent along W; 0 = cubic.  1 with ksize > 1: the W taps were unfolded into
                        // the channels (c' = kw*Cin + ci, pack_unfold_w), so K waste of a small-Cin stem
                        // (18 -> 32 per tap) becomes 7*18 = 126 -> 128 per (kd, kh)
  int kw() const { return ksize_w > 0 ? ksize_w : ksize; }
  int cin = 0;          // padded K extent (multiple of 16) = channels of src0 (+ src1)
  int n = 0;            // output channels of one N tile (multiple of 16, <= 256)
  int ntiles = 1;       // number of N tiles packed back to back
  long long batch_stride = 0;  // elements between per-sample weight sets (0 = shared)
  int cin_real = 0, cout_real = 0;  // unpadded channel counts (algorithmic-FLOP accounting only)
  size_t tile_elems() const { return (size_t)ksize * ksize * kw() * cin * n; }
};

// y = ((acc * rs + bias) * rinv) * mul[b][c] + add[b][c], then SiLU, then + resid.
//   rs   = 1/max(||src0 voxel||, 1e-12) when prenorm (pre-attention RMSNorm folded into a 1x1 conv)
//   rinv = 1/max(||acc*rs + bias|| over the N channels, 1e-12) when norm (RMSNorm, :127-128)
//   mul  = g*sqrt(C) (RMSNorm gain), for Block1 pre-multiplied by the FiLM (scale+1) (:241)
//   add  = FiLM shift
struct ConvEpilogue {
  const float* bias = nullptr;   // [ntiles*n]
  bool norm = false;
  const float* mul = nullptr;    // [B or 1][mul_stride]: per-channel factor (first n entries used)
  int mul_stride = 0;            // 0: shared by all samples
  const float* add = nullptr;    // [B or 1][add_stride]
  int add_stride = 0;
  bool silu = false;
  const Act* resid = nullptr;    // residual added last (blocked bf16, same spatial dims)
  int resid_cgoff = 0;
  bool prenorm = false;          // row scale 1/max(||src0 voxel||_2, 1e-12) (fused pre-RMSNorm)
  const float* prenorm_ss = nullptr;  // optional [B][voxels]: ||src0 voxel||^2 from the producer's sumsq_out
  float* sumsq_out = nullptr;    // optional [B][voxels]: sum over channels of the stored output squared
  int q_softmax_heads = 0;       // >0: softmax over each dim_head group of the first N tile
  int q_dim_head = 0;
  float q_scale = 1.f;
  float* out_f32 = nullptr;      // NCDHW fp32 output (final conv) instead of blocked bf16
  int out_f32_c = 0;
  bool out_f32_accum = false;    // fp32 mode: add to out_f32 instead of overwriting it
  // training: also store the pre-norm tensor (acc*rs + bias, shaped like `out`) that the backward of the
  // RMSNorm / FiLM / SiLU needs, and apply nn.Dropout after the activation (mask: ftb_common.cuh drop_mask8)
  const Act* pre_out = nullptr;
  float drop_p = 0.f;
  unsigned long long drop_key = 0;
};

struct ConvSrc {
  const Act* t = nullptr;
  int cgoff = 0;   // first channel group used
  int cg = 0;      // channel groups used (multiple of 2)
};

// out channels [out_cgoff*8, out_cgoff*8 + ntiles*n) of `out` are written.
int conv_igemm(const ConvSrc& s0, const ConvSrc& s1, const ConvWeights& w, const ConvEpilogue& e,
               Act& out, int out_cgoff, cudaStream_t st);
// Same contract, plain CUDA-core direct convolution (validation / FTB_CONV_IMPL=naive).
int conv_naive(const ConvSrc& s0, const ConvSrc& s1, const ConvWeights& w, const ConvEpilogue& e,
               Act& out, int out_cgoff, cudaStream_t st);
int conv_dispatch(const ConvSrc& s0, const ConvSrc& s1, const ConvWeights& w, const ConvEpilogue& e,
                  Act& out, int out_cgoff, cudaStream_t st);

int conv_debug_read(long long* host, int n);
// planner facts the fp32 mode needs to choose its launch variant
int conv_chunk_count(int cg, int K, int N);
int conv_max_chunks();

// 3-D TMA view of a blocked activation as (8 channels, voxels, B*CG); box = (8, box_vox, box_cg):
// lands in shared memory as [cg][voxel][8], the no-swizzle MN-major UMMA operand layout.
int make_voxel_tmap(CUtensorMap* tm, const Act& a, int box_vox, int box_cg, int* vdiv = nullptr);

// fp32 [Cout][Cin][k^3] -> packed bf16 tiles.  cin_map: K index -> source Cin index is
// identity for k < cin_real, zero beyond.  in_scale (optional, [cin_real]) folds a per-input-
// channel factor (pre-norm gain * sqrt(C)) into the weights.
// unfold_w: pack for a W-unfolded input: K index c' = kw*cin_real + ci, taps (kd, kh) only.
int pack_conv_weights(const float* w, int cout, int cin_real, int ksize, int cin_pad, int ntile_n,
                      int ntiles, const float* in_scale, bf16* dst, cudaStream_t st, bool unfold_w = false);

// batched form: a device table of jobs, one launch.  Plain job: as pack_conv_weights.  Transposed job (data
// gradient): dst = pack of wT[co'][ci'][flipped taps] = w[ci'][ci0 + co'] * in_scale[ci0 + co'], with cout = number
// of source channels, cin_real = forward Cout, src_cin = forward Cin.
struct PackJob {
  const float* w;
  const float* in_scale;
  bf16* dst;
  int cout, cin_real, ksize, cin_pad, n, ntiles, unfold_w;
  int transposed, ci0, src_cin;
  int part;   // 0 / 1: bf16(w);  2: lo = bf16(w - bf16(w));  fp32 mode (3 x bf16 split) with the K extent repeated:
              // 3: cin_pad = 3 x pad16(cin), weights [hi | hi | lo];  4: cin_pad = 2 x pad16(cin), weights [hi | hi]
};
int pack_conv_weights_batched(const PackJob* d_jobs, int njobs, cudaStream_t st);

// ---------------------------------------------------------------- layout / resample
int pack_ncdhw_to_blocked(const float* x, int B, int C, int D, int H, int W, Act& out, cudaStream_t st);
// W-unfolded pack: out channel kw*C + c of voxel (d,h,w) = x[c][d][h][w + kw - K/2] (zero outside)
int pack_unfold_w(const float* x, int B, int C, int D, int H, int W, int K, Act& out, cudaStream_t st);
int unpack_blocked_to_ncdhw(const Act& in, int cgoff, int C, float* out, cudaStream_t st);
int trilinear_resample(const Act& in, Act& out, cudaStream_t st);  // align_corners=True
// out[:, :C] = x*mul[:C]+add[:C]; out[:, C:] = atb*mul[C:]+add[C:], film row = [mul(2C) | add(2C)] per sample
int film_concat(const Act& x, const Act& atb, const float* film, int film_stride, int c_real, Act& out,
                cudaStream_t st);

// ---------------------------------------------------------------- time path
struct TimeMlpParams {
  const float *freqs, *phases, *w1, *b1, *w2, *b2;  // Fourier + Linear(tr->td) + Linear(td->td)
  int time_res, time_dim;
};
// save (optional, training): [B][time_res + 2*time_dim] = Fourier features | pre-GELU | GELU
int time_embed(const TimeMlpParams& p, const float* t, int B, float* temb, float* temb_silu,
               cudaStream_t st, float* save = nullptr);
// all per-block FiLM MLPs in one launch: out[b][off_j + o] = W_j[o,:] . silu(temb[b]) + b_j[o]
// (first half of each block = scale, emitted as (scale+1)*gs_j; second half = shift)
struct FilmTable {
  const float* const* w;  // device array of weight pointers [nblk] (each [rows_j][time_dim])
  const float* const* b;  // device array of bias pointers
  const float* const* gs; // device array of Block1 gain vectors g*sqrt(C) [rows_j/2] (entries may be null):
                          // the scale half of block j is emitted as (scale + 1) * gs, ready for the conv epilogue
  const int* row_off;     // device prefix offsets [nblk+1]
  int nblk, total_rows, time_dim;
};
int film_mlps(const FilmTable& ft, const float* temb_silu, int B, float* out, cudaStream_t st);

// ---------------------------------------------------------------- attention
// qkv: blocked tensor with 3*heads*dh channels (q | k | v).  Linear attention, phase A:
// per-(b,head,split) online-softmax partials of k over voxels and ctx = softmax(k) v^T.
int linattn_kmax(const Act& qkv, int heads, int dh, int nsplit, float* kmax, cudaStream_t st);
int linattn_context_partial(const Act& qkv, int heads, int dh, int nsplit, const float* kmax,
                            float* part, cudaStream_t st);
// phase A2: merge partials + memory kv, fold W_out -> per-sample packed 1x1 weights M_b[C][heads*dh]
// fused k/v projection + context (never writes k, v): x blocked bf16 (channel groups [cgoff, cgoff+cg)),
// ss = ||x voxel||^2, wk / wv = packed K-major tiles of the k / v rows of to_qkv, shift[hd] >= |k|
// shift[d] = 1.02 ||W[row0 + d, :] (x) in_scale||_2: row0 = hd for the k rows of to_qkv (softmax shift), 0 for the q rows
int linattn_kshift(const float* w_qkv, const float* in_scale, int hd, int cin, float* shift, cudaStream_t st, int row0);
int linattn_kv_context(const Act& x, int cgoff, int cg, const float* ss, const bf16* wk, const bf16* wv,
                       const float* shift, int heads, int dh, int nsplit, float* part, cudaStream_t st);
// fused q projection + q softmax + (context . q) + to_out + bias + RMSNorm + residual (never writes q):
// wq = packed q rows of to_qkv, mb = per-sample packed folded projection from linattn_combine
int linattn_q_out(const Act& x, const float* ss, const bf16* wq, const bf16* mb, long long mb_bstride,
                  const float* bias, const float* gs, int heads, int dh, Act& out, cudaStream_t st,
                  bool q_bounded = false);
// kmax_bstride: elements between samples of `kmax` (0: one shift vector for the whole batch)
int linattn_combine(const float* part, int nsplit, const float* kmax, int kmax_bstride, int B, int heads, int dh,
                    const float* mem_kv, int n_mem, const float* w_out, int C, float q_scale,
                    bf16* wpack_out, float* ctx_dbg, cudaStream_t st, float* kstat = nullptr);
// softmax attention over n tokens (+ n_mem memory kv), one CTA per (b, head, query tile)
int full_attention(const Act& qkv, int heads, int dh, const float* mem_kv, int n_mem, Act& out,
                   cudaStream_t st);

// ---------------------------------------------------------------- training step (wgrad.cu, train_ops.cu)
// dw (fp32 [cout][cin_tot][K][K][K]) += weight gradient of source x (channel groups [x_cgoff, x_cgoff+x_cg)),
// which feeds the weight's input channels [ci_base, ci_base+ci_real).  unfold_cin > 0: x is the W-unfolded
// stem input (channel kw*unfold_cin + ci).  dw_bstride != 0: one gradient slab per sample.
// partials: scratch of conv_wgrad_partial_bytes() (stream-ordered reuse between launches is fine).
int conv_wgrad(const Act& x, int x_cgoff, int x_cg, const Act& dy, int dy_cgoff, int cout_real, int ksize,
               int unfold_cin, float* dw, int cin_tot, int ci_base, int ci_real, long long dw_bstride,
               cudaStream_t st, float* partials = nullptr, size_t partial_bytes = 0);
// scratch for the per-CTA partial gradients of one conv_wgrad launch (one wave of CTAs x 128 rows x 512 columns fp32);
// without it (or when a launch needs more) the kernel falls back to its atomic epilogue
size_t conv_wgrad_partial_bytes();
// out = act(n * gain[c] * s1[b][c] + sh[b][c]) + resid, n = u / max(||u||_C, 1e-12) when norm
// drop_p > 0: dropout after the activation with a counter-based mask keyed by drop_key (regenerated in the backward)
int normact_fwd(const Act& u, bool norm, const float* gain, const float* s1, const float* sh, int fstride, bool silu,
                const Act* resid, Act& out, cudaStream_t st, float drop_p = 0.f, unsigned long long drop_key = 0);
// du (may alias dout); R[b][c] += sum_v dz*n, S[b*sstride + c] += sum_v dz, dbias[c] += sum du (each optional)
int normact_bwd(const Act& dout, const Act& u, bool norm, const float* gain, const float* s1, const float* sh,
                int fstride, bool silu, Act& du, float* R, float* S, int sstride, float* dbias, cudaStream_t st,
                float drop_p = 0.f, unsigned long long drop_key = 0);
int normact_finish(const float* R, int B, int C, const float* gain, const float* s1, int fstride, float sqrt_c,
                   float* ds1, float* dg, cudaStream_t st);
int bias_grad(const Act& dy, int cgoff, int C, float* db, cudaStream_t st);
int act_accum(Act& dst, const Act& src, bool add, cudaStream_t st);
int trilinear_resample_bwd(const Act& dout, Act& din, bool add, cudaStream_t st);
int transpose_flip(const float* w, int cout, int cin, int ksize, int ci0, int cin_sub, const float* row_scale,
                   float* wt, cudaStream_t st);
int fold_gain_bwd(const float* dwp, const float* w, const float* gs, int cout, int cin, float sqrt_c, float* dw,
                  float* dg, cudaStream_t st);
int blockdiag(const float* src, int B, int heads, int dh, bool transpose, float mul, float* out, cudaStream_t st);
int dctx_extract(const float* full, const float* ctx, int B, int heads, int dh, float* dctx, float* ssum,
                 cudaStream_t st);
int ksoftmax_apply(Act& qkv, int hd, const float* kstat, cudaStream_t st);
int ksoftmax_bwd(Act& dqkv, const Act& qkv, int hd, const float* ssum, cudaStream_t st);
int qsoftmax_bwd(Act& dqkv, const Act& qkv, int heads, int dh, cudaStream_t st);
int linattn_mem_bwd(const float* mem_kv, int n_mem, const float* kstat, const float* dctx, const float* ssum, int B,
                    int heads, int dh, float* dmem, cudaStream_t st);
int full_attention_bwd(const Act& qkv, const Act& ao, const Act& dao, int heads, int dh, const float* mem_kv,
                       int n_mem, float* scratch, Act& dqkv, float* dmem, cudaStream_t st);
int linear_bwd(const float* dy, int dy_stride, const float* x, const float* w, int B, int rows, int cols, float* dw,
               float* db, float* dx, bool dx_add, cudaStream_t st);
int act_bwd(float* g, const float* pre, size_t n, int mode, cudaStream_t st);   // 0: SiLU', 1: GELU(erf)'
int fourier_bwd(const float* dy, const float* t, const float* f, const float* phi, int B, int n, float* df, float* dphi,
                cudaStream_t st);
int mse_ratio_grad(const float* v, const float* vhat, long long n, const double* acc2, float gscale, float* dout,
                   cudaStream_t st);
int grad_sumsq(const float* g, long long n, double* acc, cudaStream_t st);
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              float wd, int decoupled, int step, const double* sumsq, float gscale, float max_norm, cudaStream_t st);

// ---------------------------------------------------------------- fp32 mode (f32_ops.cu): NCDHW fp32 tensors
// x0 (|| x1) [B][C][vox] fp32 -> blocked bf16 with 2*CP channels: [hi(CP) | lo(CP)], hi = bf16(x), lo = bf16(x - hi)
int f32_pack_split(const float* x0, int c0, const float* x1, int c1, int B, size_t vox, Act& out, cudaStream_t st);
// out = act(n * gain[c] * s1[b][c] + sh[b][c]) + resid on NCDHW fp32, n = u / max(||u||_C, 1e-12) when norm
int f32_normact(const float* u, int B, int C, size_t vox, bool norm, const float* gain, const float* s1, const float* sh,
                int fstride, bool silu, const float* resid, float* out, cudaStream_t st);
int f32_trilinear(const float* in, int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo, float* out, cudaStream_t st);
// LinearAttention on qkv [B][3*hd][n] fp32 (q | k | v): out [B][hd][n]; scratch: f32_linattn_scratch() floats
size_t f32_linattn_scratch(int B, int heads, int dh, size_t n);
int f32_linear_attention(float* qkv, int B, int heads, int dh, size_t n, const float* mem_kv, int n_mem, float* scratch,
                         float* out, cudaStream_t st);
// softmax attention, qkv [B][3*hd][n] fp32, mem_kv [2][heads][n_mem][dh]: out [B][hd][n]
int f32_full_attention(const float* qkv, int B, int heads, int dh, int n, const float* mem_kv, int n_mem, float* out,
                       cudaStream_t st);

// ---------------------------------------------------------------- sampler / task kernels (fp32, NCDHW flat)
int interp_xt_bt(int kind, int one_sided, float gamma_a, const float* x0, const float* x1,
                 const float* z, const float* t, float* xt, float* bt, int B, long long n,
                 cudaStream_t st);
int axpy_out(float* out, const float* x, const float* k, float h, long long n, const unsigned char* frozen,
             long long inner, cudaStream_t st);  // out = x + h*k   (Euler / RK stage input)
int heun_combine(float* out, const float* x, const float* k1, const float* k2, double h, long long n,
                 cudaStream_t st);
int rk4_combine(float* out, const float* x, const float* k1, const float* k2, const float* k3,
                const float* k4, double h, long long n, cudaStream_t st);
int denoise_drift(float* out, const float* x, const float* eta, const float* noise, float a, float b,
                  float ad, float bd, float eps, int use_sde, long long n, cudaStream_t st);
int decode_argmax(const float* x, const float* en, long long* out, int B, int E, int ncat,
                  long long n, cudaStream_t st);
int decode_logits(const float* x, const float* en, float* logits, int B, int E, int ncat, long long n, cudaStream_t st);
// ---- adaptive-step Runge-Kutta passes (ode_adaptive.cu); k: nk device pointers, coef: nk host doubles (already * dt)
int ode_lincomb(float* out, const float* y0, const float* const* k, const double* coef, int nk, long long n,
                cudaStream_t st);
int ode_error_ratio(const float* y0, const float* y1, const float* const* k, const double* coef, int nk, float rtol,
                    float atol, long long n, double* acc, cudaStream_t st);
int ode_scaled_sumsq(const float* a1, const float* a2, const float* y, float rtol, float atol, long long n, double* acc,
                     cudaStream_t st);
int ode_dense_eval(float* out, const float* y0, const float* y1, const float* ymid, const float* f0, const float* f1,
                   double dt, double x, long long n, cudaStream_t st);
// device-resident step controller (16 doubles, layout in include/ftb.h); coef: tableau weights NOT multiplied by dt
int ode_ctl_init(double* ctl, double t0, cudaStream_t st);
int ode_ctl_first_step(double* ctl, int phase, long long n, int order, cudaStream_t st);
int ode_ctl_stage_time(float* tbuf, const double* ctl, double alpha, int B, cudaStream_t st);
int ode_lincomb_dev(float* out, const float* y0, const float* const* k, const double* coef, int nk, long long n,
                    const double* ctl, cudaStream_t st);
int ode_error_ratio_dev(const float* y0, const float* y1, const float* const* k, const double* coef, int nk, float rtol,
                        float atol, long long n, double* ctl, cudaStream_t st);
int ode_ctl_step(double* ctl, const double* grid, int n_out, long long n, int order, long long max_steps, cudaStream_t st);
int ode_advance(float* y0, float* f0, const float* y1, const float* f1, const float* const* k, const double* c_mid, int nk,
                const double* ctl, const double* grid, int n_out, float* traj, float* last, long long n, cudaStream_t st);
int denoise_drift_dev(float* out, const float* x, const float* eta, const float* noise, const float* coef, int use_sde,
                      long long n, cudaStream_t st);
// ---- conditional project / ensemble kernels (cond_ops.cu, elementwise.cu)
// surface + borehole mask, X1 = W[cat + shift], ATb = X1 * mask in one pass; bores [B][max_b][2] int32 (x, y),
// nb [B] counts; mask (u8 [B][X*Y*Z]), x1, atb may each be null
int cond_frontend(const long long* cats, const int* bores, const int* nb, int max_b, const float* w, int B, int E,
                  int ncat, int shift, int X, int Y, int Z, int surface, unsigned char* mask, float* x1, float* atb,
                  cudaStream_t st);
int cond_loss_partial(const float* vt, const float* vh, const float* xt, const float* x1c, const float* x1n,
                      const unsigned char* mask, const float* T, int B, int E, long long n, double* acc6, cudaStream_t st);
int cond_loss_grad(const float* vt, const float* vh, const float* xt, const float* x1c, const unsigned char* mask,
                   const float* T, int B, int E, long long n, const double* acc6, float lambda, float scale, float* dout,
                   cudaStream_t st);
// decode S samples [S][E][n] and add them to counts [ncat][n] (int32); decoded [S][n] int64 optional
int decode_vote(const float* x, const float* en, int S, int E, int ncat, long long n, long long* decoded, int* counts,
                cudaStream_t st);
int vote_finalize(const int* counts, int S, int ncat, long long n, int shift, float* probs, float* entropy,
                  long long* most, float* entropy_masked, cudaStream_t st);
int embed_lookup(const long long* cats, const float* w, float* out, int B, int E, int ncat,
                 long long n, int shift, cudaStream_t st);
int ema_update(float* shadow, const float* param, long long n, double decay, cudaStream_t st);
int mse_ratio_partial(const float* v, const float* vhat, long long n, double* acc2, cudaStream_t st);

}  // namespace ftb
