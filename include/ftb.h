/* flowtrain-b200 C ABI — the drop-in boundary of the B200-native hot path.
 *
 * The reference (chipnbits/flowtrain_stochastic_interpolation) is pure Python and has no FFI;
 * its seam is the Python calling convention `model(XT, T)` (src/flowtrain/solvers/solvers.py:70,
 * :136, :198, :235-238) / `self.net(XT, T)` (project/geodata-3d-unconditional/
 * model_train_inference.py:440).  Each entry point below replaces the reference code cited at
 * it; the Python host (flowtrain_stochastic_interpolation_b200/) binds them with ctypes.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; tensors are contiguous,
 * fp32, in the reference's NCDHW order ([B,C,X,Y,Z], Z fastest); `stream` is a cudaStream_t
 * passed as void*; calls enqueue work and return without synchronising.  Return value 0 = ok,
 * negative = error (ftb_last_error() gives a thread-local message).  Nothing throws across the
 * ABI.  Ownership: the caller owns every tensor and the workspace; a handle owns only its
 * packed weights.  A handle is bound to the device current at creation; calls on one handle must
 * be serialised by the caller (one host thread per rank), distinct handles are independent.
 * There is no CPU fallback: without a CUDA device every compute entry returns an error.
 */
#ifndef FTB_H_
#define FTB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTB_MAX_STAGES 8

/* Constructor arguments of Unet3D (src/flowtrain/models/unet_attn_3d.py:509-525). */
typedef struct ftb_unet_cfg {
  int dim;                          /* base channels (multiple of 16) */
  int n_stages;                     /* len(dim_mults) */
  int dim_mults[FTB_MAX_STAGES];
  int data_channels;                /* embedding dim E (18 unconditional) */
  int time_resolution;              /* Fourier features */
  int attn_heads;
  int attn_dim_head;                /* 16 or 32 */
  int full_attn[FTB_MAX_STAGES];    /* per stage: 1 = softmax Attention, 0 = LinearAttention */
  int num_mem_kv;                   /* 4 in the reference (:290, :345) */
  int conditional;                  /* 1: Unet3DCond v3 (unet_attn_3d_cond_v3.py:598-828), 0: Unet3D */
} ftb_unet_cfg;

typedef struct ftb_unet ftb_unet;

const char* ftb_last_error(void);
int ftb_version(void);
int ftb_device_sm_count(void);
/* kernels launched by this library in this process so far (bench.py "gpu_launches") */
int64_t ftb_launch_count(void);
/* optional CUDA-event timing around every conv_igemm launch on its own stream (roofline leg of
 * bench.py).  collect() synchronises the recorded events and sums, per kind (0: 3x3x3/5^3/7^3
 * convs, 1: 1x1x1 convs), algorithmic FLOPs, algorithmic bytes, milliseconds and launches. */
int ftb_profile_enable(int on);
int ftb_profile_collect(double* flops, double* bytes, double* ms, int* launches, int nkinds);

/* ---- Unet3D velocity field v_theta(x, t): replaces Unet3D.forward (unet_attn_3d.py:673-719) */
/* create() only builds the plan (works without a GPU, so names/shapes can be queried);
 * device storage is allocated on the first set_param()/forward() on the current device. */
int ftb_unet3d_create(const ftb_unet_cfg* cfg, ftb_unet** out);
int ftb_unet3d_destroy(ftb_unet* h);
int ftb_unet3d_num_params(const ftb_unet* h);
const char* ftb_unet3d_param_name(const ftb_unet* h, int i);   /* state_dict() key order */
int64_t ftb_unet3d_param_numel(const ftb_unet* h, int i);
/* writes the shape into dims[0..ndim) and returns ndim (or -1) */
int ftb_unet3d_param_shape(const ftb_unet* h, int i, int* dims, int max_dims);
/* copy one fp32 parameter (device pointer) into the handle; repacked lazily before forward */
int ftb_unet3d_set_param(ftb_unet* h, const char* name, const float* data, int64_t numel, void* stream);
size_t ftb_unet3d_workspace_bytes(ftb_unet* h, int B, int X, int Y, int Z);
/* x [B,C,X,Y,Z] fp32, t [B] fp32 -> out [B,C,X,Y,Z] fp32.  bf16 tensor-core path. */
int ftb_unet3d_forward(ftb_unet* h, const float* x, const float* t, float* out, int B, int X, int Y,
                       int Z, void* workspace, size_t workspace_bytes, void* stream);
/* fp32 accuracy mode of the same call (BASELINE: fp32 velocity field within 1e-4 relative L2 of the reference,
 * src/flowtrain/models/unet_attn_3d.py:673-719): NCDHW fp32 activations, every Conv3d as three bf16 tcgen05
 * products (W_hi x_hi + W_hi x_lo + W_lo x_hi) accumulated in fp32.  About 3-4x the cost of the bf16 path. */
size_t ftb_unet3d_f32_workspace_bytes(ftb_unet* h, int B, int X, int Y, int Z);
int ftb_unet3d_forward_f32(ftb_unet* h, const float* x, const float* t, float* out, int B, int X, int Y, int Z,
                           void* workspace, size_t workspace_bytes, void* stream);
/* conditional model in the fp32 mode: atb [B,C,X,Y,Z] (one conditioning volume per sample; nothing is cached) */
int ftb_unet3d_cond_forward_f32(ftb_unet* h, const float* x, const float* atb, const float* t, float* out, int B,
                                int X, int Y, int Z, void* workspace, size_t workspace_bytes, void* stream);
/* after ftb_unet3d_forward_f32: dims[5] = (B, C, X, Y, Z) of a named intermediate; out (may be NULL) receives it */
int ftb_unet3d_get_tap_f32(ftb_unet* h, const char* name, float* out, int* dims, void* stream);
/* ---- conditional velocity field: replaces Unet3DCond.forward(x, ATb, time)
 *      (src/flowtrain/models/unet_attn_3d_cond_v3.py:769-828; handle created with cfg.conditional = 1).
 * atb is [atb_B, C, X, Y, Z] fp32 with atb_B = B, or 1 when one conditioning volume is shared by the
 * whole batch (the ensemble case, project/geodata-3d-conditional/model_inference_experiments.py:232).
 * init_conv_ATb and the ten EmbedATb outputs depend on ATb only: they live at the start of the
 * workspace, and reuse_atb != 0 skips recomputing them when the caller passes the SAME workspace,
 * shapes and ATb as in the previous call (the reference recomputes them on every evaluation). */
size_t ftb_unet3d_cond_workspace_bytes(ftb_unet* h, int B, int atb_B, int X, int Y, int Z);
int ftb_unet3d_cond_forward(ftb_unet* h, const float* x, const float* atb, int atb_B, const float* t, float* out,
                            int B, int X, int Y, int Z, void* workspace, size_t workspace_bytes, int reuse_atb,
                            void* stream);
/* after a forward: copy a named intermediate (same names as oracle/unet3d.py taps) as NCDHW fp32 */
int ftb_unet3d_tap_channels(ftb_unet* h, const char* name, int* C, int* X, int* Y, int* Z);
int ftb_unet3d_get_tap(ftb_unet* h, const char* name, float* out, void* stream);
/* number of kernels the last forward launched (for bench.py "gpu_launches") */
int ftb_unet3d_last_launches(const ftb_unet* h);

/* ---- interpolant: replaces StochasticInterpolator.get_XT / get_BT / flow_objective
 *      (src/flowtrain/interpolation/interpolation.py:78-117, :156-216).
 *      kind: 0 linear, 1 trig, 2 enc-dec, 3 SBDM, 4 mirror (:379-546).  z and bt may be NULL. */
int ftb_interp_xt_bt(int kind, int one_sided, float gamma_a, const float* x0, const float* x1,
                     const float* z, const float* t, float* xt, float* bt, int B, int64_t n_per_sample,
                     void* stream);

/* ---- fixed-grid integrator updates around model(x,t) (solvers.py:66-77, odeSol_RK4 :235-240) */
/* h is a double on the host; the kernels use float(h), float(h/2), float(h/6) like torch
 * does for `tensor * python_float`.  out = x + h*k ; frozen (optional, bytes, length `inner`) zeroes k where frozen[i % inner] (:73) */
int ftb_ode_axpy(float* out, const float* x, const float* k, double h, int64_t n,
                 const unsigned char* frozen, int64_t inner, void* stream);
int ftb_ode_heun_combine(float* out, const float* x, const float* k1, const float* k2, double h,
                         int64_t n, void* stream);
int ftb_ode_rk4_combine(float* out, const float* x, const float* k1, const float* k2, const float* k3,
                        const float* k4, double h, int64_t n, void* stream);
/* ---- adaptive-step Runge-Kutta (torchdiffeq's dopri5 / adaptive_heun behind ODEFlowSolver & co, solvers.py:77, :148,
 *      :220-222; torchdiffeq >=0.2.5,<0.3 is an un-vendored dependency: restated from its published algorithm, parity
 *      unpinned).  The step controller is host code; these are the passes over the fp32 state.  k: nk host-array of
 *      device pointers (stage derivatives; a pointer may be NULL where its weight is 0), coef: nk host doubles
 *      (tableau weight * dt).  n must be a multiple of 4.
 *   lincomb:      out = y0 + sum_j coef[j] k[j]
 *   error_ratio:  acc[0] += sum ((sum_j coef[j] k[j]) / (atol + rtol max(|y0|, |y1|)))^2      (device double)
 *   scaled_sumsq: acc[0] += sum ((a1 - a2) / (atol + rtol |y|))^2   (a2 may be NULL; initial-step heuristic)
 *   dense_eval:   out = quartic interpolant through (y0, ymid, y1) with end slopes (f0, f1) at x in [0, 1] */
int ftb_ode_lincomb(float* out, const float* y0, const float* const* k, const double* coef, int nk, int64_t n,
                    void* stream);
int ftb_ode_error_ratio(const float* y0, const float* y1, const float* const* k, const double* coef, int nk, float rtol,
                        float atol, int64_t n, double* acc, void* stream);
int ftb_ode_scaled_sumsq(const float* a1, const float* a2, const float* y, float rtol, float atol, int64_t n,
                         double* acc, void* stream);
int ftb_ode_dense_eval(float* out, const float* y0, const float* y1, const float* ymid, const float* f0,
                       const float* f1, double dt, double x, int64_t n, void* stream);
/* ---- device-resident step controller of the adaptive solvers (replaces torchdiffeq's host-side
 *      RKAdaptiveStepsizeODESolver loop behind solvers.py:77, :148, :220-222): time, step size, accept / reject, output
 *      cursor and counters live in `ctl`, 16 device doubles:
 *        [0] t  [1] dt  [2] error-ratio accumulator  [3..5] first-step norm accumulators  [6] tp0 [7] tp1 [8] dtp
 *        (interval and size of the last accepted step)  [9] accepted  [10] rejected  [11] out_lo [12] out_hi
 *        (outputs [out_lo, out_hi) are emitted by this step)  [13] flags: 1 accepted this step, 2 finished,
 *        4 non-finite error estimate, 8 max_num_steps exceeded  [14] h0  [15] attempted steps
 *      A whole solve is enqueued without a device->host read; the host polls an asynchronous copy of `ctl` from an
 *      earlier step to stop enqueueing (steps after `finished` are no-ops on the state).
 *   ctl_init:        zero the controller, t = t0
 *   ctl_first_step:  torchdiffeq _select_initial_step; phase 0: dt = h0 from ctl[3], ctl[4] (sums of squares of
 *                    y0 / scale and f0 / scale, written by ftb_ode_scaled_sumsq); phase 1: dt = min(100 h0, h1), ctl[5]
 *   ctl_stage_time:  tbuf[b] = float(t + alpha dt), the model's time input of a stage
 *   lincomb_dev / error_ratio_dev: as lincomb / error_ratio with coef = tableau weights, multiplied by ctl's dt on
 *                    the device; error_ratio_dev accumulates into ctl[2]
 *   ctl_step:        ratio = sqrt(ctl[2] / n); accept iff <= 1; t, dt (factor min(10, max(0.9 ratio^(-1/order),
 *                    accepted ? 1 : 0.2))), counters, which grid points (device doubles, n_out of them) fall in the step
 *   advance:         if accepted: write those outputs (quartic dense output; traj [n_out, n] or, when NULL, only the
 *                    final point into `last`) and y0 <- y1, f0 <- f1 in place; otherwise nothing */
int ftb_ode_ctl_init(double* ctl, double t0, void* stream);
int ftb_ode_ctl_first_step(double* ctl, int phase, int64_t n, int order, void* stream);
int ftb_ode_ctl_stage_time(float* tbuf, const double* ctl, double alpha, int B, void* stream);
int ftb_ode_lincomb_dev(float* out, const float* y0, const float* const* k, const double* coef, int nk, int64_t n,
                        const double* ctl, void* stream);
int ftb_ode_error_ratio_dev(const float* y0, const float* y1, const float* const* k, const double* coef, int nk,
                            float rtol, float atol, int64_t n, double* ctl, void* stream);
int ftb_ode_ctl_step(double* ctl, const double* grid, int n_out, int64_t n, int order, int64_t max_steps, void* stream);
int ftb_ode_advance(float* y0, float* f0, const float* y1, const float* f1, const float* const* k, const double* c_mid,
                    int nk, const double* ctl, const double* grid, int n_out, float* traj, float* last, int64_t n,
                    void* stream);
/* eq. 6.7 drift from a denoiser eta (solvers.py:130-143); SDE term (:205-216) when use_sde */
int ftb_denoise_drift(float* out, const float* x, const float* eta, const float* noise, float alpha,
                      float beta, float alpha_dot, float beta_dot, float eps, int use_sde, int64_t n,
                      void* stream);
/* the same with coef = device floats (alpha, beta, alpha_dot, beta_dot, eps) of the evaluation's (device-resident) time */
int ftb_denoise_drift_dev(float* out, const float* x, const float* eta, const float* noise, const float* coef,
                          int use_sde, int64_t n, void* stream);

/* ---- categorical embedding (model_train_inference.py:361-370, :373-404) */
/* x [B,E,n] fp32, en [ncat,E] = F.normalize(embedding.weight) -> out [B,n] int64, bit-exact order */
int ftb_decode(const float* x, const float* en, int64_t* out, int B, int E, int ncat, int64_t n,
               void* stream);
/* decode(x, return_logits=True) (:398-399): logits [B,ncat,n] fp32, same op order as ftb_decode */
int ftb_decode_logits(const float* x, const float* en, float* logits, int B, int E, int ncat, int64_t n, void* stream);
/* cats [B,n] int64 (+shift, :366) , w [ncat,E] -> out [B,E,n] */
int ftb_embed(const int64_t* cats, const float* w, float* out, int B, int E, int ncat, int64_t n,
              int shift, void* stream);

/* ---- conditioning front-end: replaces make_combined_mask + embed + ATb = X1 * mask
 *      (project/geodata-3d-conditional/boreholes.py:45-126; model_train_sh_inference_cond.py:413-420).
 * cats [B,X,Y,Z] int64 (-1 = air); bores [B,max_bores,2] int32 borehole (x, y) columns, n_bores [B] how many of
 * them are valid (the random draw stays with the caller); w [ncat,E].  Outputs (each may be NULL):
 * mask [B,X,Y,Z] uint8 = (surface != 0: top slice | air | voxel below air) | borehole column;  w may be NULL when
 * only the mask is wanted;  x1 [B,E,X,Y,Z] = w[cat + shift];
 * atb [B,E,X,Y,Z] = x1 * mask. */
int ftb_cond_frontend(const int64_t* cats, const int32_t* bores, const int32_t* n_bores, int max_bores, const float* w,
                      int B, int E, int ncat, int shift, int X, int Y, int Z, int surface, uint8_t* mask, float* x1,
                      float* atb, void* stream);
/* ---- loss of the conditional training step (model_train_sh_inference_cond.py:432-452):
 *      loss = mse(VT, VThat) / (mse(VT, 0) + 1e-6) + lambda * mean(T) * mse(b, b_hat) / (mse(X1, 0) + 1e-6),
 *      b = X1_clean[mask], b_hat = (XT + (1 - T) VThat)[mask].  accumulate adds the six partial sums
 *      (sum (v-vh)^2, sum v^2, sum_mask (b-b_hat)^2, #masked elements, sum X1_noisy^2, sum T) to acc6 (device doubles);
 *      grad writes dout = scale * d loss / d VThat from them (no host sync in between). */
int ftb_cond_loss_accumulate(const float* vt, const float* vhat, const float* xt, const float* x1_clean,
                             const float* x1_noisy, const uint8_t* mask, const float* t, int B, int E, int64_t n,
                             double* acc6, void* stream);
int ftb_cond_loss_grad(const float* vt, const float* vhat, const float* xt, const float* x1_clean, const uint8_t* mask,
                       const float* t, int B, int E, int64_t n, const double* acc6, float lambda_reconstruct, float scale,
                       float* dout, void* stream);
/* ---- ensemble statistics: replaces decode -> one_hot -> mean / entropy / argmax
 *      (project/geodata-3d-conditional/model_inference_experiments.py:442-459; inference_demo.ipynb cell 21).
 * decode_vote: x [S,E,n] fp32 samples, en [ncat,E] normalised embedding -> counts [ncat,n] int32 += votes (bit-exact
 * decode order as ftb_decode); decoded [S,n] int64 optional.  Counts of several launches / ranks add up (one
 * all-reduce(sum) of counts across GPUs).  vote_finalize: S = total samples; probs [ncat,n] = counts / S,
 * entropy [n] = -sum p log(p + 1e-8), most_probable [n] = first argmax + shift (-1: air),
 * entropy_masked [n] = entropy, -1 where most_probable == -1.  Each output may be NULL. */
int ftb_decode_vote(const float* x, const float* en, int S, int E, int ncat, int64_t n, int64_t* decoded,
                    int32_t* counts, void* stream);
int ftb_vote_finalize(const int32_t* counts, int S, int ncat, int64_t n, int shift, float* probs, float* entropy,
                      int64_t* most_probable, float* entropy_masked, void* stream);

/* ---- training-step pieces (model_train_inference.py:443; callbacks.py:263-266) */
int ftb_ema_update(float* shadow, const float* param, int64_t n, double decay, void* stream);
/* acc2[0] += sum (v-vhat)^2, acc2[1] += sum v^2 (device doubles; loss = acc2[0]/acc2[1]) */
int ftb_mse_ratio_accumulate(const float* v, const float* vhat, int64_t n, double* acc2, void* stream);

/* ---- training step: replaces autograd through Unet3D.forward inside Geo3DStochInterp.training_step
 *      (project/geodata-3d-unconditional/model_train_inference.py:417-457) and the optimiser of
 *      configure_optimizers (:465-473).  Unconditional model, dropout 0.
 * Parameters may live in ONE caller-owned flat fp32 buffer in state_dict order (bind_params; offsets
 * from param_offset, param_offset(h, num_params) = total); after changing it (optimiser step) call
 * mark_dirty so the packed copies are rebuilt before the next forward.
 * forward_train keeps every intermediate in the workspace and records the backward; backward (once per
 * forward_train, same workspace) ADDS the parameter gradients into grads_flat (state_dict order, fp32).
 * bucket_cb(user, offset, count), optional, is called on the host right after the launches that complete
 * the gradient range [offset, offset+count) were enqueued, last layers first, so a data-parallel caller
 * can start an all-reduce of that range while the rest of the backward runs (DDP-style overlap). */
int64_t ftb_unet3d_param_offset(ftb_unet* h, int i);
int ftb_unet3d_bind_params(ftb_unet* h, float* flat_params, void* stream);
int ftb_unet3d_mark_dirty(ftb_unet* h);
/* nn.Dropout(p) at the end of every ResnetBlock.block1 (unet_attn_3d.py:244, :261) for the NEXT forward_train:
 * counter-based keep mask keyed by `seed` (pass a fresh seed per step), regenerated by the backward, scale 1/(1-p).
 * The random stream is this library's own (torch's Philox dropout stream cannot be matched); p = 0 turns it off. */
int ftb_unet3d_set_dropout(ftb_unet* h, float p, uint64_t seed);
size_t ftb_unet3d_train_workspace_bytes(ftb_unet* h, int B, int X, int Y, int Z);
int ftb_unet3d_forward_train(ftb_unet* h, const float* x, const float* t, float* out, int B, int X, int Y, int Z,
                             void* workspace, size_t workspace_bytes, void* stream);
/* conditional model (Unet3DCond.forward(x, ATb, time) under autograd, model_train_sh_inference_cond.py:431):
 * atb [B,C,X,Y,Z] fp32 (one conditioning volume per sample, ATb = X1 * mask, :414-420).  The backward is the same
 * ftb_unet3d_backward; ATb and x are data, so no input gradient is produced. */
int ftb_unet3d_cond_forward_train(ftb_unet* h, const float* x, const float* atb, const float* t, float* out, int B,
                                  int X, int Y, int Z, void* workspace, size_t workspace_bytes, void* stream);
int ftb_unet3d_backward(ftb_unet* h, const float* dout, float* grads_flat, void* workspace, size_t workspace_bytes,
                        void (*bucket_cb)(void* user, int64_t offset, int64_t count), void* cb_user, void* stream);
/* gradient of the flow loss (model_train_inference.py:443) w.r.t. vhat, from the sums ftb_mse_ratio_accumulate
 * produced: dout = scale * 2 (vhat - v) / acc2[1] */
int ftb_mse_ratio_grad(const float* v, const float* vhat, int64_t n, const double* acc2, float scale, float* dout,
                       void* stream);
/* acc[0] += sum g^2 (device double) */
int ftb_grad_sumsq(const float* g, int64_t n, double* acc, void* stream);
/* torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.Adam (decoupled = 0, weight_decay as L2) or AdamW
 * (decoupled = 1) on flat buffers.  The gradient is first multiplied by grad_scale (1/world_size after a
 * sum all-reduce); sumsq (device double, may be NULL) is the squared norm of the UNSCALED gradient and
 * max_norm <= 0 disables clipping.  step counts from 1. */
int ftb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int decoupled, int step, const double* sumsq, float grad_scale,
                  float max_norm, void* stream);

/* ---- single-op test hooks (allocate scratch internally; not for hot paths) */
/* conv3d "same", stride 1, on NCDHW fp32 tensors through the blocked bf16 kernels.
 * impl: 0 = tcgen05 implicit GEMM, 1 = direct CUDA-core kernel.  x2/resid/bias/g/scale/shift may
 * be NULL.  flags: bit0 SiLU, bit1 pre-norm row scale.  Epilogue: bias -> RMSNorm(g) -> FiLM -> SiLU
 * -> +resid. */
int ftb_test_conv3d(const float* x, int c1, const float* x2, int c2, const float* w, const float* bias,
                    int cout, int ksize, const float* g, const float* scale, const float* shift,
                    const float* resid, int flags, float* out, int B, int X, int Y, int Z, int impl,
                    void* stream);
/* conv3d weight / input gradients through the tcgen05 kernels (operands rounded to bf16 inside) */
int ftb_test_conv_wgrad(const float* x, int c1, const float* x2, int c2, const float* dy, int cout, int ksize,
                        float* dw, int B, int X, int Y, int Z, int unfold, void* stream);
int ftb_test_conv_dgrad(const float* dy, const float* w, int cout, int cin, int ksize, const float* acc, float* dx,
                        int B, int X, int Y, int Z, void* stream);
int ftb_test_trilinear(const float* x, int B, int C, int X, int Y, int Z, int Xo, int Yo, int Zo,
                       float* out, void* stream);
int ftb_test_trilinear_bwd(const float* dout, int B, int C, int X, int Y, int Z, int Xo, int Yo, int Zo,
                           const float* acc, float* din, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FTB_H_ */
