// Adaptive-step Runge-Kutta support kernels: the reference hands ode_func to torchdiffeq's adaptive
// dopri5 (ODEFlowSolver / ODEOneSidedDenoisingSolver, src/flowtrain/solvers/solvers.py:77, :148) and adaptive_heun
// (SDEOneSidedDenoisingSolver, :220-222).  torchdiffeq (pinned >=0.2.5,<0.3 in pyproject.toml:19) is not vendored and
// not installed here, so this is a restatement of its published algorithm (rk_common.py: _runge_kutta_step,
// _compute_error_ratio, _interp_fit / _interp_evaluate), "parity unpinned" (DESIGN §2).  The step controller itself is
// host code (solvers.py); these are the HBM-bound passes over the fp32 state, one launch each:
//   lincomb     : out = y0 + sum_j c[j] * k[j]                (stage inputs y_i, the 5th-order solution y1, y_mid)
//   error_ratio : acc += sum ((sum_j e[j] k[j]) / (atol + rtol max(|y0|, |y1|)))^2   (double accumulator)
//   dense_eval  : out = quartic Hermite-style interpolant through (y0, y_mid, y1, f0, f1) at x = (t - t0)/(t1 - t0)
#include "ops.h"

namespace ftb {

namespace {

inline int grid_for(size_t work_items, int threads) {
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

struct KSet {
  const float* k[8];
  float c[8];
  int n;
};

__global__ void __launch_bounds__(256)
lincomb_kernel(float* __restrict__ out, const float* __restrict__ y0, KSet ks, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(y0) + i);
    // torchdiffeq accumulates y0 + sum_j k_j * (beta_j * dt) left to right in the state dtype
#pragma unroll 8
    for (int j = 0; j < ks.n; ++j) {
      if (ks.c[j] == 0.f) continue;
      const float4 k = __ldg(reinterpret_cast<const float4*>(ks.k[j]) + i);
      a.x = fmaf(k.x, ks.c[j], a.x); a.y = fmaf(k.y, ks.c[j], a.y);
      a.z = fmaf(k.z, ks.c[j], a.z); a.w = fmaf(k.w, ks.c[j], a.w);
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

__global__ void __launch_bounds__(256)
error_ratio_kernel(const float* __restrict__ y0, const float* __restrict__ y1, KSet ks, float rtol, float atol,
                   size_t n4, double* __restrict__ acc) {
  __shared__ double red[8];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(y0) + i), b = __ldg(reinterpret_cast<const float4*>(y1) + i);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int j = 0; j < ks.n; ++j) {
      if (ks.c[j] == 0.f) continue;
      const float4 k = __ldg(reinterpret_cast<const float4*>(ks.k[j]) + i);
      e.x = fmaf(k.x, ks.c[j], e.x); e.y = fmaf(k.y, ks.c[j], e.y);
      e.z = fmaf(k.z, ks.c[j], e.z); e.w = fmaf(k.w, ks.c[j], e.w);
    }
    const float rx = e.x / (atol + rtol * fmaxf(fabsf(a.x), fabsf(b.x)));
    const float ry = e.y / (atol + rtol * fmaxf(fabsf(a.y), fabsf(b.y)));
    const float rz = e.z / (atol + rtol * fmaxf(fabsf(a.z), fabsf(b.z)));
    const float rw = e.w / (atol + rtol * fmaxf(fabsf(a.w), fabsf(b.w)));
    s += (double)rx * rx + (double)ry * ry + (double)rz * rz + (double)rw * rw;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(acc, t);
  }
}

// sum (a[i] * sa / (atol + rtol |y[i]|))^2 with a = a1 - a2 (a2 optional): the three norms of _select_initial_step
__global__ void __launch_bounds__(256)
scaled_sumsq_kernel(const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ y, float rtol,
                    float atol, size_t n, double* __restrict__ acc) {
  __shared__ double red[8];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = (__ldg(a1 + i) - (a2 ? __ldg(a2 + i) : 0.f)) / (atol + rtol * fabsf(__ldg(y + i)));
    s += (double)v * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(acc, t);
  }
}

__global__ void __launch_bounds__(256)
dense_eval_kernel(float* __restrict__ out, const float* __restrict__ y0, const float* __restrict__ y1,
                  const float* __restrict__ ym, const float* __restrict__ f0, const float* __restrict__ f1, float dt,
                  float x, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a0 = __ldg(y0 + i), a1 = __ldg(y1 + i), am = __ldg(ym + i), g0 = __ldg(f0 + i), g1 = __ldg(f1 + i);
    // _interp_fit (rk_common.py): quartic through y0, y_mid, y1 with end slopes f0, f1
    const float a = 2.f * dt * (g1 - g0) - 8.f * (a1 + a0) + 16.f * am;
    const float b = dt * (5.f * g0 - 3.f * g1) + 18.f * a0 + 14.f * a1 - 32.f * am;
    const float c = dt * (g1 - 4.f * g0) - 11.f * a0 - 5.f * a1 + 16.f * am;
    const float d = dt * g0;
    out[i] = a0 + x * (d + x * (c + x * (b + x * a)));   // _interp_evaluate
  }
}

int make_kset(KSet* ks, const float* const* k, const double* coef, int nk) {
  FTB_CHECK(nk >= 1 && nk <= 8 && k && coef, "adaptive RK: 1..8 stage derivatives");
  ks->n = nk;
  for (int j = 0; j < 8; ++j) { ks->k[j] = j < nk ? k[j] : nullptr; ks->c[j] = j < nk ? (float)coef[j] : 0.f; }
  for (int j = 0; j < nk; ++j) FTB_CHECK(k[j] != nullptr || coef[j] == 0.0, "adaptive RK: null stage with a non-zero weight");
  return 0;
}


// ------------------------------------------------------------------------------------------------------------
// Device-resident step controller: t, dt, the accept / reject decision, the output cursor and the
// counters live in ONE device struct, so a whole adaptive solve is enqueued without a device->host read.  The host
// only watches an asynchronously copied status word of an OLDER step to know when to stop enqueueing.
// Layout of ftb_ode_ctl (doubles; documented in include/ftb.h):
//   [0] t   [1] dt   [2] acc (error-ratio sum of squares)   [3..5] d0^2 n, d1^2 n, d2^2 n accumulators of the first step
//   [6] tp0 [7] tp1 [8] dtp: interval and step size of the last ACCEPTED step (dense output)
//   [9] accepted [10] rejected [11] out_lo [12] out_hi (outputs to emit for this step: [out_lo, out_hi))
//   [13] flags: bit0 accept (this step), bit1 done, bit2 non-finite error estimate, bit3 max_num_steps exceeded
//   [14] h0 (first-step heuristic scratch)   [15] attempted steps
enum { C_T = 0, C_DT, C_ACC, C_D0, C_D1, C_D2, C_TP0, C_TP1, C_DTP, C_NACC, C_NREJ, C_OLO, C_OHI, C_FLAGS, C_H0, C_NATT, C_N };

__global__ void ctl_init_kernel(double* ctl, double t0) {
  for (int i = threadIdx.x; i < C_N; i += blockDim.x) ctl[i] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) { ctl[C_T] = t0; ctl[C_OLO] = 1.0; ctl[C_OHI] = 1.0; }   // output 0 is y0 itself
}

// _select_initial_step, first half: h0 from d0 = rms(y0 / scale), d1 = rms(f0 / scale); left in dt for the trial point
__global__ void ctl_h0_kernel(double* ctl, double n) {
  const double d0 = sqrt(ctl[C_D0] / n), d1 = sqrt(ctl[C_D1] / n);
  const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
  ctl[C_H0] = h0;
  ctl[C_DT] = h0;
}
// second half: d2 = rms((f1 - f0) / scale) / h0; h1; dt = min(100 h0, h1)
__global__ void ctl_h1_kernel(double* ctl, double n, int order) {
  const double h0 = ctl[C_H0];
  const double d1 = sqrt(ctl[C_D1] / n), d2 = sqrt(ctl[C_D2] / n) / h0;
  const double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / fmax(d1, d2), 1.0 / (double)order);
  ctl[C_DT] = fmin(100.0 * h0, h1);
}

// T[b] = float(t + alpha dt)  (t1 = t + dt for alpha == 1, the same expression in double)
__global__ void ctl_stage_time_kernel(float* tbuf, const double* ctl, double alpha, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) tbuf[i] = (float)(ctl[C_T] + alpha * ctl[C_DT]);
}

struct KSetD {
  const float* k[8];
  double c[8];     // tableau weights, NOT yet multiplied by dt
  int n;
};

// out = y0 + sum_j (float)(c_j dt) k_j with dt read from the controller (same rounding as the host-side product)
__global__ void __launch_bounds__(256)
lincomb_dev_kernel(float* __restrict__ out, const float* __restrict__ y0, KSetD ks, const double* __restrict__ ctl, size_t n4) {
  const double dt = ctl[C_DT];
  float c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = j < ks.n ? (float)(ks.c[j] * dt) : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = __ldg(reinterpret_cast<const float4*>(y0) + i);
#pragma unroll 8
    for (int j = 0; j < ks.n; ++j) {
      if (c[j] == 0.f) continue;
      const float4 k = __ldg(reinterpret_cast<const float4*>(ks.k[j]) + i);
      a.x = fmaf(k.x, c[j], a.x); a.y = fmaf(k.y, c[j], a.y);
      a.z = fmaf(k.z, c[j], a.z); a.w = fmaf(k.w, c[j], a.w);
    }
    reinterpret_cast<float4*>(out)[i] = a;
  }
}

__global__ void __launch_bounds__(256)
error_ratio_dev_kernel(const float* __restrict__ y0, const float* __restrict__ y1, KSetD ks, float rtol, float atol,
                       size_t n4, double* __restrict__ ctl) {
  __shared__ double red[8];
  const double dt = ctl[C_DT];
  float c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) c[j] = j < ks.n ? (float)(ks.c[j] * dt) : 0.f;
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(y0) + i), b = __ldg(reinterpret_cast<const float4*>(y1) + i);
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int j = 0; j < ks.n; ++j) {
      if (c[j] == 0.f) continue;
      const float4 k = __ldg(reinterpret_cast<const float4*>(ks.k[j]) + i);
      e.x = fmaf(k.x, c[j], e.x); e.y = fmaf(k.y, c[j], e.y);
      e.z = fmaf(k.z, c[j], e.z); e.w = fmaf(k.w, c[j], e.w);
    }
    const float rx = e.x / (atol + rtol * fmaxf(fabsf(a.x), fabsf(b.x)));
    const float ry = e.y / (atol + rtol * fmaxf(fabsf(a.y), fabsf(b.y)));
    const float rz = e.z / (atol + rtol * fmaxf(fabsf(a.z), fabsf(b.z)));
    const float rw = e.w / (atol + rtol * fmaxf(fabsf(a.w), fabsf(b.w)));
    s += (double)rx * rx + (double)ry * ry + (double)rz * rz + (double)rw * rw;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(ctl + C_ACC, t);
  }
}

// accept / reject, next dt (_optimal_step_size), which outputs this step emits; one thread
__global__ void ctl_step_kernel(double* ctl, const double* __restrict__ grid, int n_out, double n, int order,
                                double max_steps) {
  int flags = (int)ctl[C_FLAGS] & ~1;
  const double acc = ctl[C_ACC];
  ctl[C_ACC] = 0.0;
  ctl[C_OLO] = ctl[C_OHI];
  if (flags & (2 | 4 | 8)) { ctl[C_FLAGS] = (double)flags; return; }   // finished (or failed): later steps are no-ops
  const double ratio = sqrt(acc / n);
  if (!(ratio == ratio) || isinf(ratio)) { ctl[C_FLAGS] = (double)(flags | 4); return; }
  ctl[C_NATT] += 1.0;
  const double t = ctl[C_T], dt = ctl[C_DT];
  if (ratio <= 1.0) {
    const double t1 = t + dt;
    ctl[C_TP0] = t; ctl[C_TP1] = t1; ctl[C_DTP] = dt;
    ctl[C_T] = t1;
    ctl[C_NACC] += 1.0;
    flags |= 1;
    int hi = (int)ctl[C_OHI];
    while (hi < n_out && !(grid[hi] > t1)) ++hi;      // torchdiffeq integrates `while next_t > t1`: emit grid[hi] <= t1
    ctl[C_OHI] = (double)hi;
    if (hi >= n_out) flags |= 2;
  } else {
    ctl[C_NREJ] += 1.0;
  }
  double factor;
  if (ratio == 0.0) factor = 10.0;
  else {
    const double dfactor = ratio < 1.0 ? 1.0 : 0.2;
    factor = fmin(10.0, fmax(0.9 / pow(ratio, 1.0 / (double)order), dfactor));
  }
  ctl[C_DT] = dt * factor;
  if (!(flags & 2) && ctl[C_NATT] >= max_steps) flags |= 8;
  ctl[C_FLAGS] = (double)flags;
}

// After an accepted step: emit the outputs that fall inside it (quartic dense output through y0, y_mid, y1 with end
// slopes f0, f1; y_mid = y0 + sum_j (c_mid_j dt) k_j), then advance the state in place: y0 <- y1, f0 <- f1.
// A rejected (or post-finish) step leaves everything untouched.  traj == nullptr: only `last` (output n_out - 1).
__global__ void __launch_bounds__(256)
advance_kernel(float* __restrict__ y0, float* __restrict__ f0, const float* __restrict__ y1, const float* __restrict__ f1,
               KSetD mid, const double* __restrict__ ctl, const double* __restrict__ grid, int n_out,
               float* __restrict__ traj, float* __restrict__ last, size_t n) {
  const int flags = (int)ctl[C_FLAGS];
  if (!(flags & 1)) return;
  const int lo = (int)ctl[C_OLO], hi = (int)ctl[C_OHI];
  const double tp0 = ctl[C_TP0], tp1 = ctl[C_TP1];
  const float dt = (float)ctl[C_DTP];
  float cm[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) cm[j] = j < mid.n ? (float)(mid.c[j] * ctl[C_DTP]) : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a0 = y0[i], a1 = __ldg(y1 + i), g0 = f0[i], g1 = __ldg(f1 + i);
    if (hi > lo) {
      float am = fmaf(g0, cm[0], a0);   // k[0] IS f0 (first-same-as-last): read above, overwritten below
#pragma unroll 8
      for (int j = 1; j < mid.n; ++j)
        if (cm[j] != 0.f) am = fmaf(__ldg(mid.k[j] + i), cm[j], am);
      const float a = 2.f * dt * (g1 - g0) - 8.f * (a1 + a0) + 16.f * am;
      const float b = dt * (5.f * g0 - 3.f * g1) + 18.f * a0 + 14.f * a1 - 32.f * am;
      const float c = dt * (g1 - 4.f * g0) - 11.f * a0 - 5.f * a1 + 16.f * am;
      const float d = dt * g0;
      for (int gi = lo; gi < hi; ++gi) {
        const float x = (float)((grid[gi] - tp0) / (tp1 - tp0));
        const float v = a0 + x * (d + x * (c + x * (b + x * a)));
        if (traj) traj[(size_t)gi * n + i] = v;
        else if (gi == n_out - 1) last[i] = v;
      }
    }
    y0[i] = a1;
    f0[i] = g1;
  }
}

int make_ksetd(KSetD* ks, const float* const* k, const double* coef, int nk) {
  FTB_CHECK(nk >= 1 && nk <= 8 && k && coef, "adaptive RK: 1..8 stage derivatives");
  ks->n = nk;
  for (int j = 0; j < 8; ++j) { ks->k[j] = j < nk ? k[j] : nullptr; ks->c[j] = j < nk ? coef[j] : 0.0; }
  for (int j = 0; j < nk; ++j) FTB_CHECK(k[j] != nullptr || coef[j] == 0.0, "adaptive RK: null stage with a non-zero weight");
  return 0;
}

}  // namespace

int ode_ctl_init(double* ctl, double t0, cudaStream_t st) {
  ctl_init_kernel<<<1, 32, 0, st>>>(ctl, t0);
  FTB_LAUNCH_OK();
  return 0;
}
int ode_ctl_first_step(double* ctl, int phase, long long n, int order, cudaStream_t st) {
  if (phase == 0) ctl_h0_kernel<<<1, 1, 0, st>>>(ctl, (double)n);
  else ctl_h1_kernel<<<1, 1, 0, st>>>(ctl, (double)n, order);
  FTB_LAUNCH_OK();
  return 0;
}
int ode_ctl_stage_time(float* tbuf, const double* ctl, double alpha, int B, cudaStream_t st) {
  ctl_stage_time_kernel<<<(B + 127) / 128, 128, 0, st>>>(tbuf, ctl, alpha, B);
  FTB_LAUNCH_OK();
  return 0;
}
int ode_lincomb_dev(float* out, const float* y0, const float* const* k, const double* coef, int nk, long long n,
                    const double* ctl, cudaStream_t st) {
  FTB_CHECK(n % 4 == 0, "adaptive RK: element count must be a multiple of 4");
  KSetD ks;
  FTB_TRY(make_ksetd(&ks, k, coef, nk));
  lincomb_dev_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>(out, y0, ks, ctl, (size_t)n / 4);
  FTB_LAUNCH_OK();
  return 0;
}
int ode_error_ratio_dev(const float* y0, const float* y1, const float* const* k, const double* coef, int nk, float rtol,
                        float atol, long long n, double* ctl, cudaStream_t st) {
  FTB_CHECK(n % 4 == 0, "adaptive RK: element count must be a multiple of 4");
  KSetD ks;
  FTB_TRY(make_ksetd(&ks, k, coef, nk));
  error_ratio_dev_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>(y0, y1, ks, rtol, atol, (size_t)n / 4, ctl);
  FTB_LAUNCH_OK();
  return 0;
}
int ode_ctl_step(double* ctl, const double* grid, int n_out, long long n, int order, long long max_steps, cudaStream_t st) {
  ctl_step_kernel<<<1, 1, 0, st>>>(ctl, grid, n_out, (double)n, order, (double)max_steps);
  FTB_LAUNCH_OK();
  return 0;
}
int ode_advance(float* y0, float* f0, const float* y1, const float* f1, const float* const* k, const double* c_mid, int nk,
                const double* ctl, const double* grid, int n_out, float* traj, float* last, long long n, cudaStream_t st) {
  FTB_CHECK(traj || last, "adaptive RK: an output buffer is required");
  FTB_CHECK(k && k[0] == f0, "adaptive RK: stage 0 must be f0 (it is advanced in place)");
  KSetD ks;
  FTB_TRY(make_ksetd(&ks, k, c_mid, nk));
  advance_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(y0, f0, y1, f1, ks, ctl, grid, n_out, traj, last, (size_t)n);
  FTB_LAUNCH_OK();
  return 0;
}

int ode_lincomb(float* out, const float* y0, const float* const* k, const double* coef, int nk, long long n,
                cudaStream_t st) {
  FTB_CHECK(n % 4 == 0, "adaptive RK: element count must be a multiple of 4");
  KSet ks;
  FTB_TRY(make_kset(&ks, k, coef, nk));
  lincomb_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>(out, y0, ks, (size_t)n / 4);
  FTB_LAUNCH_OK();
  return 0;
}

int ode_error_ratio(const float* y0, const float* y1, const float* const* k, const double* coef, int nk, float rtol,
                    float atol, long long n, double* acc, cudaStream_t st) {
  FTB_CHECK(n % 4 == 0, "adaptive RK: element count must be a multiple of 4");
  KSet ks;
  FTB_TRY(make_kset(&ks, k, coef, nk));
  error_ratio_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>(y0, y1, ks, rtol, atol, (size_t)n / 4, acc);
  FTB_LAUNCH_OK();
  return 0;
}

int ode_scaled_sumsq(const float* a1, const float* a2, const float* y, float rtol, float atol, long long n, double* acc,
                     cudaStream_t st) {
  scaled_sumsq_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(a1, a2, y, rtol, atol, (size_t)n, acc);
  FTB_LAUNCH_OK();
  return 0;
}

int ode_dense_eval(float* out, const float* y0, const float* y1, const float* ymid, const float* f0, const float* f1,
                   double dt, double x, long long n, cudaStream_t st) {
  dense_eval_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(out, y0, y1, ymid, f0, f1, (float)dt, (float)x, (size_t)n);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
