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

}  // namespace

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
