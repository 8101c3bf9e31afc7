// Conditional-project kernels either side of the velocity field:
//  * conditioning front-end: surface + borehole mask, X1 = embed(batch), ATb = X1 * mask in ONE pass
//    (project/geodata-3d-conditional/boreholes.py:45-126, model_train_sh_inference_cond.py:414-420) — the reference
//    builds the mask with Python loops and one .item() host sync per borehole;
//  * loss of the conditional training step and its gradient w.r.t. the network output
//    (model_train_sh_inference_cond.py:432-452);
//  * ensemble statistics: decode -> per-voxel vote histogram in one pass (the decoded volumes never need to
//    reach HBM), then probabilities / entropy / most probable category
//    (project/geodata-3d-conditional/model_inference_experiments.py:442-457, inference_demo.ipynb cell 21).
// All HBM-bound, one thread per voxel with the innermost (Z) axis fastest, grid-stride over 16 blocks per SM.
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

// mask[b][x][y][z] = (z == Z-1)                      top slice          (make_surface_mask, boreholes.py:103)
//                 | cat == -1 | cat[z+1] == -1       air and the voxel below it   (:104-110)
//                 | (x, y) is a borehole of sample b  full-depth column  (make_boreholes_mask, :66-73)
__global__ void __launch_bounds__(256)
cond_frontend_kernel(const long long* __restrict__ cats, const int* __restrict__ bores, const int* __restrict__ nb,
                     int max_b, const float* __restrict__ w, int E, int ncat, int shift, int X, int Y, int Z,
                     int surface, unsigned char* __restrict__ mask, float* __restrict__ x1, float* __restrict__ atb) {
  extern __shared__ int s_pts[];   // [cnt][2] borehole (x, y) of this sample
  const int b = blockIdx.y;
  const size_t n = (size_t)X * Y * Z;
  int cnt = nb ? nb[b] : 0;
  cnt = cnt < 0 ? 0 : (cnt > max_b ? max_b : cnt);
  for (int i = threadIdx.x; i < 2 * cnt; i += blockDim.x) s_pts[i] = bores[(size_t)b * max_b * 2 + i];
  __syncthreads();
  const long long* cb = cats + (size_t)b * n;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(v % Z);
    const int y = (int)((v / Z) % Y);
    const int x = (int)(v / ((size_t)Z * Y));
    const long long c = cb[v];
    bool m = surface && ((z == Z - 1) || c == -1 || (z + 1 < Z && cb[v + 1] == -1));
    for (int i = 0; i < cnt && !m; ++i) m = (s_pts[2 * i] == x) && (s_pts[2 * i + 1] == y);
    if (mask) mask[(size_t)b * n + v] = m ? 1 : 0;
    long long ci = c + shift;   // embed(): indices = x + 1 (model_train_sh_inference_cond.py:352)
    ci = ci < 0 ? 0 : (ci >= ncat ? ncat - 1 : ci);
    const float mf = m ? 1.f : 0.f;
    for (int e = 0; e < E && w; ++e) {
      const float val = __ldg(w + ci * E + e);
      const size_t o = ((size_t)b * E + e) * n + v;
      if (x1) x1[o] = val;
      if (atb) atb[o] = val * mf;   // X1 * mask (:420): a true product, so the sign of the zeros matches too
    }
  }
}

// acc[0] += sum (v - vh)^2, acc[1] += sum v^2, acc[2] += sum_mask (b - b_hat)^2, acc[3] += #masked elements,
// acc[4] += sum x1n^2, acc[5] += sum_b T[b];  b_hat = XT + (1 - T) * VT_hat on the mask (:433-436)
__global__ void __launch_bounds__(256)
cond_loss_kernel(const float* __restrict__ vt, const float* __restrict__ vh, const float* __restrict__ xt,
                 const float* __restrict__ x1c, const float* __restrict__ x1n, const unsigned char* __restrict__ mask,
                 const float* __restrict__ T, int B, int E, size_t n, double* __restrict__ acc) {
  __shared__ double red[5][8];
  double s[5] = {0, 0, 0, 0, 0};
  const size_t total = (size_t)B * E * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % n;
    const int b = (int)(i / (n * E));
    const float a = __ldg(vt + i), h = __ldg(vh + i), d = a - h, xn = __ldg(x1n + i);
    s[0] += (double)d * d;
    s[1] += (double)a * a;
    s[4] += (double)xn * xn;
    if (mask[(size_t)b * n + v]) {
      const float bh = __ldg(xt + i) + (1.f - __ldg(T + b)) * h;
      const float r = __ldg(x1c + i) - bh;
      s[2] += (double)r * r;
      s[3] += 1.0;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double t = s[k];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) red[k][wid] = t;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0;
    for (int wv = 0; wv < 8; ++wv) t += red[threadIdx.x][wv];
    atomicAdd(acc + threadIdx.x, t);
  }
  if (blockIdx.x == 0 && threadIdx.x == 5) {
    double t = 0;
    for (int b = 0; b < B; ++b) t += (double)T[b];
    atomicAdd(acc + 5, t);
  }
}

// d loss / d VT_hat for loss = mse(VT, VT_hat) / (mse(VT, 0) + 1e-6)
//                            + lambda * mean(T) * mse(b, b_hat) / (mse(X1, 0) + 1e-6)          (:438-451)
__global__ void __launch_bounds__(256)
cond_loss_grad_kernel(const float* __restrict__ vt, const float* __restrict__ vh, const float* __restrict__ xt,
                      const float* __restrict__ x1c, const unsigned char* __restrict__ mask,
                      const float* __restrict__ T, int B, int E, size_t n, const double* __restrict__ acc,
                      float lambda, float scale, float* __restrict__ dout) {
  const size_t total = (size_t)B * E * n;
  const double N = (double)total;
  const float c0 = (float)((double)scale * 2.0 / (N * (acc[1] / N + 1e-6)));
  const float c1 = acc[3] > 0.0
                       ? (float)((double)scale * (double)lambda * (acc[5] / B) * 2.0 / (acc[3] * (acc[4] / N + 1e-6)))
                       : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % n;
    const int b = (int)(i / (n * E));
    const float h = __ldg(vh + i);
    float g = c0 * (h - __ldg(vt + i));
    if (mask[(size_t)b * n + v]) {
      const float omt = 1.f - __ldg(T + b);
      const float bh = __ldg(xt + i) + omt * h;
      g += c1 * (bh - __ldg(x1c + i)) * omt;
    }
    dout[i] = g;
  }
}

// probabilities p[c] = count[c] / S (one-hot mean, :442-447), entropy = -sum p log(p + 1e-8) (:449-452),
// most probable = first argmax - 1 (:454-455), entropy_masked = -1 where the most probable category is air (:458-459)
__global__ void __launch_bounds__(256)
vote_finalize_kernel(const int* __restrict__ counts, int S, int ncat, size_t n, int shift, float* __restrict__ probs,
                     float* __restrict__ entropy, long long* __restrict__ most, float* __restrict__ entropy_masked) {
  const float fs = (float)S;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
    float H = 0.f;
    int best = -1, arg = 0;
    for (int c = 0; c < ncat; ++c) {
      const int k = __ldg(counts + (size_t)c * n + v);
      const float p = __fdiv_rn((float)k, fs);
      if (probs) probs[(size_t)c * n + v] = p;
      H = __fadd_rn(H, __fmul_rn(p, logf(__fadd_rn(p, 1e-8f))));
      if (k > best) { best = k; arg = c; }
    }
    H = -H;
    if (entropy) entropy[v] = H;
    if (most) most[v] = (long long)arg + shift;
    if (entropy_masked) entropy_masked[v] = (arg + shift == -1) ? -1.f : H;
  }
}

}  // namespace

int cond_frontend(const long long* cats, const int* bores, const int* nb, int max_b, const float* w, int B, int E,
                  int ncat, int shift, int X, int Y, int Z, int surface, unsigned char* mask, float* x1, float* atb,
                  cudaStream_t st) {
  FTB_CHECK(max_b >= 0 && max_b <= 4096, "cond_frontend: at most 4096 boreholes per sample");
  FTB_CHECK(!w || (E >= 1 && ncat >= 1), "cond_frontend: embedding shape");
  const size_t n = (size_t)X * Y * Z;
  int gx = grid_for(n, 256) / (B > 0 ? B : 1);
  gx = gx < 1 ? 1 : gx;
  cond_frontend_kernel<<<dim3(gx, B), 256, (size_t)(max_b > 0 ? max_b : 1) * 2 * sizeof(int), st>>>(
      cats, bores, nb, max_b, w, E, ncat, shift, X, Y, Z, surface, mask, x1, atb);
  FTB_LAUNCH_OK();
  return 0;
}

int cond_loss_partial(const float* vt, const float* vh, const float* xt, const float* x1c, const float* x1n,
                      const unsigned char* mask, const float* T, int B, int E, long long n, double* acc6,
                      cudaStream_t st) {
  cond_loss_kernel<<<grid_for((size_t)B * E * n, 256), 256, 0, st>>>(vt, vh, xt, x1c, x1n, mask, T, B, E, (size_t)n, acc6);
  FTB_LAUNCH_OK();
  return 0;
}

int cond_loss_grad(const float* vt, const float* vh, const float* xt, const float* x1c, const unsigned char* mask,
                   const float* T, int B, int E, long long n, const double* acc6, float lambda, float scale, float* dout,
                   cudaStream_t st) {
  cond_loss_grad_kernel<<<grid_for((size_t)B * E * n, 256), 256, 0, st>>>(vt, vh, xt, x1c, mask, T, B, E, (size_t)n, acc6,
                                                                         lambda, scale, dout);
  FTB_LAUNCH_OK();
  return 0;
}

int vote_finalize(const int* counts, int S, int ncat, long long n, int shift, float* probs, float* entropy,
                  long long* most, float* entropy_masked, cudaStream_t st) {
  FTB_CHECK(S >= 1 && ncat >= 1, "vote_finalize: need at least one sample and one category");
  vote_finalize_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(counts, S, ncat, (size_t)n, shift, probs, entropy, most,
                                                                entropy_masked);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
