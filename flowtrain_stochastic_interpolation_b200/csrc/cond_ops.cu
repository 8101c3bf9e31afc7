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
template <int VEC>   // voxels per thread along z (4: Z % 4 == 0, 16-byte stores; 1: any Z)
__global__ void __launch_bounds__(256)
cond_frontend_kernel(const long long* __restrict__ cats, const int* __restrict__ bores, const int* __restrict__ nb,
                     int max_b, const float* __restrict__ w, int E, int ncat, int shift, int X, int Y, int Z,
                     int surface, unsigned char* __restrict__ mask, float* __restrict__ x1, float* __restrict__ atb) {
  extern __shared__ int s_pts[];   // [cnt][2] borehole (x, y) of this sample, then the embedding matrix [ncat][E]
  const int b = blockIdx.y;
  const size_t n = (size_t)X * Y * Z;
  int cnt = nb ? nb[b] : 0;
  cnt = cnt < 0 ? 0 : (cnt > max_b ? max_b : cnt);
  float* s_w = reinterpret_cast<float*>(s_pts + 2 * (max_b > 0 ? max_b : 1));
  for (int i = threadIdx.x; i < 2 * cnt; i += blockDim.x) s_pts[i] = bores[(size_t)b * max_b * 2 + i];
  if (w) for (int i = threadIdx.x; i < ncat * E; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const long long* cb = cats + (size_t)b * n;
  const size_t nv = n / VEC;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nv; q += (size_t)gridDim.x * blockDim.x) {
    const size_t v = q * VEC;
    const int z = (int)(v % Z);
    const int y = (int)((v / Z) % Y);
    const int x = (int)(v / ((size_t)Z * Y));
    long long c[VEC + 1];
#pragma unroll
    for (int k = 0; k < VEC; ++k) c[k] = cb[v + k];
    c[VEC] = (z + VEC < Z) ? cb[v + VEC] : 0;   // the voxel above the last one (same column) or "not air"
    bool bore = false;
    for (int i = 0; i < cnt && !bore; ++i) bore = (s_pts[2 * i] == x) && (s_pts[2 * i + 1] == y);
    bool m[VEC];
    long long ci[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      m[k] = bore || (surface && ((z + k == Z - 1) || c[k] == -1 || c[k + 1] == -1));
      long long t = c[k] + shift;   // embed(): indices = x + 1 (model_train_sh_inference_cond.py:352)
      ci[k] = t < 0 ? 0 : (t >= ncat ? ncat - 1 : t);
    }
    if (mask) {
      if (VEC == 4) *reinterpret_cast<uchar4*>(mask + (size_t)b * n + v) = make_uchar4(m[0], m[1], m[2], m[3]);
      else mask[(size_t)b * n + v] = m[0] ? 1 : 0;
    }
    if (!w) continue;
    for (int e = 0; e < E; ++e) {
      float val[VEC], ma[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        val[k] = s_w[ci[k] * E + e];
        ma[k] = val[k] * (m[k] ? 1.f : 0.f);   // X1 * mask (:420): a true product, so the sign of the zeros matches too
      }
      const size_t o = ((size_t)b * E + e) * n + v;
      if (VEC == 4) {
        if (x1) *reinterpret_cast<float4*>(x1 + o) = make_float4(val[0], val[1], val[2], val[3]);
        if (atb) *reinterpret_cast<float4*>(atb + o) = make_float4(ma[0], ma[1], ma[2], ma[3]);
      } else {
        if (x1) x1[o] = val[0];
        if (atb) atb[o] = ma[0];
      }
    }
  }
}

// acc[0] += sum (v - vh)^2, acc[1] += sum v^2, acc[2] += sum_mask (b - b_hat)^2, acc[3] += #masked elements,
// acc[4] += sum x1n^2, acc[5] += sum_b T[b];  b_hat = XT + (1 - T) * VT_hat on the mask (:433-436)
template <int VEC>   // elements per thread (4 when n % 4 == 0: 16-byte loads; else 1)
__global__ void __launch_bounds__(256)
cond_loss_kernel(const float* __restrict__ vt, const float* __restrict__ vh, const float* __restrict__ xt,
                 const float* __restrict__ x1c, const float* __restrict__ x1n, const unsigned char* __restrict__ mask,
                 const float* __restrict__ T, int B, int E, size_t n, double* __restrict__ acc) {
  __shared__ double red[5][8];
  double s[5] = {0, 0, 0, 0, 0};
  const size_t total = (size_t)B * E * n / VEC;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    const size_t i = q * VEC;
    const size_t v = i % n;
    const int b = (int)(i / (n * E));
    __align__(16) float a[VEC], h[VEC], xn[VEC];
    __align__(4) unsigned char m[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(a) = __ldg(reinterpret_cast<const float4*>(vt + i));
      *reinterpret_cast<float4*>(h) = __ldg(reinterpret_cast<const float4*>(vh + i));
      *reinterpret_cast<float4*>(xn) = __ldg(reinterpret_cast<const float4*>(x1n + i));
      *reinterpret_cast<uchar4*>(m) = __ldg(reinterpret_cast<const uchar4*>(mask + (size_t)b * n + v));
    } else {
      a[0] = __ldg(vt + i); h[0] = __ldg(vh + i); xn[0] = __ldg(x1n + i); m[0] = mask[(size_t)b * n + v];
    }
    float f0 = 0.f, f1 = 0.f, f4 = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float d = a[k] - h[k];
      f0 = fmaf(d, d, f0); f1 = fmaf(a[k], a[k], f1); f4 = fmaf(xn[k], xn[k], f4);
    }
    s[0] += f0; s[1] += f1; s[4] += f4;
    bool any = false;
#pragma unroll
    for (int k = 0; k < VEC; ++k) any |= m[k] != 0;
    if (any) {
      const float omt = 1.f - __ldg(T + b);
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        if (m[k]) {
          const float bh = __ldg(xt + i + k) + omt * h[k];
          const float r = __ldg(x1c + i + k) - bh;
          s[2] += (double)r * r;
          s[3] += 1.0;
        }
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
template <int VEC>
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
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total / VEC; q += (size_t)gridDim.x * blockDim.x) {
    const size_t i = q * VEC;
    const size_t v = i % n;
    const int b = (int)(i / (n * E));
    __align__(16) float a[VEC], h[VEC], g[VEC];
    __align__(4) unsigned char m[VEC];
    if (VEC == 4) {
      *reinterpret_cast<float4*>(a) = __ldg(reinterpret_cast<const float4*>(vt + i));
      *reinterpret_cast<float4*>(h) = __ldg(reinterpret_cast<const float4*>(vh + i));
      *reinterpret_cast<uchar4*>(m) = __ldg(reinterpret_cast<const uchar4*>(mask + (size_t)b * n + v));
    } else {
      a[0] = __ldg(vt + i); h[0] = __ldg(vh + i); m[0] = mask[(size_t)b * n + v];
    }
    const float omt = 1.f - __ldg(T + b);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      g[k] = c0 * (h[k] - a[k]);
      if (m[k]) {
        const float bh = __ldg(xt + i + k) + omt * h[k];
        g[k] += c1 * (bh - __ldg(x1c + i + k)) * omt;
      }
    }
    if (VEC == 4) *reinterpret_cast<float4*>(dout + i) = *reinterpret_cast<float4*>(g);
    else dout[i] = g[0];
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
  const bool vec = Z % 4 == 0;
  int gx = grid_for(n / (vec ? 4 : 1), 256) / (B > 0 ? B : 1);
  gx = gx < 1 ? 1 : gx;
  const size_t smem = (size_t)(max_b > 0 ? max_b : 1) * 2 * sizeof(int) + (w ? (size_t)ncat * E * sizeof(float) : 0);
  FTB_CHECK(smem <= 48 * 1024, "cond_frontend: borehole table + embedding matrix exceed 48 KB of shared memory");
  if (vec) cond_frontend_kernel<4><<<dim3(gx, B), 256, smem, st>>>(cats, bores, nb, max_b, w, E, ncat, shift, X, Y, Z,
                                                                   surface, mask, x1, atb);
  else cond_frontend_kernel<1><<<dim3(gx, B), 256, smem, st>>>(cats, bores, nb, max_b, w, E, ncat, shift, X, Y, Z,
                                                                surface, mask, x1, atb);
  FTB_LAUNCH_OK();
  return 0;
}

int cond_loss_partial(const float* vt, const float* vh, const float* xt, const float* x1c, const float* x1n,
                      const unsigned char* mask, const float* T, int B, int E, long long n, double* acc6,
                      cudaStream_t st) {
  if (n % 4 == 0)
    cond_loss_kernel<4><<<grid_for((size_t)B * E * n / 4, 256), 256, 0, st>>>(vt, vh, xt, x1c, x1n, mask, T, B, E, (size_t)n, acc6);
  else
    cond_loss_kernel<1><<<grid_for((size_t)B * E * n, 256), 256, 0, st>>>(vt, vh, xt, x1c, x1n, mask, T, B, E, (size_t)n, acc6);
  FTB_LAUNCH_OK();
  return 0;
}

int cond_loss_grad(const float* vt, const float* vh, const float* xt, const float* x1c, const unsigned char* mask,
                   const float* T, int B, int E, long long n, const double* acc6, float lambda, float scale, float* dout,
                   cudaStream_t st) {
  if (n % 4 == 0)
    cond_loss_grad_kernel<4><<<grid_for((size_t)B * E * n / 4, 256), 256, 0, st>>>(vt, vh, xt, x1c, mask, T, B, E, (size_t)n,
                                                                                  acc6, lambda, scale, dout);
  else
    cond_loss_grad_kernel<1><<<grid_for((size_t)B * E * n, 256), 256, 0, st>>>(vt, vh, xt, x1c, mask, T, B, E, (size_t)n,
                                                                              acc6, lambda, scale, dout);
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
