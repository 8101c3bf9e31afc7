// Bandwidth kernels: layout conversion, trilinear resample, interpolant construction,
// integrator updates, eq-6.7 drift, categorical decode/embed, EMA, loss partials.
// All are coalesced, 16-byte vectorised where the layout allows, grid-stride over
// (a multiple of) the SM count.  Reference lines are cited at each kernel.
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

// ------------------------------------------------------------------ NCDHW fp32 <-> blocked bf16
__global__ void pack_kernel(const float* __restrict__ x, int B, int C, size_t vox, int CG,
                            bf16* __restrict__ out) {
  const size_t total = (size_t)B * CG * vox;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % vox;
    const int cg = (int)((i / vox) % CG);
    const int b = (int)(i / (vox * CG));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      f[j] = c < C ? __ldg(x + ((size_t)b * C + c) * vox + v) : 0.f;
    }
    *reinterpret_cast<uint4*>(out + i * 8) = pack_bf16x8(f);
  }
}

// One block per (b, d, group of kUnfRows h rows): the C fp32 rows of every h are staged in shared memory with a zero
// halo, then every (channel group, w) output reads its 8 (kw, c) taps from there (coalesced both ways).  Several rows
// per block amortise the per-block set-up (tap tables, barrier): with one row per block the 32 768 small blocks of a
// 64^3 B=8 input ran at 3.3 TB/s.
constexpr int kUnfRows = 4;
__global__ void __launch_bounds__(256)
pack_unfold_w_kernel(const float* __restrict__ x, int C, int D, int H, int W, int K, int CG,
                     bf16* __restrict__ out) {
  extern __shared__ float s_row[];                  // [kUnfRows][C][W + K - 1]
  __shared__ unsigned char s_kw[256], s_c[256];     // per unfolded channel: tap and source channel
  const int pad = K / 2, RW = W + K - 1;
  const int hgroups = (H + kUnfRows - 1) / kUnfRows;
  int blk = blockIdx.x;
  const int hg = blk % hgroups; blk /= hgroups;
  const int d = blk % D;
  const int b = blk / D;
  const int h0 = hg * kUnfRows;
  const int nh = min(kUnfRows, H - h0);
  const size_t vox = (size_t)D * H * W;
  for (int i = threadIdx.x; i < CG * 8; i += blockDim.x) {
    const int kw = i / C;
    s_kw[i] = (unsigned char)(kw < K ? kw : 255);
    s_c[i] = (unsigned char)(i - kw * C);
  }
  for (int i = threadIdx.x; i < nh * C * RW; i += blockDim.x) {
    const int r = i / (C * RW), j = i - r * (C * RW);
    const int c = j / RW, wr = j - c * RW, w = wr - pad;
    s_row[i] = (w >= 0 && w < W) ? __ldg(x + ((size_t)b * C + c) * vox + ((size_t)d * H + h0 + r) * W + w) : 0.f;
  }
  __syncthreads();
  // thread = (channel group, W lane): its 8 (kw, c) taps are loop invariant
  const int lanes = blockDim.x / CG;            // W lanes per channel group (blockDim = CG * lanes)
  const int cg = threadIdx.x / lanes, wl = threadIdx.x - cg * lanes;
  if (cg < CG) {
    int off[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cc = cg * 8 + j;
      off[j] = s_kw[cc] == 255 ? -1 : s_c[cc] * RW + s_kw[cc];
    }
    for (int r = 0; r < nh; ++r) {
      const float* sr = s_row + (size_t)r * C * RW;
      bf16* orow = out + ((((size_t)b * CG + cg) * vox) + ((size_t)d * H + h0 + r) * W) * 8;
      for (int w = wl; w < W; w += lanes) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = off[j] < 0 ? 0.f : sr[off[j] + w];
        *reinterpret_cast<uint4*>(orow + (size_t)w * 8) = pack_bf16x8(f);
      }
    }
  }
}

__global__ void unpack_kernel(const bf16* __restrict__ in, int B, int cgtot, int cgoff, int C,
                              size_t vox, float* __restrict__ out) {
  const int CG = (C + 7) / 8;
  const size_t total = (size_t)B * CG * vox;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % vox;
    const int cg = (int)((i / vox) % CG);
    const int b = (int)(i / (vox * CG));
    float f[8];
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(in + (((size_t)b * cgtot + cgoff + cg) * vox + v) * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      if (c < C) out[((size_t)b * C + c) * vox + v] = f[j];
    }
  }
}

// ------------------------------------------------------------------ trilinear, align_corners=True
// F.interpolate(..., mode="trilinear", align_corners=True) — unet_attn_3d.py:86,:106.
// Index rule as ATen: scale = (in-1)/(out-1) (0 if out==1) in fp32, src = scale*dst,
// i0 = (int)src, i1 = i0 + (i0 < in-1), lambda1 = src - i0, lambda0 = 1 - lambda1.
struct Lerp {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ Lerp lerp_idx(int o, int in, float scale) {
  const float src = scale * (float)o;
  Lerp r;
  r.i0 = (int)src;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

// One block per (b*cg, d, group of RH output rows): the depth lerp parameters are block-uniform.  A thread produces
// up to RH*Wo/256 voxels in an unrolled loop (8 independent 16-byte loads each, so a few dozen loads in flight per
// thread): with 4 rows per block the 64^3 upsample was 98 304 one-voxel-per-thread blocks and ran at 1.1 TB/s.
constexpr int kTriRH = 16;
__global__ void __launch_bounds__(256)
trilinear_kernel(const bf16* __restrict__ in, int CG, int Di, int Hi, int Wi, int Do, int Ho, int Wo,
                 int hgroups, float sd, float sh, float sw, bf16* __restrict__ out) {
  int blk = blockIdx.x;
  const int hg = blk % hgroups; blk /= hgroups;
  const int d = blk % Do;
  const size_t bc = blk / Do;
  const Lerp ld = lerp_idx(d, Di, sd);
  const bf16* base = in + bc * (size_t)Di * Hi * Wi * 8;
  bf16* obase = out + (bc * (size_t)Do + d) * Ho * Wo * 8;
#pragma unroll 2
  for (int idx = threadIdx.x; idx < kTriRH * Wo; idx += blockDim.x) {
    const int hr = idx / Wo, w = idx - hr * Wo;
    const int h = hg * kTriRH + hr;
    if (h >= Ho) continue;
    const Lerp lh = lerp_idx(h, Hi, sh), lw = lerp_idx(w, Wi, sw);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int dd = a ? ld.i1 : ld.i0;
      const float wd = a ? ld.l1 : ld.l0;
      float accd[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) accd[j] = 0.f;
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int hh = bb ? lh.i1 : lh.i0;
        const float wh = bb ? lh.l1 : lh.l0;
        float f0[8], f1[8];
        const size_t row = ((size_t)dd * Hi + hh) * Wi;
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(base + (row + lw.i0) * 8)), f0);
        unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(base + (row + lw.i1) * 8)), f1);
#pragma unroll
        for (int j = 0; j < 8; ++j) accd[j] += wh * (lw.l0 * f0[j] + lw.l1 * f1[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += wd * accd[j];
    }
    *reinterpret_cast<uint4*>(obase + ((size_t)h * Wo + w) * 8) = pack_bf16x8(acc);
  }
}

// Upsampling variant: an upsample reads every input voxel from ~8 outputs; with direct global loads that gather
// traffic and the 8 x (unpack + lerp) per output made the 32^3 -> 64^3 resample instruction- and L2-bound (353 us for
// 453 MB).  Block = (b*cg, d, group of RH output rows): the two source planes' rows [r0, r0 + nrows) are loaded once
// (coalesced 16-byte units) and blended along D (block-uniform weights) into fp32 shared-memory rows, then every
// output is a 4-tap bilinear blend of those rows.  (fp32 throughout; the association D-then-HW differs from ATen's in
// the last bit, far below the bf16 output rounding.)
__global__ void __launch_bounds__(256)
trilinear_smem_kernel(const bf16* __restrict__ in, int Di, int Hi, int Wi, int Do, int Ho, int Wo, int hgroups,
                      int RH, int max_rows, float sd, float sh, float sw, bf16* __restrict__ out) {
  // [2 halves][max_rows][Wi]: the two source planes already blended along D (fp32, channels 0-3 | 4-7 of the group in
  // separate arrays: the taps of neighbouring outputs are then neighbouring 16-byte words, one wavefront per quarter warp)
  extern __shared__ float4 s_t[];
  const int half_stride = max_rows * Wi;
  int blk = blockIdx.x;
  const int hg = blk % hgroups; blk /= hgroups;
  const int d = blk % Do;
  const size_t bc = blk / Do;
  const Lerp ld = lerp_idx(d, Di, sd);
  const int h_lo = hg * RH, h_hi = min(Ho, h_lo + RH) - 1;
  const int r0 = lerp_idx(h_lo, Hi, sh).i0;
  const int nrows = lerp_idx(h_hi, Hi, sh).i1 - r0 + 1;
  const uint4* base = reinterpret_cast<const uint4*>(in) + bc * (size_t)Di * Hi * Wi;
  // the depth weights are block-uniform: blend the two planes once per staged voxel instead of once per output
  for (int i = threadIdx.x; i < nrows * Wi; i += blockDim.x) {
    float f0[8], f1[8];
    unpack_bf16x8(__ldg(base + ((size_t)ld.i0 * Hi + r0) * Wi + i), f0);
    unpack_bf16x8(__ldg(base + ((size_t)ld.i1 * Hi + r0) * Wi + i), f1);
#pragma unroll
    for (int k = 0; k < 8; ++k) f0[k] = ld.l0 * f0[k] + ld.l1 * f1[k];
    s_t[i] = make_float4(f0[0], f0[1], f0[2], f0[3]);
    s_t[half_stride + i] = make_float4(f0[4], f0[5], f0[6], f0[7]);
  }
  __syncthreads();
  bf16* obase = out + (bc * (size_t)Do + d) * Ho * Wo * 8;
  for (int idx = threadIdx.x; idx < RH * Wo; idx += blockDim.x) {
    const int hr = idx / Wo, w = idx - hr * Wo;
    const int h = h_lo + hr;
    if (h >= Ho) continue;
    const Lerp lh = lerp_idx(h, Hi, sh), lw = lerp_idx(w, Wi, sw);
    const float4* ra = s_t + (size_t)(lh.i0 - r0) * Wi;
    const float4* rb = s_t + (size_t)(lh.i1 - r0) * Wi;
    const float w00 = lh.l0 * lw.l0, w01 = lh.l0 * lw.l1, w10 = lh.l1 * lw.l0, w11 = lh.l1 * lw.l1;
    float acc[8];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float4 a0 = ra[hf * half_stride + lw.i0], a1 = ra[hf * half_stride + lw.i1];
      const float4 b0 = rb[hf * half_stride + lw.i0], b1 = rb[hf * half_stride + lw.i1];
      acc[4 * hf + 0] = w00 * a0.x + w01 * a1.x + w10 * b0.x + w11 * b1.x;
      acc[4 * hf + 1] = w00 * a0.y + w01 * a1.y + w10 * b0.y + w11 * b1.y;
      acc[4 * hf + 2] = w00 * a0.z + w01 * a1.z + w10 * b0.z + w11 * b1.z;
      acc[4 * hf + 3] = w00 * a0.w + w01 * a1.w + w10 * b0.w + w11 * b1.w;
    }
    *reinterpret_cast<uint4*>(obase + ((size_t)h * Wo + w) * 8) = pack_bf16x8(acc);
  }
}

// ------------------------------------------------------------------ MixATb input: cat + FiLM
// MixATb.forward (unet_attn_3d_cond_v3.py:176-182): ATb_x = cat(x, ATb); ATb_x*(scale+1)+shift.
// The FiLM acts on the conv INPUT (the zero padding of conv1 stays zero), so it cannot fold
// into the conv; one pass writes the 2C-channel tensor.  mul = scale+1 and add = shift come
// from film_mlps; the ATb embedding may have batch 1 (one conditioning volume for an ensemble).
__global__ void film_concat_kernel(const bf16* __restrict__ x, const bf16* __restrict__ a, int B, int CG,
                                   size_t vox, size_t a_bstride, const float* __restrict__ film,
                                   int film_stride, int C, bf16* __restrict__ out) {
  const size_t total = (size_t)B * 2 * CG * vox;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % vox;
    const int cg = (int)((i / vox) % (2 * CG));
    const int b = (int)(i / (vox * 2 * CG));
    const bool second = cg >= CG;
    const int cgl = second ? cg - CG : cg;
    const bf16* src = second ? a + (size_t)b * a_bstride + ((size_t)cgl * vox + v) * 8
                             : x + (((size_t)b * CG + cgl) * vox + v) * 8;
    float f[8];
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(src)), f);
    const float* mul = film + (size_t)b * film_stride + (second ? C : 0) + cgl * 8;
    const float* add = mul + 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = cgl * 8 + j < C ? fmaf(f[j], __ldg(mul + j), __ldg(add + j)) : 0.f;
    *reinterpret_cast<uint4*>(out + i * 8) = pack_bf16x8(f);
  }
}

// ------------------------------------------------------------------ interpolant (interpolation.py:156-216, :379-546)
struct Coef {
  float a, b, g, ad, bd, gd;
};
__device__ Coef interp_coef(int kind, int one_sided, float ga, float t) {
  const float pi = 3.14159265358979323846f;
  Coef c;
  c.g = 0.f;
  c.gd = 0.f;
  const float gam = sqrtf(__fmul_rn(__fmul_rn(ga, t), 1.f - t));
  const float gamd = __fmul_rn(__fmul_rn(0.5f, ga), 1.f - __fmul_rn(2.f, t)) / gam;
  switch (kind) {
    case 0:  // linear
      c.a = 1.f - t; c.b = t; c.ad = -1.f; c.bd = 1.f;
      if (!one_sided) { c.g = gam; c.gd = gamd; }
      break;
    case 1: {  // trig
      const float ph = __fmul_rn(pi, t) / 2.f;
      c.a = cosf(ph); c.b = sinf(ph);
      c.ad = __fmul_rn(-pi / 2.f, sinf(ph)); c.bd = __fmul_rn(pi / 2.f, cosf(ph));
      if (!one_sided) { c.g = gam; c.gd = gamd; }
    } break;
    case 2: {  // enc-dec
      const float cs = cosf(__fmul_rn(pi, t));
      const float c2 = __fmul_rn(cs, cs);
      const float s2 = __fmul_rn(-pi, sinf(__fmul_rn(__fmul_rn(2.f, pi), t)));
      c.a = t < 0.5f ? c2 : 0.f; c.b = t > 0.5f ? c2 : 0.f;
      const float sn = sinf(__fmul_rn(pi, t));
      c.g = __fmul_rn(sn, sn);
      c.ad = t < 0.5f ? s2 : 0.f; c.bd = t > 0.5f ? s2 : 0.f;
      c.gd = -s2;
    } break;
    case 3:  // SBDM
      c.a = sqrtf(1.f - __fmul_rn(t, t)); c.b = t;
      c.ad = -t / sqrtf(1.f - __fmul_rn(t, t)); c.bd = 1.f;
      break;
    default:  // mirror
      c.a = 0.f; c.b = 1.f; c.ad = 0.f; c.bd = 0.f; c.g = gam; c.gd = gamd;
      break;
  }
  return c;
}

__global__ void interp_kernel(int kind, int one_sided, float ga, const float4* __restrict__ x0,
                              const float4* __restrict__ x1, const float4* __restrict__ z,
                              const float* __restrict__ t, float4* __restrict__ xt,
                              float4* __restrict__ bt, int B, size_t n4) {
  const size_t total = (size_t)B * n4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / n4);
    const Coef c = interp_coef(kind, one_sided, ga, __ldg(t + b));
    const float4 u = __ldg(x0 + i), v = __ldg(x1 + i);
    float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (z) zz = __ldg(z + i);
    float4 o, d;
#define FTB_XT(f)                                                                  \
  o.f = __fadd_rn(__fmul_rn(c.a, u.f), __fmul_rn(c.b, v.f));                       \
  d.f = __fadd_rn(__fmul_rn(c.ad, u.f), __fmul_rn(c.bd, v.f));                     \
  if (z) {                                                                         \
    o.f = __fadd_rn(o.f, __fmul_rn(c.g, zz.f));                                    \
    d.f = __fadd_rn(d.f, __fmul_rn(c.gd, zz.f));                                   \
  }
    FTB_XT(x) FTB_XT(y) FTB_XT(z) FTB_XT(w)
#undef FTB_XT
    xt[i] = o;
    if (bt) bt[i] = d;
  }
}

// ------------------------------------------------------------------ integrator updates (solvers.py:236-240; fixed grid)
__global__ void axpy_kernel(float* __restrict__ out, const float* __restrict__ x,
                            const float* __restrict__ k, float h, size_t n,
                            const unsigned char* __restrict__ frozen, size_t inner) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float kv = __ldg(k + i);
    if (frozen && frozen[i % inner]) kv = 0.f;  // dxdt[..., frozen_mask] = 0 (solvers.py:73)
    out[i] = __fadd_rn(__ldg(x + i), __fmul_rn(h, kv));
  }
}
__global__ void axpy4_kernel(float4* __restrict__ out, const float4* __restrict__ x,
                             const float4* __restrict__ k, float h, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(x + i), b = __ldg(k + i);
    float4 o;
    o.x = __fadd_rn(a.x, __fmul_rn(h, b.x));
    o.y = __fadd_rn(a.y, __fmul_rn(h, b.y));
    o.z = __fadd_rn(a.z, __fmul_rn(h, b.z));
    o.w = __fadd_rn(a.w, __fmul_rn(h, b.w));
    out[i] = o;
  }
}
__global__ void heun_kernel(float4* __restrict__ out, const float4* __restrict__ x,
                            const float4* __restrict__ k1, const float4* __restrict__ k2, float hh,
                            size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(x + i), p = __ldg(k1 + i), q = __ldg(k2 + i);
    float4 o;
    o.x = __fadd_rn(a.x, __fmul_rn(hh, __fadd_rn(p.x, q.x)));
    o.y = __fadd_rn(a.y, __fmul_rn(hh, __fadd_rn(p.y, q.y)));
    o.z = __fadd_rn(a.z, __fmul_rn(hh, __fadd_rn(p.z, q.z)));
    o.w = __fadd_rn(a.w, __fmul_rn(hh, __fadd_rn(p.w, q.w)));
    out[i] = o;
  }
}
__device__ __forceinline__ float rk4_1(float x, float a, float b, float c, float d, float h6) {
  float t = __fadd_rn(a, __fmul_rn(2.f, b));
  t = __fadd_rn(t, __fmul_rn(2.f, c));
  t = __fadd_rn(t, d);
  return __fadd_rn(x, __fmul_rn(h6, t));
}
__global__ void rk4_kernel(float4* __restrict__ out, const float4* __restrict__ x,
                           const float4* __restrict__ k1, const float4* __restrict__ k2,
                           const float4* __restrict__ k3, const float4* __restrict__ k4, float h6,
                           size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(x + i), p = __ldg(k1 + i), q = __ldg(k2 + i), r = __ldg(k3 + i),
                 s = __ldg(k4 + i);
    float4 o;
    o.x = rk4_1(a.x, p.x, q.x, r.x, s.x, h6);
    o.y = rk4_1(a.y, p.y, q.y, r.y, s.y, h6);
    o.z = rk4_1(a.z, p.z, q.z, r.z, s.z, h6);
    o.w = rk4_1(a.w, p.w, q.w, r.w, s.w, h6);
    out[i] = o;
  }
}
// eq. 6.7 drift from a denoiser (solvers.py:130-143) + SDE term (:205-216)
__global__ void drift_kernel(float* __restrict__ out, const float* __restrict__ x,
                             const float* __restrict__ eta, const float* __restrict__ noise, float a,
                             float b, float ad, float bd, float eps, int use_sde, size_t n) {
  const float bdb = bd / b;
  const float sq = use_sde ? sqrtf(__fmul_rn(2.f, eps)) : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float e = __ldg(eta + i), xv = __ldg(x + i);
    float d = __fadd_rn(__fmul_rn(ad, e), __fmul_rn(bdb, __fsub_rn(xv, __fmul_rn(a, e))));
    if (use_sde) {
      const float score = -e / a;
      const float st = __fadd_rn(__fmul_rn(eps, score), __fmul_rn(__ldg(noise + i), sq));
      d = __fadd_rn(d, st);
    }
    out[i] = d;
  }
}

// same drift with (alpha, beta, alpha_dot, beta_dot, eps) read from device memory: the adaptive solvers keep time on the
// device, so the schedule values of an evaluation are computed there too (no host round trip)
__global__ void drift_dev_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ eta,
                                 const float* __restrict__ noise, const float* __restrict__ coef, int use_sde, size_t n) {
  const float a = coef[0], b = coef[1], ad = coef[2], bd = coef[3], eps = use_sde ? coef[4] : 0.f;
  const float bdb = bd / b;
  const float sq = use_sde ? sqrtf(__fmul_rn(2.f, eps)) : 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float e = __ldg(eta + i), xv = __ldg(x + i);
    float d = __fadd_rn(__fmul_rn(ad, e), __fmul_rn(bdb, __fsub_rn(xv, __fmul_rn(a, e))));
    if (use_sde) {
      const float score = -e / a;
      d = __fadd_rn(d, __fadd_rn(__fmul_rn(eps, score), __fmul_rn(__ldg(noise + i), sq)));
    }
    out[i] = d;
  }
}

// ------------------------------------------------------------------ decode (model_train_inference.py:373-404)
// One thread per voxel; fp32 op ORDER fixed (no FMA contraction): sequential sum of squares,
// sqrt, clamp 1e-12, divide, ncat sequential dot products, first-max argmax -> int64.
__device__ __forceinline__ int decode_voxel(const float* __restrict__ xp, size_t n, const float* s_en, int E, int ncat) {
  float xv[32];
  float ss = 0.f;
  for (int e = 0; e < E; ++e) {
    xv[e] = __ldg(xp + (size_t)e * n);
    ss = __fadd_rn(ss, __fmul_rn(xv[e], xv[e]));
  }
  const float nrm = fmaxf(__fsqrt_rn(ss), 1e-12f);
  for (int e = 0; e < E; ++e) xv[e] = __fdiv_rn(xv[e], nrm);
  float best = -INFINITY;
  int arg = 0;
  for (int c = 0; c < ncat; ++c) {
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = __fadd_rn(acc, __fmul_rn(xv[e], s_en[c * E + e]));
    // strictly-greater keeps the FIRST maximum; a NaN logit wins like torch.argmax
    if (c == 0 || acc > best || (acc != acc && best == best)) { best = acc; arg = c; }
  }
  return arg;
}

// decode(x, return_logits=True) (:398-399): the cosine logits [B, ncat, n] themselves, same fp32 op order
__global__ void decode_logits_kernel(const float* __restrict__ x, const float* __restrict__ en,
                                     float* __restrict__ logits, int B, int E, int ncat, size_t n) {
  extern __shared__ float s_en[];
  for (int i = threadIdx.x; i < ncat * E; i += blockDim.x) s_en[i] = en[i];
  __syncthreads();
  const size_t total = (size_t)B * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / n);
    const size_t v = i % n;
    const float* xp = x + (size_t)b * E * n + v;
    float xv[32];
    float ss = 0.f;
    for (int e = 0; e < E; ++e) {
      xv[e] = __ldg(xp + (size_t)e * n);
      ss = __fadd_rn(ss, __fmul_rn(xv[e], xv[e]));
    }
    const float nrm = fmaxf(__fsqrt_rn(ss), 1e-12f);
    for (int e = 0; e < E; ++e) xv[e] = __fdiv_rn(xv[e], nrm);
    for (int c = 0; c < ncat; ++c) {
      float acc = 0.f;
      for (int e = 0; e < E; ++e) acc = __fadd_rn(acc, __fmul_rn(xv[e], s_en[c * E + e]));
      logits[((size_t)b * ncat + c) * n + v] = acc;
    }
  }
}

__global__ void decode_kernel(const float* __restrict__ x, const float* __restrict__ en,
                              long long* __restrict__ out, int B, int E, int ncat, size_t n) {
  extern __shared__ float s_en[];
  for (int i = threadIdx.x; i < ncat * E; i += blockDim.x) s_en[i] = en[i];
  __syncthreads();
  const size_t total = (size_t)B * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / n);
    const size_t v = i % n;
    out[i] = decode_voxel(x + (size_t)b * E * n + v, n, s_en, E, ncat);
  }
}

// The same arithmetic, 4 consecutive voxels per thread and compile-time (E, NCAT): 16-byte loads of x, the embedding
// rows read from shared memory as LDS.128 with immediate offsets once per 4 voxels (the generic kernel spends one LDS
// per multiply), 16-byte stores of the int64 result.  Every voxel still sees exactly the op sequence of decode_voxel
// (sequential unfused sums, IEEE sqrt / divide, first maximum), so the output is bit-identical.
template <int E, int NCAT>
__device__ __forceinline__ void decode_voxel4(const float* __restrict__ xp, size_t n, const float* s_en, int (&arg)[4]) {
  float xv[4][E];
  float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(xp + (size_t)e * n));
    xv[0][e] = q.x; xv[1][e] = q.y; xv[2][e] = q.z; xv[3][e] = q.w;
#pragma unroll
    for (int v = 0; v < 4; ++v) ss[v] = __fadd_rn(ss[v], __fmul_rn(xv[v][e], xv[v][e]));
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const float nrm = fmaxf(__fsqrt_rn(ss[v]), 1e-12f);
#pragma unroll
    for (int e = 0; e < E; ++e) xv[v][e] = __fdiv_rn(xv[v][e], nrm);
  }
  float best[4];
#pragma unroll
  for (int c = 0; c < NCAT; ++c) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const float w = s_en[c * E + e];
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[v] = __fadd_rn(acc[v], __fmul_rn(xv[v][e], w));
    }
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (c == 0 || acc[v] > best[v] || (acc[v] != acc[v] && best[v] == best[v])) { best[v] = acc[v]; arg[v] = c; }
  }
}

template <int E, int NCAT>
__global__ void __launch_bounds__(128)
decode4_kernel(const float* __restrict__ x, const float* __restrict__ en, long long* __restrict__ out, int B, size_t n) {
  __shared__ __align__(16) float s_en[NCAT * E];
  for (int i = threadIdx.x; i < NCAT * E; i += blockDim.x) s_en[i] = en[i];
  __syncthreads();
  const size_t n4 = n >> 2, total = (size_t)B * n4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / n4, v = (i % n4) << 2;
    int arg[4];
    decode_voxel4<E, NCAT>(x + b * E * n + v, n, s_en, arg);
    longlong2* o = reinterpret_cast<longlong2*>(out + b * n + v);
    o[0] = make_longlong2(arg[0], arg[1]);
    o[1] = make_longlong2(arg[2], arg[3]);
  }
}

template <int E, int NCAT>
__global__ void __launch_bounds__(128)
decode_vote4_kernel(const float* __restrict__ x, const float* __restrict__ en, int S, size_t n,
                    long long* __restrict__ decoded, int* __restrict__ counts) {
  __shared__ __align__(16) float s_en[NCAT * E];
  __shared__ int s_cnt[NCAT][4][128];          // histogram of this thread's 4 voxels
  for (int i = threadIdx.x; i < NCAT * E; i += blockDim.x) s_en[i] = en[i];
  __syncthreads();
  const size_t n4 = n >> 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i << 2;
#pragma unroll
    for (int c = 0; c < NCAT; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) s_cnt[c][k][threadIdx.x] = 0;
    for (int s = 0; s < S; ++s) {
      int arg[4];
      decode_voxel4<E, NCAT>(x + (size_t)s * E * n + v, n, s_en, arg);
      if (decoded) {
        longlong2* o = reinterpret_cast<longlong2*>(decoded + (size_t)s * n + v);
        o[0] = make_longlong2(arg[0], arg[1]);
        o[1] = make_longlong2(arg[2], arg[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s_cnt[arg[k]][k][threadIdx.x] += 1;
    }
#pragma unroll
    for (int c = 0; c < NCAT; ++c) {
      int4* cp = reinterpret_cast<int4*>(counts + (size_t)c * n + v);
      int4 old = *cp;
      old.x += s_cnt[c][0][threadIdx.x]; old.y += s_cnt[c][1][threadIdx.x];
      old.z += s_cnt[c][2][threadIdx.x]; old.w += s_cnt[c][3][threadIdx.x];
      *cp = old;
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Ensemble vote (model_inference_experiments.py:442-447): decode S samples of one voxel and add them to the per-voxel
// category histogram counts[ncat][n] (int32, +=, so several launches / ranks accumulate).  Thread = voxel; the
// histogram of the voxel lives in shared memory ([ncat][128] ints) while the samples stream by, so the decoded
// volumes never have to reach HBM (`decoded` is optional).
__global__ void __launch_bounds__(128)
decode_vote_kernel(const float* __restrict__ x, const float* __restrict__ en, int S, int E, int ncat, size_t n,
                   long long* __restrict__ decoded, int* __restrict__ counts) {
  extern __shared__ float s_en[];
  int* s_cnt = reinterpret_cast<int*>(s_en + ncat * E);   // [ncat][128]
  for (int i = threadIdx.x; i < ncat * E; i += blockDim.x) s_en[i] = en[i];
  __syncthreads();
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (size_t)gridDim.x * blockDim.x) {
    for (int c = 0; c < ncat; ++c) s_cnt[c * 128 + threadIdx.x] = 0;
    for (int s = 0; s < S; ++s) {
      const int arg = decode_voxel(x + (size_t)s * E * n + v, n, s_en, E, ncat);
      if (decoded) decoded[(size_t)s * n + v] = arg;
      s_cnt[arg * 128 + threadIdx.x] += 1;
    }
    for (int c = 0; c < ncat; ++c) counts[(size_t)c * n + v] += s_cnt[c * 128 + threadIdx.x];
  }
}

__global__ void embed_kernel(const long long* __restrict__ cats, const float* __restrict__ w,
                             float* __restrict__ out, int B, int E, int ncat, size_t n, int shift) {
  const size_t total = (size_t)B * E * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % n;
    const int e = (int)((i / n) % E);
    const int b = (int)(i / (n * E));
    long long c = cats[(size_t)b * n + v] + shift;  // embed(): indices = x + 1 (:366)
    c = c < 0 ? 0 : (c >= ncat ? ncat - 1 : c);
    out[i] = __ldg(w + c * E + e);
  }
}

// EMACallback.on_train_batch_end (project/geodata-3d-conditional/callbacks.py:263-266)
__global__ void ema_kernel(float* __restrict__ shadow, const float* __restrict__ param, size_t n,
                           float decay, float omd) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    shadow[i] = __fadd_rn(__fmul_rn(decay, shadow[i]), __fmul_rn(omd, __ldg(param + i)));
}

// training_step loss partials (model_train_inference.py:443): sum (v-vhat)^2 and sum v^2
__global__ void mse_kernel(const float* __restrict__ v, const float* __restrict__ vh, size_t n,
                           double* __restrict__ acc2) {
  double s0 = 0.0, s1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const float a = __ldg(v + i), d = a - __ldg(vh + i);
    s0 += (double)d * d;
    s1 += (double)a * a;
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(acc2, s0);
    atomicAdd(acc2 + 1, s1);
  }
}

}  // namespace

int pack_ncdhw_to_blocked(const float* x, int B, int C, int D, int H, int W, Act& out, cudaStream_t st) {
  FTB_CHECK(out.B == B && out.D == D && out.H == H && out.W == W && out.C >= C && out.C % 16 == 0,
            "pack: output activation shape");
  const size_t vox = (size_t)D * H * W;
  pack_kernel<<<grid_for((size_t)B * out.cg() * vox, 256), 256, 0, st>>>(x, B, C, vox, out.cg(), out.p);
  FTB_LAUNCH_OK();
  return 0;
}
int pack_unfold_w(const float* x, int B, int C, int D, int H, int W, int K, Act& out, cudaStream_t st) {
  FTB_CHECK(out.B == B && out.D == D && out.H == H && out.W == W && out.C >= K * C && out.C % 16 == 0,
            "pack_unfold_w: output activation shape");
  FTB_CHECK(out.cg() * 8 <= 256 && C <= 255, "pack_unfold_w: at most 256 unfolded channels");
  const size_t smem = (size_t)kUnfRows * C * (W + K - 1) * sizeof(float);
  FTB_CHECK(smem <= 44 * 1024, "pack_unfold_w: row too wide for shared memory");
  const int lanes = 256 / out.cg() >= 1 ? 256 / out.cg() : 1;
  FTB_CHECK(out.cg() <= 256, "pack_unfold_w: too many channel groups");
  pack_unfold_w_kernel<<<(unsigned)((size_t)B * D * cdiv(H, kUnfRows)), out.cg() * lanes, smem, st>>>(x, C, D, H, W, K, out.cg(), out.p);
  FTB_LAUNCH_OK();
  return 0;
}
int unpack_blocked_to_ncdhw(const Act& in, int cgoff, int C, float* out, cudaStream_t st) {
  const size_t vox = in.voxels();
  unpack_kernel<<<grid_for((size_t)in.B * ((C + 7) / 8) * vox, 256), 256, 0, st>>>(in.p, in.B, in.cg(), cgoff, C, vox, out);
  FTB_LAUNCH_OK();
  return 0;
}
int trilinear_resample(const Act& in, Act& out, cudaStream_t st) {
  FTB_CHECK(in.B == out.B && in.C == out.C, "trilinear: batch/channels must match");
  const int hgroups = cdiv(out.H, kTriRH);
  const long long blocks = (long long)out.B * out.cg() * out.D * hgroups;
  FTB_CHECK(blocks < (1ll << 31), "trilinear: grid too large");
  const int threads = kTriRH * out.W >= 256 ? 256 : round_up(kTriRH * out.W, 32);
  // ATen's align_corners scale, computed once in fp32 exactly like area_pixel_compute_scale
  auto scale = [](int i, int o) { return o > 1 ? (float)(i - 1) / (float)(o - 1) : 0.f; };
  // upsampling (every input voxel feeds ~8 outputs): stage the source rows of a row group in shared memory
  if (out.H > in.H && out.W >= 16 && getenv("FTB_TRILINEAR_DIRECT") == nullptr) {
    const float shf = scale(in.H, out.H);
    const int max_rows = (int)ceilf(shf * (kTriRH - 1)) + 3;
    const size_t smem = (size_t)max_rows * in.W * 2 * sizeof(float4);   // D-blended fp32 rows
    if (smem <= 48 * 1024) {
      trilinear_smem_kernel<<<(unsigned)blocks, 256, smem, st>>>(in.p, in.D, in.H, in.W, out.D, out.H, out.W, hgroups,
                                                                kTriRH, max_rows, scale(in.D, out.D), shf,
                                                                scale(in.W, out.W), out.p);
      FTB_LAUNCH_OK();
      return 0;
    }
  }
  trilinear_kernel<<<(unsigned)blocks, threads, 0, st>>>(in.p, in.cg(), in.D, in.H, in.W, out.D, out.H, out.W,
                                                         hgroups, scale(in.D, out.D), scale(in.H, out.H),
                                                         scale(in.W, out.W), out.p);
  FTB_LAUNCH_OK();
  return 0;
}

int film_concat(const Act& x, const Act& atb, const float* film, int film_stride, int c_real, Act& out,
                cudaStream_t st) {
  FTB_CHECK(x.C == atb.C && out.C == 2 * x.C && out.B == x.B, "film_concat: channel/batch mismatch");
  FTB_CHECK(atb.B == x.B || atb.B == 1, "film_concat: ATb batch must be 1 or B");
  FTB_CHECK(x.voxels() == atb.voxels() && x.voxels() == out.voxels(), "film_concat: spatial mismatch");
  const size_t total = (size_t)x.B * 2 * x.cg() * x.voxels();
  const size_t a_bstride = atb.B == 1 ? 0 : (size_t)atb.C * atb.voxels();
  film_concat_kernel<<<grid_for(total, 256), 256, 0, st>>>(x.p, atb.p, x.B, x.cg(), x.voxels(), a_bstride, film,
                                                           film_stride, c_real, out.p);
  FTB_LAUNCH_OK();
  return 0;
}
int interp_xt_bt(int kind, int one_sided, float gamma_a, const float* x0, const float* x1,
                 const float* z, const float* t, float* xt, float* bt, int B, long long n,
                 cudaStream_t st) {
  FTB_CHECK(kind >= 0 && kind <= 4, "interp: unknown interpolant kind");
  FTB_CHECK(n % 4 == 0, "interp: per-sample element count must be a multiple of 4");
  const size_t n4 = (size_t)n / 4;
  interp_kernel<<<grid_for((size_t)B * n4, 256), 256, 0, st>>>(
      kind, one_sided, gamma_a, (const float4*)x0, (const float4*)x1, (const float4*)z, t,
      (float4*)xt, (float4*)bt, B, n4);
  FTB_LAUNCH_OK();
  return 0;
}
int axpy_out(float* out, const float* x, const float* k, float h, long long n,
             const unsigned char* frozen, long long inner, cudaStream_t st) {
  if (!frozen && n % 4 == 0) {
    axpy4_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>((float4*)out, (const float4*)x, (const float4*)k, h, (size_t)n / 4);
  } else {
    axpy_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(out, x, k, h, (size_t)n, frozen, (size_t)(inner > 0 ? inner : 1));
  }
  FTB_LAUNCH_OK();
  return 0;
}
int heun_combine(float* out, const float* x, const float* k1, const float* k2, double h, long long n, cudaStream_t st) {
  FTB_CHECK(n % 4 == 0, "heun: n must be a multiple of 4");
  heun_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>((float4*)out, (const float4*)x, (const float4*)k1, (const float4*)k2, (float)(h / 2.0), (size_t)n / 4);
  FTB_LAUNCH_OK();
  return 0;
}
int rk4_combine(float* out, const float* x, const float* k1, const float* k2, const float* k3,
                const float* k4, double h, long long n, cudaStream_t st) {
  FTB_CHECK(n % 4 == 0, "rk4: n must be a multiple of 4");
  rk4_kernel<<<grid_for((size_t)n / 4, 256), 256, 0, st>>>((float4*)out, (const float4*)x, (const float4*)k1, (const float4*)k2, (const float4*)k3, (const float4*)k4, (float)(h / 6.0), (size_t)n / 4);
  FTB_LAUNCH_OK();
  return 0;
}
int denoise_drift(float* out, const float* x, const float* eta, const float* noise, float a, float b,
                  float ad, float bd, float eps, int use_sde, long long n, cudaStream_t st) {
  FTB_CHECK(!use_sde || noise != nullptr, "drift: SDE term needs a noise tensor");
  drift_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(out, x, eta, noise, a, b, ad, bd, eps, use_sde, (size_t)n);
  FTB_LAUNCH_OK();
  return 0;
}
int denoise_drift_dev(float* out, const float* x, const float* eta, const float* noise, const float* coef, int use_sde,
                      long long n, cudaStream_t st) {
  FTB_CHECK(!use_sde || noise != nullptr, "drift: SDE term needs a noise tensor");
  drift_dev_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(out, x, eta, noise, coef, use_sde, (size_t)n);
  FTB_LAUNCH_OK();
  return 0;
}
int decode_argmax(const float* x, const float* en, long long* out, int B, int E, int ncat,
                  long long n, cudaStream_t st) {
  FTB_CHECK(E >= 1 && E <= 32, "decode: embedding dim must be in [1,32]");
  FTB_CHECK(ncat >= 1 && ncat <= 256, "decode: category count");
  // the shipped shapes (15 categories in an 18- / 15- / 20-d embedding) run 4 voxels per thread; anything else, ragged
  // or unaligned volumes take the generic one-voxel kernel (same arithmetic, bit-identical output)
  if (ncat == 15 && (n & 3) == 0 && aligned16(x) && aligned16(out) && (E == 18 || E == 15 || E == 20)) {
    const int g = grid_for((size_t)B * (n >> 2), 128);
    if (E == 18) decode4_kernel<18, 15><<<g, 128, 0, st>>>(x, en, out, B, (size_t)n);
    else if (E == 15) decode4_kernel<15, 15><<<g, 128, 0, st>>>(x, en, out, B, (size_t)n);
    else decode4_kernel<20, 15><<<g, 128, 0, st>>>(x, en, out, B, (size_t)n);
    FTB_LAUNCH_OK();
    return 0;
  }
  decode_kernel<<<grid_for((size_t)B * n, 128), 128, (size_t)ncat * E * sizeof(float), st>>>(x, en, out, B, E, ncat, (size_t)n);
  FTB_LAUNCH_OK();
  return 0;
}
int decode_logits(const float* x, const float* en, float* logits, int B, int E, int ncat, long long n, cudaStream_t st) {
  FTB_CHECK(E >= 1 && E <= 32, "decode: embedding dim must be in [1,32]");
  FTB_CHECK(ncat >= 1 && ncat <= 256, "decode: category count");
  decode_logits_kernel<<<grid_for((size_t)B * n, 128), 128, (size_t)ncat * E * sizeof(float), st>>>(x, en, logits, B, E, ncat,
                                                                                                  (size_t)n);
  FTB_LAUNCH_OK();
  return 0;
}
int decode_vote(const float* x, const float* en, int S, int E, int ncat, long long n, long long* decoded, int* counts,
                cudaStream_t st) {
  FTB_CHECK(E >= 1 && E <= 32, "decode: embedding dim must be in [1,32]");
  FTB_CHECK(ncat >= 1 && ncat <= 64, "decode_vote: at most 64 categories");
  if (ncat == 15 && (n & 3) == 0 && aligned16(x) && aligned16(counts) && (!decoded || aligned16(decoded)) &&
      (E == 18 || E == 15 || E == 20)) {
    const int g = grid_for((size_t)(n >> 2), 128);
    if (E == 18) decode_vote4_kernel<18, 15><<<g, 128, 0, st>>>(x, en, S, (size_t)n, decoded, counts);
    else if (E == 15) decode_vote4_kernel<15, 15><<<g, 128, 0, st>>>(x, en, S, (size_t)n, decoded, counts);
    else decode_vote4_kernel<20, 15><<<g, 128, 0, st>>>(x, en, S, (size_t)n, decoded, counts);
    FTB_LAUNCH_OK();
    return 0;
  }
  const size_t smem = (size_t)ncat * E * sizeof(float) + (size_t)ncat * 128 * sizeof(int);
  decode_vote_kernel<<<grid_for((size_t)n, 128), 128, smem, st>>>(x, en, S, E, ncat, (size_t)n, decoded, counts);
  FTB_LAUNCH_OK();
  return 0;
}
int embed_lookup(const long long* cats, const float* w, float* out, int B, int E, int ncat,
                 long long n, int shift, cudaStream_t st) {
  embed_kernel<<<grid_for((size_t)B * E * n, 256), 256, 0, st>>>(cats, w, out, B, E, ncat, (size_t)n, shift);
  FTB_LAUNCH_OK();
  return 0;
}
int ema_update(float* shadow, const float* param, long long n, double decay, cudaStream_t st) {
  ema_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(shadow, param, (size_t)n, (float)decay, (float)(1.0 - decay));
  FTB_LAUNCH_OK();
  return 0;
}
int mse_ratio_partial(const float* v, const float* vhat, long long n, double* acc2, cudaStream_t st) {
  mse_kernel<<<grid_for((size_t)n, 256), 256, 0, st>>>(v, vhat, (size_t)n, acc2);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
