// Training-step kernels around the tensor-core convs: the standalone RMSNorm / FiLM / SiLU / residual
// pass and its backward, bias gradients, the adjoint of the trilinear resample, the softmax backward
// passes of LinearAttention / Attention, the time-path backward, and the fused clip + Adam update.
// Reference: Block.forward (src/flowtrain/models/unet_attn_3d.py:232-244), RMSNorm (:111-128),
// LinearAttention (:308-341), Attention (:357-373, :436-465), time MLPs (:203-218, :551-556,
// :255-257), training_step / configure_optimizers
// (project/geodata-3d-unconditional/model_train_inference.py:417-473).
// All of these are HBM-bound (or tiny): coalesced 16-byte accesses on the blocked bf16 layout,
// warp-shuffle + shared-memory block reductions, fp32 atomics for the per-channel sums.
#include "ops.h"

namespace ftb {

namespace {

__device__ __forceinline__ float sigm(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }

struct NormActP {
  const bf16* u;
  bf16* out;
  const bf16* resid;
  const bf16* dout;
  bf16* du;
  int CG;
  size_t vox;
  int norm, silu;
  const float* gain;            // [C] shared factor (g*sqrt(C)) or null
  const float *s1, *sh;         // per-sample factor / offset rows (FiLM scale+1, shift) or null
  int fstride;
  float *R, *S, *dbias;         // backward sums: R[b][c] = sum dz*n, S[b][c] = sum dz, dbias[c] = sum du
  int sstride;                  // row stride of S
  float drop_p;                 // nn.Dropout after the activation (Block.forward :244), 0 = off
  unsigned long long drop_key;  // per-(step, block) key of the counter-based mask
};

__device__ __forceinline__ DropMask drop_mask8(const NormActP& p, size_t group) {
  return ::ftb::drop_mask8(p.drop_p, p.drop_key, group);
}

// Thread mapping of both passes: LPV lanes per voxel (8, 16 or 32 >= C/8), lane = one 8-channel group (one 16-byte
// unit of the blocked layout).  A thread therefore holds its 8 channels of u / dY / dz in registers for the whole
// computation (every tensor is read ONCE), the channel reductions of the RMSNorm (sum of squares, n.dn) are
// log2(LPV) xor-shuffles, and the per-channel sums over voxels needed by the FiLM / gain gradients accumulate in 8
// registers per thread with no cross-lane traffic until the block ends.  Lanes of a warp with the same channel
// group read consecutive voxels (64-byte runs at LPV = 8).
constexpr int kNaUnr = 4;   // voxels in flight per thread

template <int LPV>
__device__ __forceinline__ float lanes_sum(float v) {
#pragma unroll
  for (int o = 1; o < LPV; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out = act(n * gain * s1 + sh) * dropout + resid,  n = u / max(||u||_C, 1e-12) (norm) or u
template <int LPV, bool kNorm, bool kSilu, bool kDrop>
__global__ void __launch_bounds__(256)
normact_fwd_kernel(const NormActP p, int iters) {
  constexpr int VPB = 256 / LPV;
  const int b = blockIdx.y;
  const int cgi = threadIdx.x % LPV, vi = threadIdx.x / LPV;
  const bool act = cgi < p.CG;
  float m[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cgi * 8 + j;
    m[j] = 1.f; sh[j] = 0.f;
    if (act) {
      if (p.gain) m[j] = __ldg(p.gain + c);
      if (p.s1) { m[j] *= __ldg(p.s1 + (size_t)b * p.fstride + c); sh[j] = __ldg(p.sh + (size_t)b * p.fstride + c); }
    }
  }
  const size_t cbase = ((size_t)b * p.CG + (act ? cgi : 0)) * p.vox * 8;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int it0 = 0; it0 < iters; it0 += kNaUnr) {
    // all loads of kNaUnr voxels first: enough bytes in flight per SM to cover the HBM latency
    uint4 Uq[kNaUnr], Rq[kNaUnr];
#pragma unroll
    for (int k = 0; k < kNaUnr; ++k) {
      const size_t v = ((size_t)blockIdx.x * iters + it0 + k) * VPB + vi;
      const bool in = act && it0 + k < iters && v < p.vox;
      const size_t off = cbase + (in ? v : 0) * 8;
      Uq[k] = in ? __ldg(reinterpret_cast<const uint4*>(p.u + off)) : zero4;
      Rq[k] = (in && p.resid) ? __ldg(reinterpret_cast<const uint4*>(p.resid + off)) : zero4;
    }
#pragma unroll
    for (int k = 0; k < kNaUnr; ++k) {
    const size_t v = ((size_t)blockIdx.x * iters + it0 + k) * VPB + vi;
    const bool in = act && it0 + k < iters && v < p.vox;
    const size_t off = cbase + (in ? v : 0) * 8;
    float f[8], r[8];
    unpack_bf16x8(Uq[k], f);
    unpack_bf16x8(Rq[k], r);
    float rinv = 1.f;
    if (kNorm) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
      rinv = 1.f / fmaxf(sqrtf(lanes_sum<LPV>(ss)), 1e-12f);
    }
    DropMask dm;
    if (kDrop) dm = drop_mask8(p, off >> 3);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(f[j] * rinv, m[j], sh[j]);
      if (kSilu) y = y * sigm(y);
      if (kDrop) y *= dm(j);
      f[j] = y + r[j];   // r is zero without a residual
    }
    if (in) *reinterpret_cast<uint4*>(p.out + off) = pack_bf16x8(f);
    }
  }
}

// backward of the above: du (may alias dout), R[b][c] += sum_v dz*n, S += sum_v dz, dbias[c] += sum du
template <int LPV, bool kNorm, bool kSilu, bool kDrop>
__global__ void __launch_bounds__(256)
normact_bwd_kernel(const NormActP p, int iters) {
  constexpr int VPB = 256 / LPV;
  __shared__ float s_red[3 * 256];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int cgi = threadIdx.x % LPV, vi = threadIdx.x / LPV;
  const bool act = cgi < p.CG;
  for (int i = threadIdx.x; i < 3 * 256; i += 256) s_red[i] = 0.f;
  __syncthreads();
  float m[8], sh[8], accR[8], accS[8], accB[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cgi * 8 + j;
    m[j] = 1.f; sh[j] = 0.f; accR[j] = 0.f; accS[j] = 0.f; accB[j] = 0.f;
    if (act) {
      if (p.gain) m[j] = __ldg(p.gain + c);
      if (p.s1) { m[j] *= __ldg(p.s1 + (size_t)b * p.fstride + c); sh[j] = __ldg(p.sh + (size_t)b * p.fstride + c); }
    }
  }
  const size_t cbase = ((size_t)b * p.CG + (act ? cgi : 0)) * p.vox * 8;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int it0 = 0; it0 < iters; it0 += kNaUnr) {
    uint4 Uq[kNaUnr], Gq[kNaUnr];
#pragma unroll
    for (int k = 0; k < kNaUnr; ++k) {
      const size_t v = ((size_t)blockIdx.x * iters + it0 + k) * VPB + vi;
      const bool in = act && it0 + k < iters && v < p.vox;
      const size_t off = cbase + (in ? v : 0) * 8;
      Uq[k] = in ? __ldg(reinterpret_cast<const uint4*>(p.u + off)) : zero4;
      Gq[k] = in ? *reinterpret_cast<const uint4*>(p.dout + off) : zero4;
    }
#pragma unroll
    for (int k = 0; k < kNaUnr; ++k) {
    const size_t v = ((size_t)blockIdx.x * iters + it0 + k) * VPB + vi;
    const bool in = act && it0 + k < iters && v < p.vox;
    const size_t off = cbase + (in ? v : 0) * 8;
    float f[8], g[8];
    unpack_bf16x8(Uq[k], f);
    unpack_bf16x8(Gq[k], g);
    float rinv = 1.f;
    if (kNorm) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
      rinv = 1.f / fmaxf(sqrtf(lanes_sum<LPV>(ss)), 1e-12f);
    }
    DropMask dm;
    if (kDrop) dm = drop_mask8(p, off >> 3);
    float dotp = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float n = f[j] * rinv;
      float dz = g[j];
      if (kDrop) dz *= dm(j);
      if (kSilu) {
        const float z = fmaf(n, m[j], sh[j]);
        const float sg = sigm(z);
        dz *= sg * (1.f + z * (1.f - sg));
      }
      accR[j] = fmaf(dz, n, accR[j]);
      accS[j] += dz;
      f[j] = n;
      g[j] = dz * m[j];          // dn
      dotp = fmaf(n, g[j], dotp);
    }
    if (kNorm) {
      const float dot = lanes_sum<LPV>(dotp);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = rinv * (g[j] - f[j] * dot);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) accB[j] += g[j];
    if (in) *reinterpret_cast<uint4*>(p.du + off) = pack_bf16x8(g);
    }
  }
  // lanes of the warp that share a channel group: combine, then one shared-memory add per warp and channel
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = LPV; o < 32; o <<= 1) {
      accR[j] += __shfl_xor_sync(0xffffffffu, accR[j], o);
      accS[j] += __shfl_xor_sync(0xffffffffu, accS[j], o);
      accB[j] += __shfl_xor_sync(0xffffffffu, accB[j], o);
    }
  }
  if (lane < LPV && act) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&s_red[cgi * 8 + j], accR[j]);
      atomicAdd(&s_red[256 + cgi * 8 + j], accS[j]);
      atomicAdd(&s_red[512 + cgi * 8 + j], accB[j]);
    }
  }
  __syncthreads();
  const int C = p.CG * 8;
  for (int c = threadIdx.x; c < C; c += 256) {
    if (p.R) atomicAdd(p.R + (size_t)b * C + c, s_red[c]);
    if (p.S) atomicAdd(p.S + (size_t)b * p.sstride + c, s_red[256 + c]);
    if (p.dbias) atomicAdd(p.dbias + c, s_red[512 + c]);
  }
}

// ds1[b][c] = gain[c] * R[b][c];  dg[c] += sqrt_c * sum_b s1[b][c] * R[b][c]   (gain = g*sqrt_c)
__global__ void normact_finish_kernel(const float* __restrict__ R, int B, int C, const float* __restrict__ gain,
                                      const float* __restrict__ s1, int fstride, float sqrt_c,
                                      float* __restrict__ ds1, float* __restrict__ dg) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int b = 0; b < B; ++b) {
    const float r = R[(size_t)b * C + c];
    if (ds1) ds1[(size_t)b * fstride + c] = (gain ? gain[c] : 1.f) * r;
    a += (s1 ? s1[(size_t)b * fstride + c] : 1.f) * r;
  }
  if (dg) atomicAdd(dg + c, a * sqrt_c);
}

// db[c] += sum over (b, voxels) of dy[b][c][v]; grid (blocks, CG, B)
__global__ void __launch_bounds__(256)
chan_sum_kernel(const bf16* __restrict__ dy, int cgtot, int cgoff, size_t vox, int C, float* __restrict__ db) {
  __shared__ float red[8][8];
  const int cg = blockIdx.y, b = blockIdx.z;
  const bf16* base = dy + ((size_t)b * cgtot + cgoff + cg) * vox * 8;
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = 0.f;
  for (size_t v = (size_t)blockIdx.x * 256 + threadIdx.x; v < vox; v += (size_t)gridDim.x * 256) {
    float f[8];
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(base + v * 8)), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += f[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = warp_sum(a[j]);
  if ((threadIdx.x & 31) == 0)
    for (int j = 0; j < 8; ++j) red[threadIdx.x >> 5][j] = a[j];
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int c = cg * 8 + threadIdx.x;
    if (c < C) atomicAdd(db + c, s);
  }
}

// dst (+)= src  (blocked bf16, same shape)
__global__ void accum_kernel(bf16* __restrict__ dst, const bf16* __restrict__ src, size_t n8, int add) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float s[8];
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(src) + i), s);
    if (add) {
      float d[8];
      unpack_bf16x8(reinterpret_cast<const uint4*>(dst)[i], d);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += d[j];
    }
    reinterpret_cast<uint4*>(dst)[i] = pack_bf16x8(s);
  }
}

// ---------------------------------------------------------------- adjoint of the trilinear resample
// Forward (elementwise.cu trilinear_kernel): out[o] = sum_corners w * in[i], index rule as ATen.
// Adjoint as a gather: input voxel i collects w(o, i) * dout[o] over the few outputs o whose two
// source indices include i; the candidate range comes from inverting src = scale * o.
struct AxisRange { int lo, hi; };
__device__ __forceinline__ AxisRange cand(int i, int in, int out, float scale) {
  AxisRange r;
  if (scale <= 0.f) { r.lo = 0; r.hi = out - 1; return r; }
  r.lo = max(0, (int)floorf((float)(i - 1) / scale) - 1);
  r.hi = min(out - 1, (int)ceilf((float)(i + 1) / scale) + 1);
  return r;
}
__device__ __forceinline__ float axis_w(int o, int i, int in, float scale) {
  const float src = scale * (float)o;
  const int i0 = (int)src;
  const int i1 = i0 + (i0 < in - 1 ? 1 : 0);
  const float l1 = src - (float)i0;
  float w = 0.f;
  if (i0 == i) w += 1.f - l1;
  if (i1 == i) w += l1;
  return w;
}
// Per-axis list of the outputs that read input index i, with their weights: cnt <= kTriMaxTap entries, or -1 when the
// axis has more contributors than that (extreme up-scales: the caller falls back to the range scan).
constexpr int kTriMaxTap = 6;
struct AxisTaps {
  int idx[kTriMaxTap];
  float w[kTriMaxTap];
  int cnt;
};
__device__ __forceinline__ AxisTaps axis_taps(int i, int in, int out, float scale) {
  AxisTaps t;
  t.cnt = 0;
  const AxisRange r = cand(i, in, out, scale);
#pragma unroll
  for (int k = 0; k < kTriMaxTap; ++k) { t.idx[k] = 0; t.w[k] = 0.f; }
  for (int o = r.lo; o <= r.hi; ++o) {
    const float w = axis_w(o, i, in, scale);
    if (w == 0.f) continue;
    if (t.cnt == kTriMaxTap) { t.cnt = -1; return t; }
#pragma unroll
    for (int k = 0; k < kTriMaxTap; ++k)
      if (k == t.cnt) { t.idx[k] = o; t.w[k] = w; }
    ++t.cnt;
  }
  return t;
}

// grid (blocks over input voxels, B*CG); din (+)= adjoint(dout).  The contributors of an input voxel form a product
// set (taps along D) x (taps along H) x (taps along W): the three short lists are built once per thread (3 x ~8
// weight evaluations), then the gather runs over exactly the contributing outputs (4 x 4 x 4 for a 2x upsample)
// instead of re-evaluating the weights inside a 7 x 7 x 7 range scan.
__global__ void __launch_bounds__(256)
trilinear_bwd_kernel(const bf16* __restrict__ dout, int Di, int Hi, int Wi, int Do, int Ho, int Wo, float sd, float sh,
                     float sw, bf16* __restrict__ din, int add) {
  const size_t bc = blockIdx.y;
  const size_t vin = (size_t)Di * Hi * Wi, vout = (size_t)Do * Ho * Wo;
  const size_t v = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= vin) return;
  const int w = (int)(v % Wi), h = (int)((v / Wi) % Hi), d = (int)(v / ((size_t)Wi * Hi));
  const bf16* ob = dout + bc * vout * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const AxisTaps td = axis_taps(d, Di, Do, sd), th = axis_taps(h, Hi, Ho, sh), tw = axis_taps(w, Wi, Wo, sw);
  if (td.cnt >= 0 && th.cnt >= 0 && tw.cnt >= 0) {
#pragma unroll
    for (int a = 0; a < kTriMaxTap; ++a) {
      if (a >= td.cnt) break;
#pragma unroll
      for (int b = 0; b < kTriMaxTap; ++b) {
        if (b >= th.cnt) break;
        const float wdh = td.w[a] * th.w[b];
        const bf16* row = ob + (((size_t)td.idx[a] * Ho + th.idx[b]) * Wo) * 8;
#pragma unroll
        for (int c = 0; c < kTriMaxTap; ++c) {
          if (c >= tw.cnt) break;
          const float ww = tw.w[c] * wdh;
          float f[8];
          unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(row + (size_t)tw.idx[c] * 8)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, f[j], acc[j]);
        }
      }
    }
  } else {   // many contributors per axis (large up-scale factors): scan the candidate ranges
    const AxisRange rd = cand(d, Di, Do, sd), rh = cand(h, Hi, Ho, sh), rw = cand(w, Wi, Wo, sw);
    for (int od = rd.lo; od <= rd.hi; ++od) {
      const float wd = axis_w(od, d, Di, sd);
      if (wd == 0.f) continue;
      for (int oh = rh.lo; oh <= rh.hi; ++oh) {
        const float wh = axis_w(oh, h, Hi, sh) * wd;
        if (wh == 0.f) continue;
        for (int ow = rw.lo; ow <= rw.hi; ++ow) {
          const float ww = axis_w(ow, w, Wi, sw) * wh;
          if (ww == 0.f) continue;
          float f[8];
          unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(ob + (((size_t)od * Ho + oh) * Wo + ow) * 8)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(ww, f[j], acc[j]);
        }
      }
    }
  }
  bf16* ip = din + (bc * vin + v) * 8;
  if (add) {
    float f[8];
    unpack_bf16x8(*reinterpret_cast<const uint4*>(ip), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  *reinterpret_cast<uint4*>(ip) = pack_bf16x8(acc);
}

// ---------------------------------------------------------------- weights for the backward convs
// wT[ci][co][K-1-kd][K-1-kh][K-1-kw] = w[co][ci0+ci][kd][kh][kw] * row_scale[ci0+ci]
__global__ void transpose_flip_kernel(const float* __restrict__ w, int cout, int cin, int k3, int ci0, int cin_sub,
                                      const float* __restrict__ row_scale, float* __restrict__ wt) {
  const size_t total = (size_t)cin_sub * cout * k3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % k3);
    const int co = (int)((i / k3) % cout);
    const int ci = (int)(i / ((size_t)k3 * cout));
    const float s = row_scale ? row_scale[ci0 + ci] : 1.f;
    wt[i] = w[((size_t)co * cin + ci0 + ci) * k3 + (k3 - 1 - t)] * s;
  }
}

// to_qkv with the pre-norm gain folded in (W' = W diag(gs)): dW = dW' diag(gs), dg[ci] += sqrt_c sum_co dW'[co][ci] W[co][ci]
__global__ void __launch_bounds__(128)
fold_gain_bwd_kernel(const float* __restrict__ dwp, const float* __restrict__ w, const float* __restrict__ gs,
                     int cout, int cin, float sqrt_c, float* __restrict__ dw, float* __restrict__ dg) {
  __shared__ float red[4];
  const int ci = blockIdx.x;
  float a = 0.f;
  const float g = gs[ci];
  for (int co = threadIdx.x; co < cout; co += 128) {
    const float d = dwp[(size_t)co * cin + ci];
    dw[(size_t)co * cin + ci] += d * g;
    a = fmaf(d, w[(size_t)co * cin + ci], a);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) dg[ci] += (red[0] + red[1] + red[2] + red[3]) * sqrt_c;
}

// ---------------------------------------------------------------- LinearAttention (train path)
// out[b][(h,i)][(h',j)] = (h == h') ? (transpose ? src[b][h][j][i] : src[b][h][i][j]) * mul : 0
__global__ void blockdiag_kernel(const float* __restrict__ src, int heads, int dh, int transpose, float mul,
                                 float* __restrict__ out) {
  const int hd = heads * dh;
  const int b = blockIdx.y;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < hd * hd; o += gridDim.x * blockDim.x) {
    const int r = o / hd, c = o % hd;
    const int h = r / dh, i = r % dh, h2 = c / dh, j = c % dh;
    float v = 0.f;
    if (h == h2) v = src[(((size_t)b * heads + h) * dh + (transpose ? j : i)) * dh + (transpose ? i : j)] * mul;
    out[(size_t)b * hd * hd + o] = v;
  }
}
// ctx diagonal blocks of the per-sample [hd][hd] gradient; also Ssum[b][h][d] = sum_e dctx*ctx
__global__ void dctx_extract_kernel(const float* __restrict__ full, const float* __restrict__ ctx, int heads, int dh,
                                    float* __restrict__ dctx, float* __restrict__ ssum) {
  const int h = blockIdx.x, b = blockIdx.y, hd = heads * dh;
  for (int d = threadIdx.x; d < dh; d += blockDim.x) {
    float s = 0.f;
    for (int e = 0; e < dh; ++e) {
      const float g = full[((size_t)b * hd + h * dh + d) * hd + h * dh + e];
      const size_t o = (((size_t)b * heads + h) * dh + d) * dh + e;
      dctx[o] = g;
      s = fmaf(g, ctx[o], s);
    }
    ssum[((size_t)b * heads + h) * dh + d] = s;
  }
}
// k third of qkv, in place: k~ = exp(k - m[d]) / s[d]   (kstat[b][hd][2] = (m, s) from the combine step)
__global__ void ksoftmax_apply_kernel(bf16* __restrict__ qkv, int cgtot, int kcg0, int kcgs, size_t vox,
                                      const float* __restrict__ kstat, int hd) {
  const int b = blockIdx.z, cg = blockIdx.y;
  bf16* base = qkv + ((size_t)b * cgtot + kcg0 + cg) * vox * 8;
  float m[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m[j] = kstat[((size_t)b * hd + cg * 8 + j) * 2];
    rs[j] = 1.f / kstat[((size_t)b * hd + cg * 8 + j) * 2 + 1];
  }
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += (size_t)gridDim.x * blockDim.x) {
    float f[8];
    unpack_bf16x8(*reinterpret_cast<const uint4*>(base + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __expf(f[j] - m[j]) * rs[j];
    *reinterpret_cast<uint4*>(base + v * 8) = pack_bf16x8(f);
  }
}
// dk third of dqkv, in place: dk = k~ * (dk~ - S[d])
__global__ void ksoftmax_bwd_kernel(bf16* __restrict__ dqkv, const bf16* __restrict__ qkv, int cgtot, int kcg0,
                                    size_t vox, const float* __restrict__ ssum, int hd) {
  const int b = blockIdx.z, cg = blockIdx.y;
  const size_t off = ((size_t)b * cgtot + kcg0 + cg) * vox * 8;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = ssum[(size_t)b * hd + cg * 8 + j];
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < vox; v += (size_t)gridDim.x * blockDim.x) {
    float f[8], k[8];
    unpack_bf16x8(*reinterpret_cast<const uint4*>(dqkv + off + v * 8), f);
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(qkv + off + v * 8)), k);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = k[j] * (f[j] - s[j]);
    *reinterpret_cast<uint4*>(dqkv + off + v * 8) = pack_bf16x8(f);
  }
}
// dq third of dqkv, in place: q~ = softmax_d(q)*scale;  dq = q~ * (dq~ - sum_d q~ dq~ / scale)
// One thread per (voxel, head); the head's dh <= 32 channels of dq~ and q~ stay in registers between the dot product
// and the update, so both tensors are read once (the two-pass version read them twice: 2.7 instead of 1.6 GB per
// 64^3 layer at B=8).
template <int CGH>
__global__ void __launch_bounds__(256)
qsoftmax_bwd_kernel(bf16* __restrict__ dqkv, const bf16* __restrict__ qkv, int cgtot, size_t vox, int heads, float scale) {
  const int b = blockIdx.z, h = blockIdx.y;
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= vox) return;
  float f[CGH][8], q[CGH][8];
  float dot = 0.f;
#pragma unroll
  for (int cg = 0; cg < CGH; ++cg) {
    const size_t off = (((size_t)b * cgtot + h * CGH + cg) * vox + v) * 8;
    unpack_bf16x8(*reinterpret_cast<const uint4*>(dqkv + off), f[cg]);
    unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(qkv + off)), q[cg]);
  }
#pragma unroll
  for (int cg = 0; cg < CGH; ++cg)
#pragma unroll
    for (int j = 0; j < 8; ++j) dot = fmaf(f[cg][j], q[cg][j], dot);
  dot /= scale;
#pragma unroll
  for (int cg = 0; cg < CGH; ++cg) {
    const size_t off = (((size_t)b * cgtot + h * CGH + cg) * vox + v) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) f[cg][j] = q[cg][j] * (f[cg][j] - dot);
    *reinterpret_cast<uint4*>(dqkv + off) = pack_bf16x8(f[cg]);
  }
}
// mem_kv [2][heads][dh][n_mem] gradient; grid (heads, B), block dh*n_mem threads (<= 256); samples add atomically
// (one block per head looping over the batch was a 50 us serial chain per layer)
__global__ void linattn_mem_bwd_kernel(const float* __restrict__ mem_kv, int n_mem, const float* __restrict__ kstat,
                                       const float* __restrict__ dctx, const float* __restrict__ ssum, int B, int heads,
                                       int dh, float* __restrict__ dmem) {
  __shared__ float s_kt[32 * 8];   // softmaxed memory keys kt[d][j] of this (head, sample)
  const int h = blockIdx.x, b = blockIdx.y, hd = heads * dh;
  const int t = threadIdx.x;
  const bool on = t < dh * n_mem;
  const int x = on ? t / n_mem : 0, j = on ? t % n_mem : 0;   // x = d for the k gradient, e for the v gradient
  const float* dc = dctx + ((size_t)b * heads + h) * dh * dh;
  const float* ks = kstat + ((size_t)b * hd + h * dh) * 2;
  float kt = 0.f;
  if (on) {
    kt = __expf(mem_kv[((size_t)h * dh + x) * n_mem + j] - ks[2 * x]) / ks[2 * x + 1];
    s_kt[x * n_mem + j] = kt;
  }
  __syncthreads();
  if (!on) return;
  float a = 0.f, g = 0.f;
#pragma unroll 8
  for (int e = 0; e < dh; ++e) {
    a = fmaf(__ldg(mem_kv + ((size_t)hd + h * dh + e) * n_mem + j), __ldg(dc + x * dh + e), a);   // dk: d = x
    g = fmaf(s_kt[e * n_mem + j], __ldg(dc + e * dh + x), g);                                     // dv: e = x
  }
  atomicAdd(dmem + ((size_t)h * dh + x) * n_mem + j, kt * (a - ssum[(size_t)b * hd + h * dh + x]));
  atomicAdd(dmem + ((size_t)hd + h * dh + x) * n_mem + j, g);
}

// ---------------------------------------------------------------- softmax Attention backward
// tokens n (+ n_mem memory kv, mem_kv[2][heads][n_mem][dh]).  Pass A (one warp per query): recompute
// P = softmax(q k^T scale), dP = dO v^T, dS = P (dP - dO.O), dq = scale dS k; P and dS go to scratch.
// Pass B (one warp per key): dk = scale dS^T q, dv = P^T dO.
__device__ __forceinline__ float ld_ch(const bf16* t, int cgtot, size_t n, int b, int ch, size_t tok) {
  return __bfloat162float(t[(((size_t)b * cgtot + (ch >> 3)) * n + tok) * 8 + (ch & 7)]);
}
__global__ void __launch_bounds__(128)
attn_bwd_q_kernel(const bf16* __restrict__ qkv, int cgtot, const bf16* __restrict__ ao, const bf16* __restrict__ dao,
                  int ocgtot, int heads, int dh, int n, const float* __restrict__ mem_kv, int n_mem, float scale,
                  float* __restrict__ P, float* __restrict__ dS, bf16* __restrict__ dqkv) {
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads, hd = heads * dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  if (i >= n) return;
  const int nk = n + n_mem;
  // lane = dim (dh <= 32)
  const float qd = lane < dh ? ld_ch(qkv, cgtot, n, b, h * dh + lane, i) : 0.f;
  const float od = lane < dh ? ld_ch(ao, ocgtot, n, b, h * dh + lane, i) : 0.f;
  const float gd = lane < dh ? ld_ch(dao, ocgtot, n, b, h * dh + lane, i) : 0.f;
  const float D = warp_sum(od * gd);
  float* Pr = P + ((size_t)bh * n + i) * nk;
  float* Sr = dS + ((size_t)bh * n + i) * nk;
  float mx = -INFINITY;
  for (int j = 0; j < nk; ++j) {
    float kd = 0.f, vd = 0.f;
    if (lane < dh) {
      if (j < n_mem) {
        kd = mem_kv[((size_t)h * n_mem + j) * dh + lane];
        vd = mem_kv[(((size_t)heads + h) * n_mem + j) * dh + lane];
      } else {
        kd = ld_ch(qkv, cgtot, n, b, hd + h * dh + lane, j - n_mem);
        vd = ld_ch(qkv, cgtot, n, b, 2 * hd + h * dh + lane, j - n_mem);
      }
    }
    const float s = warp_sum(qd * kd) * scale;
    const float dp = warp_sum(gd * vd);
    if (lane == 0) { Pr[j] = s; Sr[j] = dp; }
    mx = fmaxf(mx, s);
  }
  __syncwarp();
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) sum += __expf(Pr[j] - mx);
  sum = warp_sum(sum);
  for (int j = lane; j < nk; j += 32) {
    const float pj = __expf(Pr[j] - mx) / sum;
    Pr[j] = pj;
    Sr[j] = pj * (Sr[j] - D);
  }
  __syncwarp();
  float dq = 0.f;
  if (lane < dh)
    for (int j = 0; j < nk; ++j) {
      const float kd = j < n_mem ? mem_kv[((size_t)h * n_mem + j) * dh + lane]
                                 : ld_ch(qkv, cgtot, n, b, hd + h * dh + lane, j - n_mem);
      dq = fmaf(Sr[j], kd, dq);
    }
  if (lane < dh) {
    const int ch = h * dh + lane;
    dqkv[(((size_t)b * cgtot + (ch >> 3)) * n + i) * 8 + (ch & 7)] = __float2bfloat16(dq * scale);
  }
}
__global__ void __launch_bounds__(128)
attn_bwd_kv_kernel(const bf16* __restrict__ qkv, int cgtot, const bf16* __restrict__ dao, int ocgtot, int heads, int dh,
                   int n, int n_mem, float scale, const float* __restrict__ P, const float* __restrict__ dS,
                   bf16* __restrict__ dqkv, float* __restrict__ dmem) {
  const int bh = blockIdx.y, b = bh / heads, h = bh % heads, hd = heads * dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + warp;
  const int nk = n + n_mem;
  if (j >= nk || lane >= dh) return;
  float dk = 0.f, dv = 0.f;
  for (int i = 0; i < n; ++i) {
    const float ds = dS[((size_t)bh * n + i) * nk + j], pj = P[((size_t)bh * n + i) * nk + j];
    dk = fmaf(ds, ld_ch(qkv, cgtot, n, b, h * dh + lane, i), dk);
    dv = fmaf(pj, ld_ch(dao, ocgtot, n, b, h * dh + lane, i), dv);
  }
  dk *= scale;
  if (j < n_mem) {
    atomicAdd(dmem + ((size_t)h * n_mem + j) * dh + lane, dk);
    atomicAdd(dmem + (((size_t)heads + h) * n_mem + j) * dh + lane, dv);
  } else {
    const int ck = hd + h * dh + lane, cv = 2 * hd + h * dh + lane;
    dqkv[(((size_t)b * cgtot + (ck >> 3)) * n + (j - n_mem)) * 8 + (ck & 7)] = __float2bfloat16(dk);
    dqkv[(((size_t)b * cgtot + (cv >> 3)) * n + (j - n_mem)) * 8 + (cv & 7)] = __float2bfloat16(dv);
  }
}

// ---------------------------------------------------------------- time path backward (tiny, fp32)
// y[b][r] = sum_i W[r][i] x[b][i] + bias[r]:  dW += dy^T x, db += sum_b dy, dx (+)= dy W
__global__ void linear_bwd_w_kernel(const float* __restrict__ dy, int dy_stride, const float* __restrict__ x, int B,
                                    int rows, int cols, float* __restrict__ dw, float* __restrict__ db) {
  const size_t total = (size_t)rows * cols;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(o / cols), i = (int)(o % cols);
    float a = 0.f, s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float d = dy[(size_t)b * dy_stride + r];
      a = fmaf(d, x[(size_t)b * cols + i], a);
      s += d;
    }
    dw[o] += a;
    if (i == 0 && db) db[r] += s;
  }
}
// grid (cols / 32, B), block 256 = 32 columns x 8 row slices (a serial loop over up to 384 rows per thread made each of
// the 25 launches of a step a 27 us latency chain)
__global__ void __launch_bounds__(256)
linear_bwd_x_kernel(const float* __restrict__ dy, int dy_stride, const float* __restrict__ w, int B,
                    int rows, int cols, float* __restrict__ dx, int add) {
  __shared__ float s_part[8][32];
  const int b = blockIdx.y;
  const int ci = threadIdx.x & 31, rs = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ci;
  float a0 = 0.f, a1 = 0.f;
  if (i < cols) {
    int r = rs;
    for (; r + 8 < rows; r += 16) {
      a0 = fmaf(__ldg(dy + (size_t)b * dy_stride + r), __ldg(w + (size_t)r * cols + i), a0);
      a1 = fmaf(__ldg(dy + (size_t)b * dy_stride + r + 8), __ldg(w + (size_t)(r + 8) * cols + i), a1);
    }
    if (r < rows) a0 = fmaf(__ldg(dy + (size_t)b * dy_stride + r), __ldg(w + (size_t)r * cols + i), a0);
  }
  s_part[rs][ci] = a0 + a1;
  __syncthreads();
  if (rs == 0 && i < cols) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a += s_part[k][ci];
    if (add) dx[(size_t)b * cols + i] += a; else dx[(size_t)b * cols + i] = a;
  }
}
// mode 0: g *= silu'(pre)   mode 1: g *= gelu_erf'(pre)
__global__ void act_bwd_kernel(float* __restrict__ g, const float* __restrict__ pre, size_t n, int mode) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float z = pre[i];
  float d;
  if (mode == 0) {
    const float s = 1.f / (1.f + expf(-z));
    d = s * (1.f + z * (1.f - s));
  } else {
    d = 0.5f * (1.f + erff(z * 0.70710678118654752440f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
  }
  g[i] *= d;
}
// y = sqrt2 cos(t f + phi): dphi[i] += sum_b dy * (-sqrt2 sin(arg)), df[i] += sum_b dy * (-sqrt2 sin(arg)) t[b]
__global__ void fourier_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ t, const float* __restrict__ f,
                                   const float* __restrict__ phi, int B, int n, float* __restrict__ df,
                                   float* __restrict__ dphi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f, c = 0.f;
  for (int b = 0; b < B; ++b) {
    const float arg = __fadd_rn(__fmul_rn(t[b], f[i]), phi[i]);
    const float g = dy[(size_t)b * n + i] * (-1.41421356237309504880f) * sinf(arg);
    a += g * t[b];
    c += g;
  }
  df[i] += a;
  dphi[i] += c;
}

// ---------------------------------------------------------------- NCDHW fp32 gradient -> blocked bf16
// (pack_ncdhw_to_blocked does this; declared in ops.h)

// ---------------------------------------------------------------- optimiser
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, double* __restrict__ acc) {
  __shared__ double red[8];
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = g[i];
    s += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    atomicAdd(acc, t);
  }
}
// torch.nn.utils.clip_grad_norm_ (coef = min(1, max_norm / (norm + 1e-6))) + torch.optim.Adam / AdamW
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float b1, float b2, float eps, float wd, int decoupled, float bc1, float bc2,
                            const double* __restrict__ sumsq, float gscale, float max_norm) {
  float coef = gscale;
  if (sumsq && max_norm > 0.f) {
    const float norm = (float)sqrt(*sumsq) * gscale;
    coef *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float gi = g[i] * coef, pi = p[i];
    if (wd != 0.f) {
      if (decoupled) pi *= 1.f - lr * wd; else gi = fmaf(wd, pi, gi);
    }
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// d/dvhat of loss = sum (v - vhat)^2 / sum v^2  (acc2 = the two sums): dout = 2 (vhat - v) / acc2[1]
__global__ void mse_ratio_grad_kernel(const float* __restrict__ v, const float* __restrict__ vh, size_t n,
                                      const double* __restrict__ acc2, float gscale, float* __restrict__ dout) {
  const float k = (float)(2.0 / acc2[1]) * gscale;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dout[i] = k * (vh[i] - v[i]);
}

inline int grid1(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

#define FTB_NA_FLAGS(KERNEL, LPVV)                                                              \
  do {                                                                                          \
    const int code = (norm ? 4 : 0) | (silu ? 2 : 0) | (dr ? 1 : 0);                            \
    switch (code) {                                                                             \
      case 0: KERNEL<LPVV, false, false, false><<<grid, 256, 0, st>>>(p, iters); break;         \
      case 1: KERNEL<LPVV, false, false, true><<<grid, 256, 0, st>>>(p, iters); break;          \
      case 2: KERNEL<LPVV, false, true, false><<<grid, 256, 0, st>>>(p, iters); break;          \
      case 3: KERNEL<LPVV, false, true, true><<<grid, 256, 0, st>>>(p, iters); break;           \
      case 4: KERNEL<LPVV, true, false, false><<<grid, 256, 0, st>>>(p, iters); break;          \
      case 5: KERNEL<LPVV, true, false, true><<<grid, 256, 0, st>>>(p, iters); break;           \
      case 6: KERNEL<LPVV, true, true, false><<<grid, 256, 0, st>>>(p, iters); break;           \
      default: KERNEL<LPVV, true, true, true><<<grid, 256, 0, st>>>(p, iters); break;           \
    }                                                                                           \
  } while (0)

template <typename F8, typename F16, typename F32>
static inline void normact_dispatch(int CG, F8 f8, F16 f16, F32 f32) {
  if (CG <= 8) f8(); else if (CG <= 16) f16(); else f32();
}
static inline void normact_grid(const Act& u, int CG, dim3* grid, int* iters) {
  const int lpv = CG <= 8 ? 8 : (CG <= 16 ? 16 : 32);
  const size_t vpb = 256 / lpv;
  const size_t vblocks = (u.voxels() + vpb - 1) / vpb;
  int it = (int)(vblocks * u.B / (16 * (size_t)num_sms()));   // ~16 blocks per SM over the batch
  it = it < 1 ? 1 : (it > 32 ? 32 : it);
  *iters = it;
  *grid = dim3((unsigned)((vblocks + it - 1) / it), u.B);
}

int normact_fwd(const Act& u, bool norm, const float* gain, const float* s1, const float* sh, int fstride, bool silu,
                const Act* resid, Act& out, cudaStream_t st, float drop_p, unsigned long long drop_key) {
  FTB_CHECK(out.B == u.B && out.C == u.C && out.voxels() == u.voxels(), "normact: shapes");
  FTB_CHECK(u.cg() <= 32, "normact: at most 256 channels");
  if (resid) FTB_CHECK(resid->C == u.C && resid->voxels() == u.voxels(), "normact: residual shape");
  NormActP p{};
  p.u = u.p; p.out = out.p; p.resid = resid ? resid->p : nullptr;
  p.CG = u.cg(); p.vox = u.voxels(); p.norm = norm; p.silu = silu;
  p.gain = gain; p.s1 = s1; p.sh = sh; p.fstride = fstride;
  p.drop_p = drop_p; p.drop_key = drop_key;
  dim3 grid;
  int iters;
  normact_grid(u, p.CG, &grid, &iters);
  const bool dr = drop_p > 0.f;
  normact_dispatch(p.CG, [&] { FTB_NA_FLAGS(normact_fwd_kernel, 8); }, [&] { FTB_NA_FLAGS(normact_fwd_kernel, 16); },
                   [&] { FTB_NA_FLAGS(normact_fwd_kernel, 32); });
  FTB_LAUNCH_OK();
  return 0;
}

int normact_bwd(const Act& dout, const Act& u, bool norm, const float* gain, const float* s1, const float* sh,
                int fstride, bool silu, Act& du, float* R, float* S, int sstride, float* dbias, cudaStream_t st,
                float drop_p, unsigned long long drop_key) {
  FTB_CHECK(dout.C == u.C && du.C == u.C && dout.voxels() == u.voxels(), "normact_bwd: shapes");
  FTB_CHECK(u.cg() <= 32, "normact_bwd: at most 256 channels");
  NormActP p{};
  p.u = u.p; p.dout = dout.p; p.du = du.p;
  p.CG = u.cg(); p.vox = u.voxels(); p.norm = norm; p.silu = silu;
  p.gain = gain; p.s1 = s1; p.sh = sh; p.fstride = fstride;
  p.R = R; p.S = S; p.sstride = sstride; p.dbias = dbias;
  p.drop_p = drop_p; p.drop_key = drop_key;
  dim3 grid;
  int iters;
  const bool dr = drop_p > 0.f;
  normact_grid(u, p.CG, &grid, &iters);
  normact_dispatch(p.CG, [&] { FTB_NA_FLAGS(normact_bwd_kernel, 8); }, [&] { FTB_NA_FLAGS(normact_bwd_kernel, 16); },
                   [&] { FTB_NA_FLAGS(normact_bwd_kernel, 32); });
  FTB_LAUNCH_OK();
  return 0;
}

int normact_finish(const float* R, int B, int C, const float* gain, const float* s1, int fstride, float sqrt_c,
                   float* ds1, float* dg, cudaStream_t st) {
  normact_finish_kernel<<<cdiv(C, 128), 128, 0, st>>>(R, B, C, gain, s1, fstride, sqrt_c, ds1, dg);
  FTB_LAUNCH_OK();
  return 0;
}

int bias_grad(const Act& dy, int cgoff, int C, float* db, cudaStream_t st) {
  const size_t vox = dy.voxels();
  int bx = (int)((vox + 256 * 8 - 1) / (256 * 8));
  bx = bx < 1 ? 1 : (bx > 64 ? 64 : bx);
  dim3 grid(bx, cdiv(C, 8), dy.B);
  chan_sum_kernel<<<grid, 256, 0, st>>>(dy.p, dy.cg(), cgoff, vox, C, db);
  FTB_LAUNCH_OK();
  return 0;
}

int act_accum(Act& dst, const Act& src, bool add, cudaStream_t st) {
  FTB_CHECK(dst.elems() == src.elems(), "accum: shapes");
  const size_t n8 = dst.elems() / 8;
  accum_kernel<<<grid1(n8, 256), 256, 0, st>>>(dst.p, src.p, n8, add ? 1 : 0);
  FTB_LAUNCH_OK();
  return 0;
}

// din (+)= adjoint of trilinear_resample(in -> out) applied to dout
int trilinear_resample_bwd(const Act& dout, Act& din, bool add, cudaStream_t st) {
  FTB_CHECK(dout.B == din.B && dout.C == din.C, "trilinear_bwd: shapes");
  const float sd = dout.D > 1 ? (float)(din.D - 1) / (float)(dout.D - 1) : 0.f;
  const float sh = dout.H > 1 ? (float)(din.H - 1) / (float)(dout.H - 1) : 0.f;
  const float sw = dout.W > 1 ? (float)(din.W - 1) / (float)(dout.W - 1) : 0.f;
  dim3 grid((unsigned)((din.voxels() + 255) / 256), din.B * din.cg());
  trilinear_bwd_kernel<<<grid, 256, 0, st>>>(dout.p, din.D, din.H, din.W, dout.D, dout.H, dout.W, sd, sh, sw, din.p,
                                             add ? 1 : 0);
  FTB_LAUNCH_OK();
  return 0;
}

int transpose_flip(const float* w, int cout, int cin, int ksize, int ci0, int cin_sub, const float* row_scale,
                   float* wt, cudaStream_t st) {
  const int k3 = ksize * ksize * ksize;
  const size_t total = (size_t)cin_sub * cout * k3;
  transpose_flip_kernel<<<grid1(total, 256), 256, 0, st>>>(w, cout, cin, k3, ci0, cin_sub, row_scale, wt);
  FTB_LAUNCH_OK();
  return 0;
}

int fold_gain_bwd(const float* dwp, const float* w, const float* gs, int cout, int cin, float sqrt_c, float* dw,
                  float* dg, cudaStream_t st) {
  fold_gain_bwd_kernel<<<cin, 128, 0, st>>>(dwp, w, gs, cout, cin, sqrt_c, dw, dg);
  FTB_LAUNCH_OK();
  return 0;
}

int blockdiag(const float* src, int B, int heads, int dh, bool transpose, float mul, float* out, cudaStream_t st) {
  dim3 grid(cdiv(heads * dh * heads * dh, 256), B);
  blockdiag_kernel<<<grid, 256, 0, st>>>(src, heads, dh, transpose ? 1 : 0, mul, out);
  FTB_LAUNCH_OK();
  return 0;
}
int dctx_extract(const float* full, const float* ctx, int B, int heads, int dh, float* dctx, float* ssum,
                 cudaStream_t st) {
  dctx_extract_kernel<<<dim3(heads, B), 32, 0, st>>>(full, ctx, heads, dh, dctx, ssum);
  FTB_LAUNCH_OK();
  return 0;
}
int ksoftmax_apply(Act& qkv, int hd, const float* kstat, cudaStream_t st) {
  const size_t vox = qkv.voxels();
  int bx = (int)((vox + 1023) / 1024);
  bx = bx < 1 ? 1 : (bx > 128 ? 128 : bx);
  ksoftmax_apply_kernel<<<dim3(bx, hd / 8, qkv.B), 256, 0, st>>>(qkv.p, qkv.cg(), hd / 8, hd / 8, vox, kstat, hd);
  FTB_LAUNCH_OK();
  return 0;
}
int ksoftmax_bwd(Act& dqkv, const Act& qkv, int hd, const float* ssum, cudaStream_t st) {
  const size_t vox = qkv.voxels();
  int bx = (int)((vox + 1023) / 1024);
  bx = bx < 1 ? 1 : (bx > 128 ? 128 : bx);
  ksoftmax_bwd_kernel<<<dim3(bx, hd / 8, qkv.B), 256, 0, st>>>(dqkv.p, qkv.p, qkv.cg(), hd / 8, vox, ssum, hd);
  FTB_LAUNCH_OK();
  return 0;
}
int qsoftmax_bwd(Act& dqkv, const Act& qkv, int heads, int dh, cudaStream_t st) {
  const size_t vox = qkv.voxels();
  FTB_CHECK(dh == 8 || dh == 16 || dh == 32, "qsoftmax_bwd: dim_head must be 8, 16 or 32");
  dim3 grid((unsigned)((vox + 255) / 256), heads, qkv.B);
  const float scale = 1.f / sqrtf((float)dh);
  if (dh == 32) qsoftmax_bwd_kernel<4><<<grid, 256, 0, st>>>(dqkv.p, qkv.p, qkv.cg(), vox, heads, scale);
  else if (dh == 16) qsoftmax_bwd_kernel<2><<<grid, 256, 0, st>>>(dqkv.p, qkv.p, qkv.cg(), vox, heads, scale);
  else qsoftmax_bwd_kernel<1><<<grid, 256, 0, st>>>(dqkv.p, qkv.p, qkv.cg(), vox, heads, scale);
  FTB_LAUNCH_OK();
  return 0;
}
int linattn_mem_bwd(const float* mem_kv, int n_mem, const float* kstat, const float* dctx, const float* ssum, int B,
                    int heads, int dh, float* dmem, cudaStream_t st) {
  FTB_CHECK(dh * n_mem <= 256 && dh <= 32 && n_mem <= 8, "linattn_mem_bwd: dim_head <= 32, num_mem_kv <= 8");
  linattn_mem_bwd_kernel<<<dim3(heads, B), 256, 0, st>>>(mem_kv, n_mem, kstat, dctx, ssum, B, heads, dh, dmem);
  FTB_LAUNCH_OK();
  return 0;
}

// scratch: 2 * B*heads*n*(n+n_mem) floats
int full_attention_bwd(const Act& qkv, const Act& ao, const Act& dao, int heads, int dh, const float* mem_kv,
                       int n_mem, float* scratch, Act& dqkv, float* dmem, cudaStream_t st) {
  const int n = (int)qkv.voxels();
  FTB_CHECK(dh <= 32, "attention_bwd: dim_head must be <= 32");
  const float scale = 1.f / sqrtf((float)dh);
  float* P = scratch;
  float* dS = scratch + (size_t)qkv.B * heads * n * (n + n_mem);
  attn_bwd_q_kernel<<<dim3(cdiv(n, 4), qkv.B * heads), 128, 0, st>>>(qkv.p, qkv.cg(), ao.p, dao.p, ao.cg(), heads, dh, n,
                                                                      mem_kv, n_mem, scale, P, dS, dqkv.p);
  FTB_LAUNCH_OK();
  attn_bwd_kv_kernel<<<dim3(cdiv(n + n_mem, 4), qkv.B * heads), 128, 0, st>>>(qkv.p, qkv.cg(), dao.p, ao.cg(), heads, dh,
                                                                              n, n_mem, scale, P, dS, dqkv.p, dmem);
  FTB_LAUNCH_OK();
  return 0;
}

int linear_bwd(const float* dy, int dy_stride, const float* x, const float* w, int B, int rows, int cols, float* dw,
               float* db, float* dx, bool dx_add, cudaStream_t st) {
  if (dw) {
    linear_bwd_w_kernel<<<grid1((size_t)rows * cols, 256), 256, 0, st>>>(dy, dy_stride, x, B, rows, cols, dw, db);
    FTB_LAUNCH_OK();
  }
  if (dx) {
    linear_bwd_x_kernel<<<dim3(cdiv(cols, 32), B), 256, 0, st>>>(dy, dy_stride, w, B, rows, cols, dx, dx_add ? 1 : 0);
    FTB_LAUNCH_OK();
  }
  return 0;
}
int act_bwd(float* g, const float* pre, size_t n, int mode, cudaStream_t st) {
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, pre, n, mode);
  FTB_LAUNCH_OK();
  return 0;
}
int fourier_bwd(const float* dy, const float* t, const float* f, const float* phi, int B, int n, float* df, float* dphi,
                cudaStream_t st) {
  fourier_bwd_kernel<<<cdiv(n, 128), 128, 0, st>>>(dy, t, f, phi, B, n, df, dphi);
  FTB_LAUNCH_OK();
  return 0;
}

int mse_ratio_grad(const float* v, const float* vhat, long long n, const double* acc2, float gscale, float* dout,
                   cudaStream_t st) {
  mse_ratio_grad_kernel<<<grid1((size_t)n, 256), 256, 0, st>>>(v, vhat, (size_t)n, acc2, gscale, dout);
  FTB_LAUNCH_OK();
  return 0;
}

int grad_sumsq(const float* g, long long n, double* acc, cudaStream_t st) {
  sumsq_kernel<<<grid1((size_t)n, 256), 256, 0, st>>>(g, (size_t)n, acc);
  FTB_LAUNCH_OK();
  return 0;
}
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
              float wd, int decoupled, int step, const double* sumsq, float gscale, float max_norm, cudaStream_t st) {
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  adam_kernel<<<grid1((size_t)n, 256), 256, 0, st>>>(p, g, m, v, (size_t)n, lr, b1, b2, eps, wd, decoupled, bc1, bc2,
                                                    sumsq, gscale, max_norm);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
