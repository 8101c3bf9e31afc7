// fp32 mode of the velocity field: everything AROUND the convolutions in plain fp32 on NCDHW tensors (the
// reference's own layout), the convolutions on the tcgen05 kernel with the 3 x bf16 split
//     W x  ~=  W_hi x_hi + W_hi x_lo + W_lo x_hi        (hi = bf16(.), lo = bf16(. - hi); fp32 accumulation),
// which keeps 16 mantissa bits of both operands: the bar is <= 1e-4 relative L2 against the fp32 reference
// (BASELINE north_star; plain bf16 operands give 5.6e-3, TF32 7e-4 — SURVEY §6).  This is the accuracy mode, measured at
// 3.8x the cost of the bf16 path (14.3 vs 3.8 ms per 64^3 evaluation at B=1); the kernels here are simple coalesced fp32 passes (thread = voxel, channel stride =
// voxels).  Reference: unet_attn_3d.py RMSNorm :111-128, Block :232-244, LinearAttention :308-341, Attention
// :357-373 / :436-465, Upsample / Downsample :85-88, :105-108.
#include "ops.h"

namespace ftb {

namespace {

__global__ void pack_split_kernel(const float* __restrict__ x0, int c0, const float* __restrict__ x1, int c1, int B,
                                  size_t vox, int CGP, bf16* __restrict__ out) {
  const size_t total = (size_t)B * CGP * vox;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t v = i % vox;
    const int cg = (int)((i / vox) % CGP);
    const int b = (int)(i / (vox * CGP));
    float hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      float x = 0.f;
      if (c < c0) x = __ldg(x0 + ((size_t)b * c0 + c) * vox + v);
      else if (c < c0 + c1) x = __ldg(x1 + ((size_t)b * c1 + (c - c0)) * vox + v);
      const float h = __bfloat162float(__float2bfloat16(x));
      hi[j] = h;
      lo[j] = x - h;
    }
    bf16* ob = out + ((size_t)b * 2 * CGP * vox) * 8;
    *reinterpret_cast<uint4*>(ob + ((size_t)cg * vox + v) * 8) = pack_bf16x8(hi);
    *reinterpret_cast<uint4*>(ob + ((size_t)(CGP + cg) * vox + v) * 8) = pack_bf16x8(lo);
  }
}

__global__ void __launch_bounds__(256)
normact_f32_kernel(const float* __restrict__ u, int C, size_t vox, int norm, const float* __restrict__ gain,
                   const float* __restrict__ s1, const float* __restrict__ sh, int fstride, int silu,
                   const float* __restrict__ resid, float* __restrict__ out) {
  const int b = blockIdx.y;
  const size_t v = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= vox) return;
  const float* ub = u + (size_t)b * C * vox + v;
  float rinv = 1.f;
  if (norm) {
    float ss = 0.f;
    for (int c = 0; c < C; ++c) {
      const float x = ub[(size_t)c * vox];
      ss += x * x;
    }
    rinv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  for (int c = 0; c < C; ++c) {
    float y = ub[(size_t)c * vox] * rinv;
    if (gain) y *= gain[c];
    if (s1) y = y * s1[(size_t)b * fstride + c] + sh[(size_t)b * fstride + c];
    if (silu) y = y / (1.f + expf(-y));
    if (resid) y += resid[((size_t)b * C + c) * vox + v];
    out[((size_t)b * C + c) * vox + v] = y;
  }
}

__global__ void trilinear_f32_kernel(const float* __restrict__ in, size_t BC, int Di, int Hi, int Wi, int Do, int Ho,
                                     int Wo, float sd, float sh, float sw, float* __restrict__ out) {
  const size_t vo = (size_t)Do * Ho * Wo, vi = (size_t)Di * Hi * Wi;
  const size_t total = BC * vo;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bc = i / vo, r = i % vo;
    const int w = (int)(r % Wo), h = (int)((r / Wo) % Ho), d = (int)(r / ((size_t)Wo * Ho));
    // index rule as ATen (align_corners=True): src = scale * dst, i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0
    const float fd = sd * d, fh = sh * h, fw = sw * w;
    const int d0 = (int)fd, h0 = (int)fh, w0 = (int)fw;
    const int d1 = d0 + (d0 < Di - 1), h1 = h0 + (h0 < Hi - 1), w1 = w0 + (w0 < Wi - 1);
    const float ld1 = fd - d0, lh1 = fh - h0, lw1 = fw - w0;
    const float ld0 = 1.f - ld1, lh0 = 1.f - lh1, lw0 = 1.f - lw1;
    const float* p = in + bc * vi;
    auto at = [&](int dd, int hh, int ww) { return p[((size_t)dd * Hi + hh) * Wi + ww]; };
    out[i] = ld0 * (lh0 * (lw0 * at(d0, h0, w0) + lw1 * at(d0, h0, w1)) + lh1 * (lw0 * at(d0, h1, w0) + lw1 * at(d0, h1, w1))) +
             ld1 * (lh0 * (lw0 * at(d1, h0, w0) + lw1 * at(d1, h0, w1)) + lh1 * (lw0 * at(d1, h1, w0) + lw1 * at(d1, h1, w1)));
  }
}

// q third, in place: softmax over the dh channels of each head, times dh^-0.5; thread = (b, head, voxel)
__global__ void linattn_q_f32_kernel(float* __restrict__ qkv, int heads, int dh, size_t n, float scale) {
  const int b = blockIdx.z, h = blockIdx.y;
  const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  float* q = qkv + ((size_t)b * 3 * heads * dh + (size_t)h * dh) * n + v;
  float mx = -INFINITY;
  for (int d = 0; d < dh; ++d) mx = fmaxf(mx, q[(size_t)d * n]);
  float s = 0.f;
  for (int d = 0; d < dh; ++d) s += expf(q[(size_t)d * n] - mx);
  const float inv = scale / s;
  for (int d = 0; d < dh; ++d) q[(size_t)d * n] = expf(q[(size_t)d * n] - mx) * inv;
}
// block per (b, k channel): stat[b][ch] = (max, sum exp) over voxels + memory tokens
__global__ void __launch_bounds__(1024)
linattn_kstat_f32_kernel(const float* __restrict__ qkv, int hd, size_t n, const float* __restrict__ mem_kv, int n_mem,
                         float* __restrict__ stat) {
  __shared__ float red[32];
  const int ch = blockIdx.x, b = blockIdx.y;
  const float* k = qkv + ((size_t)b * 3 * hd + hd + ch) * n;
  const float* mk = mem_kv + (size_t)ch * n_mem;   // mem_kv[0][h][d][j], (h, d) = ch
  float mx = -INFINITY;
  for (size_t i = threadIdx.x; i < n; i += 1024) mx = fmaxf(mx, k[i]);
  for (int j = 0; j < n_mem; ++j) mx = fmaxf(mx, mk[j]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (size_t i = threadIdx.x; i < n; i += 1024) s += expf(k[i] - mx);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 32; ++w) t += red[w];
    for (int j = 0; j < n_mem; ++j) t += expf(mk[j] - mx);
    stat[((size_t)b * hd + ch) * 2] = mx;
    stat[((size_t)b * hd + ch) * 2 + 1] = t;
  }
}
// Context of one (sample, head): ctx[d][e] = sum_n softmax_n(k)[d,n] v[e,n].  Block = (voxel chunk, head, sample):
// the chunk's P = exp(k - max) [dh][TN] and V [dh][TN] tiles are staged in shared memory (k and v are read from HBM
// exactly once, one expf per element) and each of the 256 threads accumulates a 2 x 2 patch of the dh x dh result;
// per-chunk partials go to `part` and are summed in a fixed order by linattn_ctx_finish_f32_kernel (deterministic).
constexpr int kCtxTN = 128;    // voxels per staged tile
constexpr int kCtxChunk = 2048;  // voxels per block
__global__ void __launch_bounds__(256)
linattn_ctx_part_f32_kernel(const float* __restrict__ qkv, int heads, int dh, size_t n, const float* __restrict__ stat,
                            float* __restrict__ part, int nchunk) {
  __shared__ float sp[32][kCtxTN + 1], sv[32][kCtxTN + 1];
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z, hd = heads * dh;
  const float* kb = qkv + ((size_t)b * 3 * hd + hd + (size_t)h * dh) * n;
  const float* vb = qkv + ((size_t)b * 3 * hd + 2 * hd + (size_t)h * dh) * n;
  const int d0 = (threadIdx.x >> 4) * 2, e0 = (threadIdx.x & 15) * 2;   // 16 x 16 threads, 2 x 2 outputs each
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
  const size_t v_lo = (size_t)chunk * kCtxChunk;
  const size_t v_hi = v_lo + kCtxChunk < n ? v_lo + kCtxChunk : n;
  for (size_t t0 = v_lo; t0 < v_hi; t0 += kCtxTN) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * kCtxTN; i += 256) {
      const int r = i / kCtxTN, c = i - r * kCtxTN;
      const size_t v = t0 + c;
      const bool in = r < dh && v < v_hi;
      sp[r][c] = in ? expf(kb[(size_t)r * n + v] - stat[((size_t)b * hd + h * dh + r) * 2]) : 0.f;
      sv[r][c] = in ? vb[(size_t)r * n + v] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < kCtxTN; ++c) {
      const float p0 = sp[d0][c], p1 = sp[d0 + 1][c], v0 = sv[e0][c], v1 = sv[e0 + 1][c];
      a00 = fmaf(p0, v0, a00); a01 = fmaf(p0, v1, a01);
      a10 = fmaf(p1, v0, a10); a11 = fmaf(p1, v1, a11);
    }
  }
  float* o = part + (((size_t)b * heads + h) * nchunk + chunk) * 1024;
  o[d0 * 32 + e0] = a00; o[d0 * 32 + e0 + 1] = a01;
  o[(d0 + 1) * 32 + e0] = a10; o[(d0 + 1) * 32 + e0 + 1] = a11;
}
// ctx[b][h][d][e] = (sum_chunks part + memory tokens) / den[d]; block per (h, b), thread per (d, e)
__global__ void __launch_bounds__(1024)
linattn_ctx_finish_f32_kernel(const float* __restrict__ part, int nchunk, int heads, int dh,
                              const float* __restrict__ mem_kv, int n_mem, const float* __restrict__ stat,
                              float* __restrict__ ctx) {
  const int h = blockIdx.x, b = blockIdx.y, hd = heads * dh;
  const int d = threadIdx.x >> 5, e = threadIdx.x & 31;
  if (d >= dh || e >= dh) return;
  const float* pp = part + ((size_t)b * heads + h) * nchunk * 1024 + d * 32 + e;
  float t = 0.f;
  for (int c = 0; c < nchunk; ++c) t += pp[(size_t)c * 1024];
  const float mx = stat[((size_t)b * hd + h * dh + d) * 2], den = stat[((size_t)b * hd + h * dh + d) * 2 + 1];
  const float* mk = mem_kv + ((size_t)h * dh + d) * n_mem;
  const float* mv = mem_kv + ((size_t)hd + h * dh + e) * n_mem;
  for (int j = 0; j < n_mem; ++j) t += expf(mk[j] - mx) * mv[j];
  ctx[(((size_t)b * heads + h) * dh + d) * dh + e] = t / den;
}
// out[b][(h,e)][n] = sum_d ctx[b][h][d][e] q~[(h,d)][n]; thread = (b, h, voxel)
__global__ void __launch_bounds__(256)
linattn_out_f32_kernel(const float* __restrict__ qkv, int heads, int dh, size_t n, const float* __restrict__ ctx,
                       float* __restrict__ out) {
  __shared__ float sc[32 * 32];
  const int b = blockIdx.z, h = blockIdx.y, hd = heads * dh;
  for (int i = threadIdx.x; i < dh * dh; i += 256) sc[i] = ctx[((size_t)b * heads + h) * dh * dh + i];
  __syncthreads();
  const size_t v = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (v >= n) return;
  const float* q = qkv + ((size_t)b * 3 * hd + (size_t)h * dh) * n + v;
  float qd[32];
  for (int d = 0; d < dh; ++d) qd[d] = q[(size_t)d * n];
  for (int e = 0; e < dh; ++e) {
    float a = 0.f;
    for (int d = 0; d < dh; ++d) a += sc[d * dh + e] * qd[d];
    out[((size_t)b * hd + h * dh + e) * n + v] = a;
  }
}

// thread = (b, h, query): online softmax over n + n_mem keys, dh <= 32
__global__ void __launch_bounds__(64)
full_attn_f32_kernel(const float* __restrict__ qkv, int heads, int dh, int n, const float* __restrict__ mem_kv, int n_mem,
                     float scale, float* __restrict__ out) {
  const int b = blockIdx.z, h = blockIdx.y, hd = heads * dh;
  const int i = blockIdx.x * 64 + threadIdx.x;
  if (i >= n) return;
  const float* qb = qkv + ((size_t)b * 3 * hd + h * dh) * n;
  const float* kb = qkv + ((size_t)b * 3 * hd + hd + h * dh) * n;
  const float* vb = qkv + ((size_t)b * 3 * hd + 2 * hd + h * dh) * n;
  float q[32], o[32];
  for (int d = 0; d < dh; ++d) { q[d] = qb[(size_t)d * n + i]; o[d] = 0.f; }
  float m = -INFINITY, l = 0.f;
  for (int j = 0; j < n + n_mem; ++j) {
    float s = 0.f;
    if (j < n_mem) for (int d = 0; d < dh; ++d) s += q[d] * mem_kv[((size_t)h * n_mem + j) * dh + d];
    else for (int d = 0; d < dh; ++d) s += q[d] * kb[(size_t)d * n + (j - n_mem)];
    s *= scale;
    const float mn = fmaxf(m, s);
    const float c = expf(m - mn), pj = expf(s - mn);
    l = l * c + pj;
    if (j < n_mem) for (int d = 0; d < dh; ++d) o[d] = o[d] * c + pj * mem_kv[(((size_t)heads + h) * n_mem + j) * dh + d];
    else for (int d = 0; d < dh; ++d) o[d] = o[d] * c + pj * vb[(size_t)d * n + (j - n_mem)];
    m = mn;
  }
  for (int d = 0; d < dh; ++d) out[((size_t)b * hd + h * dh + d) * n + i] = o[d] / l;
}

inline int grid1(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  const size_t cap = (size_t)num_sms() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace

int f32_pack_split(const float* x0, int c0, const float* x1, int c1, int B, size_t vox, Act& out, cudaStream_t st) {
  const int CP = round_up(c0 + c1, 16);
  FTB_CHECK(out.C == 2 * CP && out.B == B && out.voxels() == vox, "f32_pack_split: output must have 2*pad16(C) channels");
  pack_split_kernel<<<grid1((size_t)B * (CP / 8) * vox, 256), 256, 0, st>>>(x0, c0, x1, c1, B, vox, CP / 8, out.p);
  FTB_LAUNCH_OK();
  return 0;
}

int f32_normact(const float* u, int B, int C, size_t vox, bool norm, const float* gain, const float* s1, const float* sh,
                int fstride, bool silu, const float* resid, float* out, cudaStream_t st) {
  dim3 grid((unsigned)((vox + 255) / 256), B);
  normact_f32_kernel<<<grid, 256, 0, st>>>(u, C, vox, norm ? 1 : 0, gain, s1, sh, fstride, silu ? 1 : 0, resid, out);
  FTB_LAUNCH_OK();
  return 0;
}

int f32_trilinear(const float* in, int B, int C, int Di, int Hi, int Wi, int Do, int Ho, int Wo, float* out,
                  cudaStream_t st) {
  const float sd = Do > 1 ? (float)(Di - 1) / (float)(Do - 1) : 0.f;
  const float sh = Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f;
  const float sw = Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f;
  const size_t total = (size_t)B * C * Do * Ho * Wo;
  trilinear_f32_kernel<<<grid1(total, 256), 256, 0, st>>>(in, (size_t)B * C, Di, Hi, Wi, Do, Ho, Wo, sd, sh, sw, out);
  FTB_LAUNCH_OK();
  return 0;
}

int f32_linattn_chunks(size_t n) { return (int)((n + kCtxChunk - 1) / kCtxChunk); }
// scratch floats f32_linear_attention needs: stat [B][hd][2] + ctx [B][heads][dh][dh] + per-chunk partials
size_t f32_linattn_scratch(int B, int heads, int dh, size_t n) {
  return (size_t)B * heads * dh * 2 + (size_t)B * heads * dh * dh + (size_t)B * heads * f32_linattn_chunks(n) * 1024;
}

int f32_linear_attention(float* qkv, int B, int heads, int dh, size_t n, const float* mem_kv, int n_mem, float* scratch,
                         float* out, cudaStream_t st) {
  FTB_CHECK(dh <= 32, "f32 linear attention: dim_head <= 32");
  const int hd = heads * dh;
  float* stat = scratch;
  float* ctx = scratch + (size_t)B * hd * 2;
  float* part = ctx + (size_t)B * heads * dh * dh;
  const int nchunk = f32_linattn_chunks(n);
  linattn_q_f32_kernel<<<dim3((unsigned)((n + 255) / 256), heads, B), 256, 0, st>>>(qkv, heads, dh, n, 1.f / sqrtf((float)dh));
  FTB_LAUNCH_OK();
  linattn_kstat_f32_kernel<<<dim3(hd, B), 1024, 0, st>>>(qkv, hd, n, mem_kv, n_mem, stat);
  FTB_LAUNCH_OK();
  linattn_ctx_part_f32_kernel<<<dim3(nchunk, heads, B), 256, 0, st>>>(qkv, heads, dh, n, stat, part, nchunk);
  FTB_LAUNCH_OK();
  linattn_ctx_finish_f32_kernel<<<dim3(heads, B), 1024, 0, st>>>(part, nchunk, heads, dh, mem_kv, n_mem, stat, ctx);
  FTB_LAUNCH_OK();
  linattn_out_f32_kernel<<<dim3((unsigned)((n + 255) / 256), heads, B), 256, 0, st>>>(qkv, heads, dh, n, ctx, out);
  FTB_LAUNCH_OK();
  return 0;
}

int f32_full_attention(const float* qkv, int B, int heads, int dh, int n, const float* mem_kv, int n_mem, float* out,
                       cudaStream_t st) {
  FTB_CHECK(dh <= 32, "f32 attention: dim_head <= 32");
  full_attn_f32_kernel<<<dim3(cdiv(n, 64), heads, B), 64, 0, st>>>(qkv, heads, dh, n, mem_kv, n_mem, 1.f / sqrtf((float)dh), out);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
