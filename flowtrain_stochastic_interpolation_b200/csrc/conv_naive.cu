// Direct (CUDA-core) convolution with the same operands, layouts and fused epilogue as
// conv_igemm.cu.  It exists to validate the tensor-core kernel on the device and to bisect
// failures (FTB_CONV_IMPL=naive routes the whole network through it); it is GPU code, not a
// CPU fallback, and is never the timed path.  Also holds the weight packer.
#include <cstdlib>
#include <cstring>

#include "ops.h"

namespace ftb {

namespace {

struct NaiveParams {
  int B, D, H, W, K, pad, Kw, padw;
  int cg0, cg1, s0_cgtot, s0_cgoff, s1_cgtot, s1_cgoff;
  int N, KS;
  const bf16 *src0, *src1, *wpack;
  long long w_batch_stride;
  bf16* out;
  int out_cgtot, out_cgoff;
  float* out_f32;
  int out_f32_c;
  const float *bias, *mul, *add;
  int mul_stride, add_stride, norm;
  const bf16* resid;
  int resid_cgtot, resid_cgoff;
  int prenorm, silu, qsoftmax, q_dh;
  const float* ss_in;
  float* ss_out;
  float q_scale;
};

// one thread = one output voxel x 16 output channels (chunk c0)
__device__ void naive_chunk(const NaiveParams& p, int b, int d, int h, int w, int c0, float (&acc)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const size_t cgs = (size_t)p.D * p.H * p.W;
  const bf16* wb = p.wpack + (long long)b * p.w_batch_stride;
  int t = 0;
  for (int kd = 0; kd < p.K; ++kd)
    for (int kh = 0; kh < p.K; ++kh)
      for (int kw = 0; kw < p.Kw; ++kw, ++t) {
        const int dz = d + kd - p.pad, hy = h + kh - p.pad, wx = w + kw - p.padw;
        if (dz < 0 || dz >= p.D || hy < 0 || hy >= p.H || wx < 0 || wx >= p.W) continue;
        const size_t vox = ((size_t)dz * p.H + hy) * p.W + wx;
        for (int ks = 0; ks < p.KS; ++ks) {
          float a[16];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int cgk = ks * 2 + half;
            const bf16* sp = cgk < p.cg0
                                 ? p.src0 + (((size_t)b * p.s0_cgtot + p.s0_cgoff + cgk) * cgs + vox) * 8
                                 : p.src1 + (((size_t)b * p.s1_cgtot + p.s1_cgoff + cgk - p.cg0) * cgs + vox) * 8;
            float f[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(sp)), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[half * 8 + j] = f[j];
          }
          // packed tile [N/8][2][8][8] for (kh, kw, ks, j = K-1-kd)
          const bf16* wt = wb + ((((size_t)(kh * p.Kw + kw) * p.KS + ks) * p.K + (p.K - 1 - kd)) * p.N) * 16;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = c0 + j;
            const bf16* wr = wt + (n >> 3) * 128 + (n & 7) * 8;
            float f0[8], f1[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(wr)), f0);
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(wr + 64)), f1);
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) s += a[k] * f0[k] + a[8 + k] * f1[k];
            acc[j] += s;
          }
        }
      }
}

__global__ void conv_naive_kernel(const NaiveParams p) {
  const size_t cgs = (size_t)p.D * p.H * p.W;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)p.B * cgs) return;
  const int b = (int)(gid / cgs);
  const size_t vox = gid % cgs;
  const int w = (int)(vox % p.W);
  const int h = (int)((vox / p.W) % p.H);
  const int d = (int)(vox / ((size_t)p.W * p.H));

  float rs = 1.f;
  if (p.prenorm && p.ss_in) {
    rs = 1.f / fmaxf(sqrtf(p.ss_in[(size_t)b * cgs + vox]), 1e-12f);
  } else if (p.prenorm) {
    float ss = 0.f;
    for (int cgi = 0; cgi < p.cg0; ++cgi) {
      float f[8];
      unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(
                        p.src0 + (((size_t)b * p.s0_cgtot + p.s0_cgoff + cgi) * cgs + vox) * 8)), f);
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    rs = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  float acc[16];
  float rinv = 1.f;
  if (p.norm) {
    float ss = 0.f;
    for (int c0 = 0; c0 < p.N; c0 += 16) {
      naive_chunk(p, b, d, h, w, c0, acc);
      for (int j = 0; j < 16; ++j) {
        float v = acc[j] * rs + (p.bias ? p.bias[c0 + j] : 0.f);
        ss += v * v;
      }
    }
    rinv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  }
  if (p.qsoftmax) {
    for (int hd = 0; hd < p.N / p.q_dh; ++hd) {
      float v[64];
      float mx = -INFINITY;
      for (int ch = 0; ch < p.q_dh / 16; ++ch) {
        naive_chunk(p, b, d, h, w, hd * p.q_dh + ch * 16, acc);
        for (int j = 0; j < 16; ++j) { v[ch * 16 + j] = acc[j] * rs; mx = fmaxf(mx, v[ch * 16 + j]); }
      }
      float sum = 0.f;
      for (int i = 0; i < p.q_dh; ++i) { v[i] = __expf(v[i] - mx); sum += v[i]; }
      const float inv = p.q_scale / sum;
      for (int i = 0; i < p.q_dh; ++i) {
        const int ch = hd * p.q_dh + i;
        p.out[(((size_t)b * p.out_cgtot + p.out_cgoff + (ch >> 3)) * cgs + vox) * 8 + (ch & 7)] =
            __float2bfloat16(v[i] * inv);
      }
    }
    return;
  }
  const float* mul = p.mul ? p.mul + (size_t)b * p.mul_stride : nullptr;
  const float* add = p.add ? p.add + (size_t)b * p.add_stride : nullptr;
  float ssq = 0.f;
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    naive_chunk(p, b, d, h, w, c0, acc);
    for (int j = 0; j < 16; ++j) {
      const int ch = c0 + j;
      float x = acc[j] * rs + (p.bias ? p.bias[ch] : 0.f);
      x *= rinv;
      if (mul) x *= mul[ch];
      if (add) x += add[ch];
      if (p.silu) x = silu_f(x);
      if (p.resid)
        x += __bfloat162float(
            p.resid[(((size_t)b * p.resid_cgtot + p.resid_cgoff + (ch >> 3)) * cgs + vox) * 8 + (ch & 7)]);
      ssq += x * x;
      if (p.out_f32) {
        if (ch < p.out_f32_c) p.out_f32[((size_t)b * p.out_f32_c + ch) * cgs + vox] = x;
      } else {
        p.out[(((size_t)b * p.out_cgtot + p.out_cgoff + (ch >> 3)) * cgs + vox) * 8 + (ch & 7)] =
            __float2bfloat16(x);
      }
    }
  }
  if (p.ss_out) p.ss_out[(size_t)b * cgs + vox] = ssq;
}

// fp32 [Cout][Cin][kd][kh][kw] -> bf16 [ntile][kh][kw][ks][j = K-1-kd][n/8][2][8][8]
__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin_real, int K,
                                    int cin_pad, int n, int ntiles, const float* __restrict__ in_scale,
                                    bf16* __restrict__ dst, int unfold_w) {
  const int taps = K * K * K;
  const int Kw = unfold_w ? 1 : K;
  const size_t total = (size_t)ntiles * K * K * Kw * cin_pad * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int k8 = r % 8; r /= 8;
    const int n8 = r % 8; r /= 8;
    const int kc = r % 2; r /= 2;
    const int ng = r % (n / 8); r /= (n / 8);
    const int j = r % K; r /= K;
    const int ks = r % (cin_pad / 16); r /= (cin_pad / 16);
    const int khw = r % (K * Kw); r /= (K * Kw);
    const int nt = (int)r;
    const int co = nt * n + ng * 8 + n8;
    int ci = ks * 16 + kc * 8 + k8;
    int t = (K - 1 - j) * K * K + khw;
    bool ok = co < cout && ci < cin_real;
    if (unfold_w) {   // K index kw*cin_real + ci, khw = kh
      const int kw = ci / cin_real;
      ci -= kw * cin_real;
      ok = co < cout && kw < K;
      t = (K - 1 - j) * K * K + khw * K + kw;
    }
    float v = 0.f;
    if (ok) {
      v = w[((size_t)co * cin_real + ci) * taps + t];
      if (in_scale) v *= in_scale[ci];
    }
    dst[i] = __float2bfloat16(v);
  }
}

// every weight tensor of the network in ONE launch (the training step repacks after each optimiser update):
// blockIdx.y = job.  transposed jobs build the data-gradient operand: output channel = input channel ci0 + co' of
// the forward weight, K index = forward output channel, taps flipped, optional per-row scale (folded pre-norm gain).
__global__ void pack_jobs_kernel(const PackJob* __restrict__ jobs) {
  const PackJob jb = jobs[blockIdx.y];
  const int K = jb.ksize, taps = K * K * K;
  const int Kw = jb.unfold_w ? 1 : K;
  const size_t total = (size_t)jb.ntiles * K * K * Kw * jb.cin_pad * jb.n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int k8 = r % 8; r /= 8;
    const int n8 = r % 8; r /= 8;
    const int kc = r % 2; r /= 2;
    const int ng = r % (jb.n / 8); r /= (jb.n / 8);
    const int j = r % K; r /= K;
    const int ks = r % (jb.cin_pad / 16); r /= (jb.cin_pad / 16);
    const int khw = r % (K * Kw); r /= (K * Kw);
    const int nt = (int)r;
    const int co = nt * jb.n + ng * 8 + n8;
    int ci = ks * 16 + kc * 8 + k8;
    int t = (K - 1 - j) * K * K + khw;
    int seg = 0;   // fp32 mode: K extent = 2 or 3 copies of the padded channels ([hi | hi] or [hi | hi | lo] weights)
    if (jb.part >= 3) {
      const int cp = jb.cin_pad / (jb.part == 3 ? 3 : 2);
      seg = ci / cp;
      ci -= seg * cp;
    }
    bool ok = co < jb.cout && ci < jb.cin_real;
    if (jb.unfold_w) {
      const int kw = ci / jb.cin_real;
      ci -= kw * jb.cin_real;
      ok = co < jb.cout && kw < K;
      t = (K - 1 - j) * K * K + khw * K + kw;
    }
    float v = 0.f;
    if (ok) {
      if (jb.transposed) {
        v = jb.w[((size_t)ci * jb.src_cin + jb.ci0 + co) * taps + (taps - 1 - t)];
        if (jb.in_scale) v *= jb.in_scale[jb.ci0 + co];
      } else {
        v = jb.w[((size_t)co * jb.cin_real + ci) * taps + t];
        if (jb.in_scale) v *= jb.in_scale[ci];
      }
    }
    if (jb.part == 2 || (jb.part == 3 && seg == 2)) v -= __bfloat162float(__float2bfloat16(v));   // low part of the 3 x bf16 split
    jb.dst[i] = __float2bfloat16(v);
  }
}

}  // namespace

int pack_conv_weights_batched(const PackJob* d_jobs, int njobs, cudaStream_t st) {
  if (njobs <= 0) return 0;
  pack_jobs_kernel<<<dim3(48, njobs), 256, 0, st>>>(d_jobs);
  FTB_LAUNCH_OK();
  return 0;
}

int pack_conv_weights(const float* w, int cout, int cin_real, int ksize, int cin_pad, int ntile_n,
                      int ntiles, const float* in_scale, bf16* dst, cudaStream_t st, bool unfold_w) {
  FTB_CHECK(cin_pad % 16 == 0 && ntile_n % 16 == 0, "pack: padded extents must be multiples of 16");
  FTB_CHECK(ntile_n * ntiles >= cout && cin_pad >= (unfold_w ? ksize : 1) * cin_real, "pack: tile does not cover the weight");
  const int taps = ksize * ksize * (unfold_w ? 1 : ksize);
  const size_t total = (size_t)ntiles * taps * cin_pad * ntile_n;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  pack_weights_kernel<<<blocks, 256, 0, st>>>(w, cout, cin_real, ksize, cin_pad, ntile_n, ntiles, in_scale, dst,
                                              unfold_w ? 1 : 0);
  FTB_LAUNCH_OK();
  return 0;
}

int conv_naive(const ConvSrc& s0, const ConvSrc& s1, const ConvWeights& w, const ConvEpilogue& e,
               Act& out, int out_cgoff, cudaStream_t st) {
  const Act& a0 = *s0.t;
  FTB_CHECK((s0.cg + s1.cg) * 8 == w.cin, "conv_naive: weight K extent does not match the sources");
  FTB_CHECK(!e.pre_out && e.drop_p == 0.f, "conv_naive: the training epilogue (pre-norm output, dropout) is tcgen05-only");
  NaiveParams p{};
  p.B = a0.B; p.D = a0.D; p.H = a0.H; p.W = a0.W; p.K = w.ksize; p.pad = (w.ksize - 1) / 2;
  p.Kw = w.kw(); p.padw = (p.Kw - 1) / 2;
  p.cg0 = s0.cg; p.cg1 = s1.t ? s1.cg : 0;
  p.s0_cgtot = a0.cg(); p.s0_cgoff = s0.cgoff;
  p.s1_cgtot = s1.t ? s1.t->cg() : 0; p.s1_cgoff = s1.cgoff;
  p.N = w.n; p.KS = w.cin / 16;
  p.src0 = a0.p; p.src1 = s1.t ? s1.t->p : nullptr;
  p.w_batch_stride = w.batch_stride;
  p.out = out.p; p.out_cgtot = out.cg();
  p.out_f32 = e.out_f32; p.out_f32_c = e.out_f32_c;
  p.norm = e.norm; p.mul = e.mul; p.add = e.add; p.mul_stride = e.mul_stride; p.add_stride = e.add_stride;
  p.resid = e.resid ? e.resid->p : nullptr;
  p.resid_cgtot = e.resid ? e.resid->cg() : 0; p.resid_cgoff = e.resid_cgoff;
  p.ss_in = e.prenorm ? e.prenorm_ss : nullptr; p.ss_out = e.sumsq_out;
  p.prenorm = e.prenorm; p.silu = e.silu; p.q_dh = e.q_dim_head; p.q_scale = e.q_scale;
  const size_t nthreads = (size_t)a0.B * a0.voxels();
  const int blocks = (int)((nthreads + 127) / 128);
  for (int nt = 0; nt < w.ntiles; ++nt) {
    NaiveParams q = p;
    q.wpack = w.w + (size_t)nt * w.tile_elems();
    q.bias = e.bias ? e.bias + (size_t)nt * w.n : nullptr;
    q.out_cgoff = out_cgoff + nt * (w.n / 8);
    q.qsoftmax = (e.q_softmax_heads && nt == 0) ? 1 : 0;
    conv_naive_kernel<<<blocks, 128, 0, st>>>(q);
    FTB_LAUNCH_OK();
  }
  return 0;
}

int conv_dispatch(const ConvSrc& s0, const ConvSrc& s1, const ConvWeights& w, const ConvEpilogue& e,
                  Act& out, int out_cgoff, cudaStream_t st) {
  static int naive = -1;
  if (naive < 0) {
    const char* v = getenv("FTB_CONV_IMPL");
    naive = (v && strcmp(v, "naive") == 0) ? 1 : 0;
  }
  return naive ? conv_naive(s0, s1, w, e, out, out_cgoff, st)
               : conv_igemm(s0, s1, w, e, out, out_cgoff, st);
}

}  // namespace ftb
