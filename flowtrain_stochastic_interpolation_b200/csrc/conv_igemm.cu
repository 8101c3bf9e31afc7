// 3-D convolution as implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces every nn.Conv3d on the Unet3D path (reference src/flowtrain/models/unet_attn_3d.py:
// stem :535, Block.proj :227, Upsample.conv :83, stage-last :614/:657, Downsample.conv :103,
// res_conv :263, to_qkv :303/:354, to_out :306/:355, final_conv :667) and fuses what follows it
// in the reference into the epilogue: bias, channel RMSNorm (:127-128), FiLM (:241), SiLU (:243),
// the residual add (:278, :695), the pre-attention RMSNorm (:311, as a per-row scale) and the
// q softmax of LinearAttention (:326,:329).
//
// Data flow per CTA (persistent, 384 threads = plane producer warp | two MMA issuer warps | weight producer warp |
// 8 epilogue warps):
//   * activations live in HBM as [B][C/8][D][H][W][8] bf16.  One TMA box = one halo PLANE
//     (BH x BW voxels x all input channels) and lands in shared memory as [cg][h][w][8], which is
//     exactly the no-swizzle K-major UMMA operand layout: 8 consecutive voxels along W form a
//     core matrix, LBO = channel-group pitch, SBO = halo-row pitch.  A filter tap (kd,kh,kw) is
//     therefore just a start-address shift of the A descriptor: no im2col, every plane is read
//     from L2 once and reused by all K^3 taps; zero padding comes from TMA out-of-bounds fill.
//   * an output tile is M = 128 voxels = 16 rows (H) x 8 columns (W) of one depth plane; the CTA
//     marches along D through a ring of plane slots, so each plane is loaded once per column.
//   * a group of NZ consecutive output planes is accumulated at once (NZ TMEM accumulators of N
//     columns, side by side).  Weights are packed per (kh, kw, k-step) as a (K*N) x 16 K-major
//     tile whose row blocks are the K depth taps in DESCENDING kd order: an input plane q of the
//     window feeds output planes q-kd, i.e. ADJACENT accumulators, so one tcgen05.mma with
//     N_mma = (#depth taps) * N covers them all ("depth-tap stacking").  A-operand reads from
//     shared memory (128 rows x 32 B per MMA whatever N is) are what bounds narrow layers
//     (Cout = 48: 4 KB of A for 1.5 KB of B); stacking divides them by up to K.
//   * weights stay resident in shared memory when everything fits, else they stream through a
//     small ring, once per group of NZ planes.
//   * accumulators are double-buffered (the epilogue of group g overlaps the MMAs of g+1); each
//     epilogue thread owns one voxel row, so the channel reduction of RMSNorm is thread-local.
//   * an MMA issuer is one elected lane of warp 1 / 2 (disjoint accumulators).  The whole warp builds a per-group
//     table of stacked MMAs; the lane decodes its entries ONCE per (group, channel chunk) pass into registers
//     (issue_cached) and then walks tap -> entry -> k-step with two adds per instruction: the issuing thread's own
//     instruction stream is what bounds the kernel, not the tensor pipe or shared memory.
#include <stdlib.h>

#include "ops.h"

namespace ftb {

namespace {

constexpr int kThreads = 384;   // TMA warp | up to 2 MMA issuer warps | spare | 8 epilogue warps
constexpr int kMaxIss = 2;
constexpr int kMaxN = 256;
#ifndef FTB_MAX_SLOTS
#define FTB_MAX_SLOTS 16   // plane ring cap: 16 instead of 12 lets the 7^3 stem prefetch more of the next pass (1.19 -> 1.13 ms, same-box A/B)
#endif
constexpr int kMaxSlots = FTB_MAX_SLOTS;
constexpr int kMaxWSlots = 32;
constexpr int kMaxEnt = 128;  // MMA table entries per group and issuer (overwrite table + stacked runs)
constexpr int kMaxKS = 32;    // k-steps (Cin_pad / 16)
constexpr int kMaxCC = 16;    // channel chunks (passes) per group

enum : int { F_SILU = 1, F_QSOFTMAX = 2, F_PLAIN = 8 };

struct IgemmParams {
  int B, D, H, W;
  int K, pad, taps;            // taps = K*Kw (kh, kw) positions; the K depth taps are stacked along N
  int Kw, padw;                // extent along W (= K, or 1 when the W taps were unfolded into channels)
  int cg0, cg1, KS;
  // Input channels are processed in chunks ("passes"): a ring slot holds ONE plane of ONE chunk (<= 64
  // channels), so the plane window stays small for any Cin.  ncc == 1: the window slides along D and
  // halo planes are kept between groups; ncc > 1: every (group, chunk) pass loads its own window.
  int ncc;
  int cc_src[kMaxCC], cc_cgoff[kMaxCC], cc_ks[kMaxCC], cc_ks0[kMaxCC];   // source, first cg, k-steps, first k-step
  int s0_cgtot, s0_cgoff, s1_cgtot, s1_cgoff;
  int N, smax;                 // smax: adjacent accumulators one MMA may cover (smax*N <= 256)
  int TH, BH, BW;
  int nHt, nWt, nSeg, LZ, NZ;
  int n_items;
  int nslot, wslot, w_resident;
  int flat, Dext;              // 1x1x1 convs: tiles are 128 CONSECUTIVE voxels (Dext tiles per sample), else Dext = D
  int n_iss;                   // MMA issuer warps (each owns a contiguous share of a group's accumulators)
  uint32_t cg_pitch, row_pitch, slot_stride, wtap_bytes;
  uint32_t wchunk_bytes;    // weight ring slot: one (tap, channel chunk) = max cc_ks * kstep_bytes
  uint32_t kstep_bytes;     // K * N * 32: one k-step of one (kh,kw) position, all depth taps
  uint32_t off_w, off_bar;
  uint32_t tmem_cols;
  const bf16* wpack;
  long long w_batch_stride;
  bf16* out;
  int out_cgtot, out_cgoff;
  float* out_f32;
  int out_f32_c;               // channels of this N tile that exist in the fp32 output
  int out_f32_cs;              // channels per sample of the fp32 output (sample stride)
  const float *bias, *mul, *add;
  int mul_stride, add_stride, norm;
  const bf16* resid;
  int resid_cgtot, resid_cgoff;
  const bf16* pre_src;
  const float* ss_in;          // per-voxel ||src0||^2 (replaces reading pre_src)
  float* ss_out;               // per-voxel sum of squares of the stored output
  long long* dbg;              // optional [grid][8] cycle counters of the first MMA issuer (FTB_CONV_DBG)
  int pre_cgtot, pre_cgoff, pre_cg;
  int flags, q_dh;
  float q_scale;
  bf16* u_out;                 // training: pre-norm output (same channel layout as `out`), or null
  float drop_p;                // training: dropout after the activation
  unsigned long long drop_key;
  int f32_accum;               // fp32 mode: out_f32 += instead of = (the hi/lo operand products of one conv)
  long long w_tile_stride;     // elements between the packed weights of consecutive N tiles (blockIdx.y = tile)
  int vdiv;                    // flat tiles: divisor of the voxel coordinate of the tensor map (32: wide rows, 1)
  int wait_test;               // per-tap slot wait also tests the next slot (FTB_CONV_WAIT_TEST=1; off by default)
  int no_fast27;               // FTB_CONV_NO_FAST27: always walk the table tap by tap (no register-cached entries)
};

struct ItemCoord {
  int b, d0, lz, h0, w0;
};

__device__ __forceinline__ ItemCoord decode_item(const IgemmParams& p, int item) {
  ItemCoord c;
  int wt = item % p.nWt;
  int r = item / p.nWt;
  int ht = r % p.nHt;
  r /= p.nHt;
  int seg = r % p.nSeg;
  c.b = r / p.nSeg;
  c.d0 = seg * p.LZ;
  c.lz = min(p.LZ, p.Dext - c.d0);
  c.h0 = ht * p.TH;
  c.w0 = wt * 8;
  return c;
}


// ------------------------------------------------------------------------------ epilogue helpers
struct EpiCtx {
  uint32_t bias, mul, add;         // shared-memory byte addresses of this half's copy for the current sample (LDS, not generic loads)
  bf16* out_b;                     // output base of sample b
  bf16* u_b;                       // pre-norm output base of sample b (or null)
  int b;                           // sample index (dropout mask key)
  const bf16* res_b;               // residual base of sample b (or null)
  float* f32_b;                    // NCDHW fp32 output base of sample b (or null)
  int f32_lim;                     // channels of this N tile that exist in the fp32 output
  size_t cgs;                      // voxels per channel-group plane (D*H*W)
  float ssq;                       // running sum of squares of this voxel's stored outputs
  bool valid;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// 16-byte load from a shared-memory byte address.  (Through a generic pointer the same load compiles to LD.E.128, which
// costs two shared-memory wavefronts instead of one: ncu r02 counted 93 % of the kernel's LSU shared wavefronts on the
// 36 per-row parameter loads of the epilogue, competing with the tensor core's operand fetches for the same banks.)
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// finish 16 consecutive channels [c0, c0+16) of one voxel: SiLU, + residual, store.
// `pre` (optional) holds the residual's two 16-byte groups, loaded by the caller ahead of time.
template <bool kTrain = false>
__device__ __forceinline__ void epi_store16(const IgemmParams& p, EpiCtx& ec, int c0, size_t vox,
                                            float (&v)[16], uint4 pre0 = uint4(), uint4 pre1 = uint4(),
                                            bool use_pre = false) {
  if (p.flags & F_SILU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = silu_fast(v[j]);
  }
  if (!ec.valid) return;
  if (kTrain && p.drop_p > 0.f) {   // nn.Dropout after the activation (training only)
    const size_t grp = ((size_t)ec.b * p.out_cgtot + p.out_cgoff + (c0 >> 3)) * ec.cgs + vox;
    const DropMask d0 = drop_mask8(p.drop_p, p.drop_key, grp), d1 = drop_mask8(p.drop_p, p.drop_key, grp + ec.cgs);
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] *= d0(j); v[8 + j] *= d1(j); }
  }
  if (ec.res_b) {
    uint4 u0, u1;
    if (use_pre) {
      u0 = pre0; u1 = pre1;
    } else {
      const bf16* rp = ec.res_b + ((size_t)(p.resid_cgoff + (c0 >> 3)) * ec.cgs + vox) * 8;
      u0 = __ldg(reinterpret_cast<const uint4*>(rp));
      u1 = __ldg(reinterpret_cast<const uint4*>(rp + ec.cgs * 8));
    }
    float f[8];
    unpack_bf16x8(u0, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += f[j];
    unpack_bf16x8(u1, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[8 + j] += f[j];
  }
  if (p.ss_out) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0 = fmaf(v[j], v[j], s0); s1 = fmaf(v[8 + j], v[8 + j], s1); }
    ec.ssq += s0 + s1;
  }
  if (ec.f32_b) {
    float* o = ec.f32_b + (size_t)c0 * ec.cgs + vox;
    if (kTrain && p.f32_accum) {   // all 16 loads in flight before the first store
      float old[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) old[j] = (c0 + j < ec.f32_lim) ? __ldcg(o + (size_t)j * ec.cgs) : 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] += old[j];
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < ec.f32_lim) o[(size_t)j * ec.cgs] = v[j];
    return;
  }
  bf16* dst = ec.out_b + ((size_t)(p.out_cgoff + (c0 >> 3)) * ec.cgs + vox) * 8;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = v[j];
  *reinterpret_cast<uint4*>(dst) = pack_bf16x8(f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = v[8 + j];
  *reinterpret_cast<uint4*>(dst + ec.cgs * 8) = pack_bf16x8(f);
}

// pull this voxel's residual row towards the SM before the accumulator is read
__device__ __forceinline__ void epi_prefetch_resid(const IgemmParams& p, const EpiCtx& ec, size_t vox) {
  if (!ec.res_b || !ec.valid) return;
  const bf16* rp = ec.res_b + ((size_t)p.resid_cgoff * ec.cgs + vox) * 8;
  for (int cg = 0; cg < (p.N >> 3); ++cg)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(rp + (size_t)cg * ec.cgs * 8));
}

// y = (acc * rs) * mul + add'   (no norm; bias already folded into add')
// NLD x 16 columns behind ONE tcgen05.wait (the rows of the 1^3 convs are bound by TMEM round trips, not arithmetic)
template <int NLD, bool kTrain>
__device__ __forceinline__ void epi_affine_chunk(const IgemmParams& p, EpiCtx& ec, uint32_t trow, int c0, size_t vox,
                                                 float rs) {
  uint32_t r[NLD][16];
#pragma unroll
  for (int i = 0; i < NLD; ++i) tmem_ld16(trow + c0 + i * 16, r[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    float v[16];
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      const float4 mu = lds_f4(ec.mul + 4u * (uint32_t)(c0 + i * 16 + j4));
      const float4 ad = lds_f4(ec.add + 4u * (uint32_t)(c0 + i * 16 + j4));
      v[j4 + 0] = fmaf(__uint_as_float(r[i][j4 + 0]) * rs, mu.x, ad.x);
      v[j4 + 1] = fmaf(__uint_as_float(r[i][j4 + 1]) * rs, mu.y, ad.y);
      v[j4 + 2] = fmaf(__uint_as_float(r[i][j4 + 2]) * rs, mu.z, ad.z);
      v[j4 + 3] = fmaf(__uint_as_float(r[i][j4 + 3]) * rs, mu.w, ad.w);
    }
    epi_store16<kTrain>(p, ec, c0 + i * 16, vox, v);
  }
}

template <bool kTrain>
__device__ __forceinline__ void epi_affine(const IgemmParams& p, EpiCtx& ec, uint32_t trow,
                                           size_t vox, float rs) {
  epi_prefetch_resid(p, ec, vox);
  int c0 = 0;
  for (; c0 + 48 <= p.N; c0 += 48) epi_affine_chunk<3, kTrain>(p, ec, trow, c0, vox, rs);
  if (c0 + 32 <= p.N) { epi_affine_chunk<2, kTrain>(p, ec, trow, c0, vox, rs); c0 += 32; }
  if (c0 + 16 <= p.N) epi_affine_chunk<1, kTrain>(p, ec, trow, c0, vox, rs);
}

// y = acc * rs: no bias, affine, activation or residual (k and v of to_qkv)
template <int NLD>
__device__ __forceinline__ void epi_plain_chunk(const IgemmParams& p, EpiCtx& ec, uint32_t trow, int c0, size_t vox,
                                                float rs) {
  uint32_t r[NLD][16];
#pragma unroll
  for (int i = 0; i < NLD; ++i) tmem_ld16(trow + c0 + i * 16, r[i]);
  tmem_ld_wait();
  if (!ec.valid) return;
  bf16* dst = ec.out_b + ((size_t)(p.out_cgoff + (c0 >> 3)) * ec.cgs + vox) * 8;
#pragma unroll
  for (int i = 0; i < NLD; ++i)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[i][hf * 8 + j]) * rs;
      *reinterpret_cast<uint4*>(dst + (size_t)(2 * i + hf) * ec.cgs * 8) = pack_bf16x8(f);
    }
}
__device__ __forceinline__ void epi_plain(const IgemmParams& p, EpiCtx& ec, uint32_t trow, size_t vox, float rs) {
  int c0 = 0;
  for (; c0 + 64 <= p.N; c0 += 64) epi_plain_chunk<4>(p, ec, trow, c0, vox, rs);
  if (c0 + 32 <= p.N) { epi_plain_chunk<2>(p, ec, trow, c0, vox, rs); c0 += 32; }
  if (c0 + 16 <= p.N) epi_plain_chunk<1>(p, ec, trow, c0, vox, rs);
}

// channel RMSNorm (unet_attn_3d.py:127-128) with the whole voxel row held in registers:
// v = acc*rs + bias; y = v / max(||v||, 1e-12) * mul + add
// residual row of voxel `vox` (all channel groups of this N tile) into registers
template <int NCG>
__device__ __forceinline__ void epi_load_resid(const IgemmParams& p, const EpiCtx& ec, size_t vox, uint4 (&pre)[NCG]) {
  const bf16* rp = ec.res_b + ((size_t)p.resid_cgoff * ec.cgs + vox) * 8;
#pragma unroll
  for (int i = 0; i < NCG; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(rp + (size_t)i * ec.cgs * 8));
}

template <int NCH, bool kTrain>
__device__ __forceinline__ void epi_norm_regs(const IgemmParams& p, EpiCtx& ec, uint32_t trow,
                                              size_t vox, float rs) {
  constexpr bool kPre = NCH <= 3;   // residual row preloaded into registers (else L1 prefetch)
  uint4 pre[kPre ? 2 * NCH : 2];
  const bool has_res = ec.res_b != nullptr && ec.valid;
  if (kPre) {
    if (has_res) {
      epi_load_resid<2 * NCH>(p, ec, vox, pre);
    }
  } else {
    epi_prefetch_resid(p, ec, vox);
  }
  uint32_t r[NCH][16];
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) tmem_ld16(trow + ch * 16, r[ch]);
  tmem_ld_wait();
  float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      const float4 bi = lds_f4(ec.bias + 4u * (uint32_t)(ch * 16 + j4));
      const float a0 = fmaf(__uint_as_float(r[ch][j4 + 0]), rs, bi.x);
      const float a1 = fmaf(__uint_as_float(r[ch][j4 + 1]), rs, bi.y);
      const float a2 = fmaf(__uint_as_float(r[ch][j4 + 2]), rs, bi.z);
      const float a3 = fmaf(__uint_as_float(r[ch][j4 + 3]), rs, bi.w);
      ss[0] = fmaf(a0, a0, ss[0]); ss[1] = fmaf(a1, a1, ss[1]);
      ss[2] = fmaf(a2, a2, ss[2]); ss[3] = fmaf(a3, a3, ss[3]);
      r[ch][j4 + 0] = __float_as_uint(a0); r[ch][j4 + 1] = __float_as_uint(a1);
      r[ch][j4 + 2] = __float_as_uint(a2); r[ch][j4 + 3] = __float_as_uint(a3);
    }
  const float rinv = 1.f / fmaxf(sqrtf((ss[0] + ss[1]) + (ss[2] + ss[3])), 1e-12f);
  if (kTrain && ec.u_b && ec.valid) {   // training: the pre-norm row
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[ch][hf * 8 + j]);
        *reinterpret_cast<uint4*>(ec.u_b + ((size_t)(p.out_cgoff + ch * 2 + hf) * ec.cgs + vox) * 8) = pack_bf16x8(f);
      }
  }
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    float v[16];
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      const float4 mu = lds_f4(ec.mul + 4u * (uint32_t)(ch * 16 + j4));
      const float4 ad = lds_f4(ec.add + 4u * (uint32_t)(ch * 16 + j4));
      v[j4 + 0] = fmaf(__uint_as_float(r[ch][j4 + 0]) * rinv, mu.x, ad.x);
      v[j4 + 1] = fmaf(__uint_as_float(r[ch][j4 + 1]) * rinv, mu.y, ad.y);
      v[j4 + 2] = fmaf(__uint_as_float(r[ch][j4 + 2]) * rinv, mu.z, ad.z);
      v[j4 + 3] = fmaf(__uint_as_float(r[ch][j4 + 3]) * rinv, mu.w, ad.w);
    }
    if (kPre) epi_store16<kTrain>(p, ec, ch * 16, vox, v, pre[2 * ch], pre[2 * ch + 1], true);
    else epi_store16<kTrain>(p, ec, ch * 16, vox, v);
  }
}

// same, any N: one TMEM pass for the norm, a second for the output
// pass 1 of the two-pass norm: sum of squares of NLD x 16 columns behind one tcgen05.wait
template <int NLD>
__device__ __forceinline__ void epi_norm_ss_chunk(const EpiCtx& ec, uint32_t trow, int c0, float rs, float (&ss)[4]) {
  uint32_t r[NLD][16];
#pragma unroll
  for (int i = 0; i < NLD; ++i) tmem_ld16(trow + c0 + i * 16, r[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NLD; ++i)
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      const float4 bi = lds_f4(ec.bias + 4u * (uint32_t)(c0 + i * 16 + j4));
      const float a0 = fmaf(__uint_as_float(r[i][j4 + 0]), rs, bi.x);
      const float a1 = fmaf(__uint_as_float(r[i][j4 + 1]), rs, bi.y);
      const float a2 = fmaf(__uint_as_float(r[i][j4 + 2]), rs, bi.z);
      const float a3 = fmaf(__uint_as_float(r[i][j4 + 3]), rs, bi.w);
      ss[0] = fmaf(a0, a0, ss[0]); ss[1] = fmaf(a1, a1, ss[1]);
      ss[2] = fmaf(a2, a2, ss[2]); ss[3] = fmaf(a3, a3, ss[3]);
    }
}
// pass 2: normalise, affine, finish NLD x 16 columns behind one tcgen05.wait
template <int NLD, bool kTrain>
__device__ __forceinline__ void epi_norm_out_chunk(const IgemmParams& p, EpiCtx& ec, uint32_t trow, int c0, size_t vox,
                                                   float rs, float rinv) {
  uint32_t r[NLD][16];
#pragma unroll
  for (int i = 0; i < NLD; ++i) tmem_ld16(trow + c0 + i * 16, r[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    const int cc = c0 + i * 16;
    float v[16];
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      const float4 bi = lds_f4(ec.bias + 4u * (uint32_t)(cc + j4));
      const float4 mu = lds_f4(ec.mul + 4u * (uint32_t)(cc + j4));
      const float4 ad = lds_f4(ec.add + 4u * (uint32_t)(cc + j4));
      const float a0 = fmaf(__uint_as_float(r[i][j4 + 0]), rs, bi.x), a1 = fmaf(__uint_as_float(r[i][j4 + 1]), rs, bi.y);
      const float a2 = fmaf(__uint_as_float(r[i][j4 + 2]), rs, bi.z), a3 = fmaf(__uint_as_float(r[i][j4 + 3]), rs, bi.w);
      r[i][j4 + 0] = __float_as_uint(a0); r[i][j4 + 1] = __float_as_uint(a1);
      r[i][j4 + 2] = __float_as_uint(a2); r[i][j4 + 3] = __float_as_uint(a3);
      v[j4 + 0] = fmaf(a0 * rinv, mu.x, ad.x);
      v[j4 + 1] = fmaf(a1 * rinv, mu.y, ad.y);
      v[j4 + 2] = fmaf(a2 * rinv, mu.z, ad.z);
      v[j4 + 3] = fmaf(a3 * rinv, mu.w, ad.w);
    }
    if (kTrain && ec.u_b && ec.valid) {   // training: the pre-norm row
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[i][hf * 8 + j]);
        *reinterpret_cast<uint4*>(ec.u_b + ((size_t)(p.out_cgoff + (cc >> 3) + hf) * ec.cgs + vox) * 8) = pack_bf16x8(f);
      }
    }
    epi_store16<kTrain>(p, ec, cc, vox, v);
  }
}

// same, any N: one TMEM pass for the norm, a second for the output (64 / 48 columns per tcgen05.wait)
template <bool kTrain>
__device__ __forceinline__ void epi_norm_2pass(const IgemmParams& p, EpiCtx& ec, uint32_t trow,
                                               size_t vox, float rs) {
  epi_prefetch_resid(p, ec, vox);
  float ss[4] = {0.f, 0.f, 0.f, 0.f};
  int c0 = 0;
  for (; c0 + 64 <= p.N; c0 += 64) epi_norm_ss_chunk<4>(ec, trow, c0, rs, ss);
  if (c0 + 32 <= p.N) { epi_norm_ss_chunk<2>(ec, trow, c0, rs, ss); c0 += 32; }
  if (c0 + 16 <= p.N) epi_norm_ss_chunk<1>(ec, trow, c0, rs, ss);
  const float rinv = 1.f / fmaxf(sqrtf((ss[0] + ss[1]) + (ss[2] + ss[3])), 1e-12f);
  c0 = 0;
  for (; c0 + 48 <= p.N; c0 += 48) epi_norm_out_chunk<3, kTrain>(p, ec, trow, c0, vox, rs, rinv);
  if (c0 + 32 <= p.N) { epi_norm_out_chunk<2, kTrain>(p, ec, trow, c0, vox, rs, rinv); c0 += 32; }
  if (c0 + 16 <= p.N) epi_norm_out_chunk<1, kTrain>(p, ec, trow, c0, vox, rs, rinv);
}

// LinearAttention q: softmax over each dim_head group of this voxel, times dim_head^-0.5
// (unet_attn_3d.py:326,:329); no bias / affine on this path
template <int DH>
__device__ __forceinline__ void epi_qsoftmax(const IgemmParams& p, const EpiCtx& ec, uint32_t trow,
                                             size_t vox, float rs) {
  for (int hd = 0; hd < p.N / DH; ++hd) {
    uint32_t r[DH / 16][16];
#pragma unroll
    for (int ch = 0; ch < DH / 16; ++ch) tmem_ld16(trow + hd * DH + ch * 16, r[ch]);
    tmem_ld_wait();
    float mx = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < DH / 16; ++ch)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float v = __uint_as_float(r[ch][j]) * rs;
        r[ch][j] = __float_as_uint(v);
        mx = fmaxf(mx, v);
      }
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < DH / 16; ++ch)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e = __expf(__uint_as_float(r[ch][j]) - mx);
        r[ch][j] = __float_as_uint(e);
        sum += e;
      }
    const float inv = __fdividef(p.q_scale, sum);
    if (ec.valid) {
#pragma unroll
      for (int ch = 0; ch < DH / 16; ++ch)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[ch][hf * 8 + j]) * inv;
          const int cg = p.out_cgoff + ((hd * DH + ch * 16) >> 3) + hf;
          *reinterpret_cast<uint4*>(ec.out_b + ((size_t)cg * ec.cgs + vox) * 8) = pack_bf16x8(f);
        }
    }
  }
}


struct IssueCtx {
  uint32_t a_hi, b_hi, planes_enc, slot_enc, nslot, kinc, kstep;
  uint32_t slot_w0, acc0;   // per group: ring slot of window plane 0, TMEM address of accumulator 0
};

// Issue the MMAs of one weight chunk: table entries [ea, ea_end) x NKS k-steps.  An entry is
// (window plane q, B row-block offset | accumulate flag, TMEM column offset, instruction
// descriptor); `overwrite_ok` = this is the first chunk, honour the entry's flag.
template <int NKS>
__device__ __forceinline__ void issue_entries(const IssueCtx& ic, uint32_t ea, uint32_t ea_end, uint32_t aoff,
                                              uint32_t wb, bool overwrite_ok) {
  if (ea >= ea_end) return;
  uint32_t ex, ey, ez, ew;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ex), "=r"(ey), "=r"(ez), "=r"(ew) : "r"(ea));
  const uint32_t a_base = ic.planes_enc + aoff;
#pragma unroll 1
  while (true) {
    ea += 16;
    const bool more = ea < ea_end;
    uint32_t nx = 0, ny = 0, nz = 0, nw = 0;
    if (more)   // next entry's load overlaps this entry's issue
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(nx), "=r"(ny), "=r"(nz), "=r"(nw) : "r"(ea));
    uint32_t slot = ic.slot_w0 + ex;            // window plane q -> ring slot
    if (slot >= ic.nslot) slot -= ic.nslot;
    const uint32_t a = a_base + slot * ic.slot_enc;
    const uint32_t b = (ey & 0x7FFFFFFFu) + wb;
    const uint32_t d = ic.acc0 + ez;
    umma_bf16_lohi(d, a, ic.a_hi, b, ic.b_hi, ew, overwrite_ok ? (ey >> 31) : 1u);
#pragma unroll
    for (int i = 1; i < NKS; ++i) umma_bf16_lohi(d, a + i * ic.kinc, ic.a_hi, b + i * ic.kstep, ic.b_hi, ew, 1u);
    if (!more) break;
    ex = nx; ey = ny; ez = nz; ew = nw;
  }
}

// The table entries of one (group, channel-chunk) pass decoded ONCE into registers, then k-steps [KS0, NKS) of every
// entry for one filter tap: two adds per MMA.  The issuer is a single thread and its instruction stream, not the
// tensor pipe, bounded the kernel (FTB_CONV_DBG + the ncu source page, r02: ~300 instructions at ~6 cycles each per
// tap for 12 MMAs per issuer = 1 900 cycles per tap against 1 200 cycles of tensor time for both issuers together;
// the table walk re-read and re-decoded every entry for every tap).
constexpr int kCache = 8;
template <int NKS, int KS0>
__device__ __forceinline__ void issue_cached(const IssueCtx& ic, const uint32_t (&ca)[kCache], const uint32_t (&cb)[kCache],
                                             const uint32_t (&cd)[kCache], const uint32_t (&ci)[kCache], int ne,
                                             uint32_t aoff, uint32_t wb) {
#pragma unroll
  for (int e = 0; e < kCache; ++e) {
    if (e < ne) {
#pragma unroll
      for (int ks = KS0; ks < NKS; ++ks)
        umma_bf16_lohi(cd[e], ca[e] + aoff + (uint32_t)ks * ic.kinc, ic.a_hi, cb[e] + wb + (uint32_t)ks * ic.kstep, ic.b_hi,
                       ci[e], 1u);
    }
  }
}

// All filter taps of one (group, channel chunk) pass for cached entries: per tap only the weight-slot hand-shake
// (wait / commit) and one call of issue_cached remain on the issuing thread.
struct TapCtx {
  uint32_t w_enc, wchunk_enc, rowp_enc, first_lo, first_hi;
  uint64_t *w_full, *w_empty;
  int K, Kw, nwslot, res_slot0;
  bool stream_w, w_waited, first_pass, wait_test;
};
template <int NKS>
__device__ __forceinline__ void run_taps_cached(const IssueCtx& ic, const TapCtx& tc, const uint32_t (&ca)[kCache],
                                                const uint32_t (&cb)[kCache], const uint32_t (&cd)[kCache],
                                                const uint32_t (&ci)[kCache], int ne, uint32_t& wslot, uint32_t& wphase,
                                                bool& w_ready) {
  uint32_t aoff_row = 0;
  uint32_t res_slot = (uint32_t)tc.res_slot0;
  for (int kh = 0; kh < tc.K; ++kh, aoff_row += tc.rowp_enc)
    for (int kw = 0; kw < tc.Kw; ++kw) {
      uint32_t slot;
      if (tc.stream_w) {
        slot = wslot;
        if (!w_ready) {   // also tests the next slot: when that is already full, the next tap skips its wait
          uint32_t ns = slot + 1, np = wphase;
          if (ns == (uint32_t)tc.nwslot) { ns = 0; np ^= 1; }
          w_ready = mbar_wait_test_next(&tc.w_full[slot], wphase, &tc.w_full[ns], np, tc.wait_test);
        } else {
          w_ready = false;
        }
        tc_fence_after();
      } else {
        slot = res_slot++;
        if (!tc.w_waited) {
          mbar_wait(&tc.w_full[slot], 0);
          tc_fence_after();
        }
      }
      const uint32_t wb = tc.w_enc + slot * tc.wchunk_enc;
      const uint32_t ao = aoff_row + (uint32_t)kw;
      if (tc.first_pass && (kh | kw) == 0) {   // first touch of every accumulator: k-step 0 through the overwrite table
        issue_entries<1>(ic, tc.first_lo, tc.first_hi, ao, wb, true);
        issue_cached<NKS, 1>(ic, ca, cb, cd, ci, ne, ao, wb);
      } else {
        issue_cached<NKS, 0>(ic, ca, cb, cd, ci, ne, ao, wb);
      }
      if (tc.stream_w) {
        umma_commit(&tc.w_empty[slot]);
        if (++wslot == (uint32_t)tc.nwslot) { wslot = 0; wphase ^= 1; }
      }
    }
}

template <bool kTrain>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                  const IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* s_planes = smem;
  uint8_t* s_w = smem + p.off_w;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* plane_full = bars;
  uint64_t* plane_empty = bars + kMaxSlots;
  uint64_t* w_full = bars + 2 * kMaxSlots;
  uint64_t* w_empty = w_full + kMaxWSlots;
  uint64_t* acc_full = w_empty + kMaxWSlots;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint4* s_tab = reinterpret_cast<uint4*>(tmem_ptr + 4);             // [kMaxEnt] MMA table of the group
  float* s_par = reinterpret_cast<float*>(s_tab + kMaxIss * kMaxEnt);          // [2 halves][bias|mul|add][kMaxN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nslot; ++i) {
      mbar_init(&plane_full[i], 1);
      mbar_init(&plane_empty[i], p.n_iss);
    }
    for (int i = 0; i < p.wslot; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], p.n_iss);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], p.n_iss);
      mbar_init(&acc_empty[i], 8);
    }
    fence_barrier_init();
    prefetch_tmap(&tm0);
    if (p.cg1 > 0) prefetch_tmap(&tm1);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Programmatic dependent launch: the prologue above (barrier init, TMEM allocation, descriptor prefetch) ran while
  // the previous kernel of the stream was still draining; its results are only read below this point.  The
  // trigger lets the NEXT kernel's CTAs take an SM as soon as this kernel's CTA there exits.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================================================================== TMA producer: activation planes
    // (The weight chunks have their own producer thread below.  With both in one thread's program order the plane
    // loads of the next group queued behind 27 weight-chunk loads that each wait for a free ring slot, so the planes
    // could not be prefetched: FTB_CONV_DBG showed the issuers waiting 10 % of their cycles for planes and 13 % for
    // weights on 48 -> 48 @64^3.)
    if (lane == 0) {
      uint32_t pslot = 0, pphase = 0;   // plane ring cursor
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        const int npl = c.lz + 2 * p.pad;
        const int ngroups = (c.lz + p.NZ - 1) / p.NZ;
        int issued = 0;   // ncc == 1: planes of this item loaded so far
        for (int g = 0; g < ngroups; ++g) {
          const int nze = min(p.NZ, c.lz - g * p.NZ);
          const int win = nze + 2 * p.pad;
          for (int cc = 0; cc < p.ncc; ++cc) {
            const bool s1 = p.cc_src[cc] != 0;
            const CUtensorMap* tm = s1 ? &tm1 : &tm0;
            const int cgc = c.b * (s1 ? p.s1_cgtot : p.s0_cgtot) + (s1 ? p.s1_cgoff : p.s0_cgoff) + p.cc_cgoff[cc];
            const uint32_t bytes = (uint32_t)p.cc_ks[cc] * 2u * p.cg_pitch;
            int q_lo, q_hi;   // item-relative planes to load for this pass
            if (p.ncc == 1) { q_lo = issued; q_hi = min(npl, g * p.NZ + win); issued = q_hi; }
            else { q_lo = g * p.NZ; q_hi = q_lo + win; }
            for (int q = q_lo; q < q_hi; ++q) {
              mbar_wait_backoff(&plane_empty[pslot], pphase ^ 1);
              mbar_expect_tx(&plane_full[pslot], bytes);
              uint8_t* dst = s_planes + (size_t)pslot * p.slot_stride;
              const int dz = c.d0 - p.pad + q;
              if (p.flat) tma_load_3d(dst, tm, &plane_full[pslot], 0, dz * 128 / p.vdiv, cgc);
              else tma_load_4d(dst, tm, &plane_full[pslot], (c.w0 - p.padw) * 8, c.h0 - p.pad, dz, cgc);
              if (++pslot == (uint32_t)p.nslot) { pslot = 0; pphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ===================================================================== bulk-copy producer: weight chunks
    if (lane == 0) {
      uint32_t wslot = 0, wphase = 0;   // weight ring cursor
      bool w_loaded = false;
      for (int item = blockIdx.x; item < p.n_items && !(p.w_resident && w_loaded); item += gridDim.x) {
        const ItemCoord c = decode_item(p, item);
        const int ngroups = (c.lz + p.NZ - 1) / p.NZ;
        for (int g = 0; g < ngroups && !(p.w_resident && w_loaded); ++g) {
          for (int cc = 0; cc < p.ncc; ++cc) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack + (long long)blockIdx.y * p.w_tile_stride +
                                                                 (long long)c.b * p.w_batch_stride) +
                                  (size_t)p.cc_ks0[cc] * p.kstep_bytes;
            const uint32_t wbytes = (uint32_t)p.cc_ks[cc] * p.kstep_bytes;
            for (int t = 0; t < p.taps; ++t) {
              uint32_t slot;
              if (p.w_resident) {
                slot = (uint32_t)(cc * p.taps + t);
              } else {
                slot = wslot;
                mbar_wait_backoff(&w_empty[slot], wphase ^ 1);
                if (++wslot == (uint32_t)p.wslot) { wslot = 0; wphase ^= 1; }
              }
              mbar_expect_tx(&w_full[slot], wbytes);
              bulk_load(s_w + (size_t)slot * p.wchunk_bytes, wsrc + (size_t)t * p.wtap_bytes, wbytes, &w_full[slot]);
            }
          }
          w_loaded = true;
        }
      }
    }
  } else if (warp >= 1 && warp <= p.n_iss) {
    // ===================================================================== MMA issuers
    // Window plane q of a group feeds the accumulators zi = q - kd for the depth taps
    // kd in [kd_lo, kd_hi]: adjacent TMEM column blocks, and adjacent row blocks j = K-1-kd of
    // the packed weight tile, so one MMA of N_mma = ns*N covers ns of them.  The warp builds a
    // small table per group ("first" entries: one per (q, kd) pair, used on the very first weight
    // chunk where each accumulator's first touch must overwrite; "main" entries: stacked runs
    // of up to smax taps), then ONE elected lane walks chunk -> entry -> k-step with two adds
    // per instruction.  A single thread cannot issue narrow MMAs (N = 48: 24 tensor cycles each)
    // fast enough, so up to two issuer warps run side by side on DISJOINT accumulators: issuer i
    // owns the output planes [i*nze/n_iss, (i+1)*nze/n_iss) of every group (tcgen05.mma from
    // different threads are unordered, which is harmless when they never share an accumulator).
    const int iss = warp - 1;
    uint4* tab = s_tab + iss * kMaxEnt;
    const bool leader = elect_one();
    IssueCtx ic;
    ic.a_hi = ((p.row_pitch >> 4) & 0x3FFFu) | (1u << 14);                 // SBO | descriptor version
    ic.b_hi = (256u >> 4) | (1u << 14);
    ic.planes_enc = (smem_u32(s_planes) >> 4) | (((p.cg_pitch >> 4) & 0x3FFFu) << 16);   // | LBO
    ic.slot_enc = p.slot_stride >> 4;
    ic.nslot = (uint32_t)p.nslot;
    ic.kinc = (2 * p.cg_pitch) >> 4;                                       // one k-step = 2 channel groups
    ic.kstep = p.kstep_bytes >> 4;
    const uint32_t w_enc = (smem_u32(s_w) >> 4) | ((128u >> 4) << 16);
    const uint32_t wchunk_enc = p.wchunk_bytes >> 4;
    const uint32_t rowp_enc = p.row_pitch >> 4;
    const uint32_t idesc0 = umma_idesc_bf16_f32(128, 0);
    const uint32_t idesc_n = (uint32_t)(p.N >> 3) << 17;
    const uint32_t nb_enc = (uint32_t)(p.N * 2);                           // one depth tap of B rows, >>4
    const uint32_t tab_addr = smem_u32(tab);
    const bool stream_w = !p.w_resident;
    // ring cursors, advanced incrementally (no divisions on the issue path)
    uint32_t slot_w0 = 0;                  // ring slot of the current window's plane 0
    uint32_t rslot = 0, rphase = 0;        // next plane_full barrier to wait for
    uint32_t wslot = 0, wphase = 0;        // next streamed weight chunk
    bool w_ready = false;                  // the slot at the cursor was already seen full (tested with the previous wait)
    uint32_t gctr = 0;
    int tab_nze = -1, n_first = 0, n_main = 0;
    bool w_waited = false;
    long long dbg_acc = 0, dbg_plane = 0, dbg_w = 0, dbg_nchunk = 0, dbg_issue = 0, dbg_tab = 0;
    const long long dbg_t0 = clock64();
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      const int npl = c.lz + 2 * p.pad;
      const int ngroups = (c.lz + p.NZ - 1) / p.NZ;
      int waited = 0;   // ncc == 1: planes of this item already waited for
      for (int g = 0; g < ngroups; ++g, ++gctr) {
        const int nze = min(p.NZ, c.lz - g * p.NZ);
        const int win = nze + 2 * p.pad;
        const uint32_t ab = gctr & 1;
        if (nze != tab_nze) {
          // ---- (re)build the table for this window shape: lane q owns window plane q.  Entries
          // are position independent: (q, B row-block offset | accumulate flag, TMEM column offset
          // of the first accumulator, instruction descriptor).
          long long tt0 = 0;
          if (p.dbg) tt0 = clock64();
          __syncwarp();
          const int z_lo = (iss * nze) / p.n_iss, z_hi = ((iss + 1) * nze) / p.n_iss;
          int cnt = 0, kd_hi = 0, nruns = 0;
          // "first" table (k-step 0 of the group's first weight chunk, where every accumulator is touched for the first
          // time and must be overwritten): the issuer's accumulators [z_lo, z_hi) are cut into blocks of c = min(smax, K)
          // and block [za, za + n) is overwritten by ONE stacked MMA from window plane za + n - 1 (depth taps n-1 .. 0);
          // all other (plane, tap) pairs follow as accumulating stacked runs.  (Before: one N-wide MMA per pair on all
          // k-steps of the first chunk - 12 x 44 instead of 372 tensor cycles per k-step for 48 -> 48, NZ 4.)
          int cnt_ow = 0, rem = 0, nrem = 0;
          if (lane < win && z_hi > z_lo) {
            const int kd_lo = max(0, lane - (z_hi - 1));
            kd_hi = min(2 * p.pad, lane - z_lo);
            cnt = max(0, kd_hi - kd_lo + 1);
            nruns = (cnt + p.smax - 1) / p.smax;
            const int cblk = min(p.smax, p.K);
            if (cnt > 0 && kd_lo == 0 && (((lane - z_lo + 1) % cblk) == 0 || lane == z_hi - 1)) cnt_ow = ((lane - z_lo) % cblk) + 1;
            rem = cnt - cnt_ow;
            nrem = (rem + p.smax - 1) / p.smax;
          }
          int so = cnt_ow > 0 ? 1 : 0, sr = nrem, sm = nruns;  // inclusive scans
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int to = __shfl_up_sync(0xffffffffu, so, o), tr = __shfl_up_sync(0xffffffffu, sr, o),
                      tm = __shfl_up_sync(0xffffffffu, sm, o);
            if (lane >= o) { so += to; sr += tr; sm += tm; }
          }
          const int n_ow = __shfl_sync(0xffffffffu, so, 31);
          n_first = n_ow + __shfl_sync(0xffffffffu, sr, 31);
          n_main = __shfl_sync(0xffffffffu, sm, 31);
          if (cnt_ow > 0) {
            const int kd = cnt_ow - 1;
            tab[so - 1] = make_uint4((uint32_t)lane, (uint32_t)(p.K - 1 - kd) * nb_enc, (uint32_t)((lane - kd) * p.N),
                                     idesc0 + (uint32_t)cnt_ow * idesc_n);
          }
          for (int r = 0; r < nrem; ++r) {
            const int kd = kd_hi - r * p.smax;
            const int ns = min(p.smax, rem - r * p.smax);
            tab[n_ow + sr - nrem + r] = make_uint4((uint32_t)lane, (uint32_t)(p.K - 1 - kd) * nb_enc | 0x80000000u,
                                                   (uint32_t)((lane - kd) * p.N), idesc0 + (uint32_t)ns * idesc_n);
          }
          for (int r = 0; r < nruns; ++r) {
            const int kd = kd_hi - r * p.smax;
            const int ns = min(p.smax, cnt - r * p.smax);
            tab[n_first + sm - nruns + r] = make_uint4((uint32_t)lane, (uint32_t)(p.K - 1 - kd) * nb_enc | 0x80000000u,
                                                       (uint32_t)((lane - kd) * p.N), idesc0 + (uint32_t)ns * idesc_n);
          }
          tab_nze = nze;
          __syncwarp();
          if (p.dbg) dbg_tab += clock64() - tt0;
        }
        if (leader) {
          long long tq0 = 0;
          if (p.dbg) tq0 = clock64();
          // (the accumulator-buffer barrier rides in the first plane batch: one round trip for the whole group start)
          ic.acc0 = tmem_base + ab * p.NZ * p.N;
          for (int cc = 0; cc < p.ncc; ++cc) {
            if (p.dbg) tq0 = clock64();
            int n_new;
            if (p.ncc == 1) { const int upto = min(npl, g * p.NZ + win); n_new = upto - waited; waited = upto; }
            else n_new = win;
            int extra = cc == 0 ? 1 : 0;   // acc_empty still to be waited for
            for (int i = 0; i < n_new || extra; ) {   // up to six barriers per round trip
              uint32_t ba[6], bp[6];
              int j = 0;
              if (extra) { ba[0] = smem_u32(&acc_empty[ab]); bp[0] = ((gctr >> 1) & 1) ^ 1; j = 1; extra = 0; }
#pragma unroll
              for (int k = 0; k < 6; ++k) {
                if (k >= j) {
                  if (i < n_new) {
                    ba[k] = smem_u32(&plane_full[rslot]); bp[k] = rphase;
                    if (++rslot == ic.nslot) { rslot = 0; rphase ^= 1; }
                    ++i;
                  } else {
                    ba[k] = ba[k > 0 ? k - 1 : 0]; bp[k] = bp[k > 0 ? k - 1 : 0];
                  }
                }
              }
              mbar_wait6(ba, bp);
            }
            if (p.dbg) dbg_plane += clock64() - tq0;
            tc_fence_after();
            ic.slot_w0 = slot_w0;
            const int nks = p.cc_ks[cc];
            // main-table entries of this pass in registers (window plane -> ring slot resolved here, once)
            const bool cached = n_main <= kCache && nks <= 4 && !p.no_fast27;
            uint32_t ca[kCache], cb[kCache], cd[kCache], ci[kCache];
            if (cached) {
#pragma unroll
              for (int e = 0; e < kCache; ++e) {
                ca[e] = 0; cb[e] = 0; cd[e] = 0; ci[e] = 0;
                if (e < n_main) {
                  uint32_t ex, ey, ez, ew;
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ex), "=r"(ey), "=r"(ez), "=r"(ew)
                               : "r"(tab_addr + (uint32_t)(n_first + e) * 16u));
                  uint32_t slot = slot_w0 + ex;
                  if (slot >= ic.nslot) slot -= ic.nslot;
                  ca[e] = ic.planes_enc + slot * ic.slot_enc;
                  cb[e] = ey & 0x7FFFFFFFu;
                  cd[e] = ic.acc0 + ez;
                  ci[e] = ew;
                }
              }
            }
            if (cached) {
              long long ti0 = 0;
              if (p.dbg) ti0 = clock64();
              TapCtx tcx;
              tcx.w_enc = w_enc; tcx.wchunk_enc = wchunk_enc; tcx.rowp_enc = rowp_enc;
              tcx.first_lo = tab_addr; tcx.first_hi = tab_addr + (uint32_t)n_first * 16u;
              tcx.w_full = w_full; tcx.w_empty = w_empty;
              tcx.K = p.K; tcx.Kw = p.Kw; tcx.nwslot = p.wslot; tcx.res_slot0 = cc * p.taps;
              tcx.stream_w = stream_w; tcx.w_waited = w_waited; tcx.first_pass = cc == 0; tcx.wait_test = p.wait_test != 0;
              switch (nks) {
                case 1: run_taps_cached<1>(ic, tcx, ca, cb, cd, ci, n_main, wslot, wphase, w_ready); break;
                case 2: run_taps_cached<2>(ic, tcx, ca, cb, cd, ci, n_main, wslot, wphase, w_ready); break;
                case 3: run_taps_cached<3>(ic, tcx, ca, cb, cd, ci, n_main, wslot, wphase, w_ready); break;
                default: run_taps_cached<4>(ic, tcx, ca, cb, cd, ci, n_main, wslot, wphase, w_ready); break;
              }
              if (p.dbg) dbg_issue += clock64() - ti0;
            } else {
            int t = 0;
            for (int kh = 0; kh < p.K; ++kh)
              for (int kw = 0; kw < p.Kw; ++kw, ++t) {
                const uint32_t aoff = kh * rowp_enc + kw;
                uint32_t wslot_i;
                if (stream_w) {
                  wslot_i = wslot;
                  long long tq1 = 0;
                  if (p.dbg) tq1 = clock64();
                  mbar_wait(&w_full[wslot_i], wphase);
                  if (p.dbg) { dbg_w += clock64() - tq1; ++dbg_nchunk; }
                  tc_fence_after();
                } else {
                  wslot_i = (uint32_t)(cc * p.taps + t);
                  if (!w_waited) {
                    mbar_wait(&w_full[wslot_i], 0);
                    tc_fence_after();
                  }
                }
                const uint32_t wb = w_enc + wslot_i * wchunk_enc;
                const bool firstc = (cc | t) == 0;
                const uint32_t ea_m = tab_addr + (uint32_t)n_first * 16u, ea_m_end = ea_m + (uint32_t)n_main * 16u;
                long long ti0 = 0;
                if (p.dbg) ti0 = clock64();
                int nk = nks;
                uint32_t ao = aoff, wo = wb;
                if (firstc) {   // k-step 0 through the overwrite table, the rest of the chunk like any other
                  issue_entries<1>(ic, tab_addr, tab_addr + (uint32_t)n_first * 16u, ao, wo, true);
                  --nk; ao += ic.kinc; wo += ic.kstep;
                }
                switch (nk) {   // straight-line k-step issue for the common chunk lengths
                  case 0: break;
                  case 1: issue_entries<1>(ic, ea_m, ea_m_end, ao, wo, false); break;
                  case 2: issue_entries<2>(ic, ea_m, ea_m_end, ao, wo, false); break;
                  case 3: issue_entries<3>(ic, ea_m, ea_m_end, ao, wo, false); break;
                  case 4: issue_entries<4>(ic, ea_m, ea_m_end, ao, wo, false); break;
                  default:
                    for (int i = 0; i < nk; ++i)
                      issue_entries<1>(ic, ea_m, ea_m_end, ao + i * ic.kinc, wo + i * ic.kstep, false);
                }
                if (p.dbg) dbg_issue += clock64() - ti0;
                if (stream_w) {
                  umma_commit(&w_empty[wslot_i]);
                  if (++wslot == (uint32_t)p.wslot) { wslot = 0; wphase ^= 1; }
                  w_ready = false;
                }
              }
            }   // !cached
            if (p.ncc > 1) {   // this pass's window is done: hand all its slots back
              for (int i = 0; i < win; ++i) {
                umma_commit(&plane_empty[slot_w0]);
                if (++slot_w0 == ic.nslot) slot_w0 = 0;
              }
            }
          }
          if (p.ncc == 1) {
            // planes that leave the sliding window: the NZ oldest, or everything at the end of the item
            const int nrel = (g == ngroups - 1) ? win : nze;
            for (int i = 0; i < nrel; ++i) {
              umma_commit(&plane_empty[slot_w0]);
              if (++slot_w0 == ic.nslot) slot_w0 = 0;
            }
          }
          umma_commit(&acc_full[ab]);
        }
        w_waited = true;
      }
    }
    if (p.dbg && leader && iss == 0) {
      long long* o = p.dbg + (size_t)blockIdx.x * 16;
      o[0] = clock64() - dbg_t0; o[1] = dbg_acc; o[2] = dbg_plane; o[3] = dbg_w; o[4] = dbg_nchunk; o[5] = dbg_issue; o[6] = dbg_tab;
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================================================================== epilogue (8 warps)
    // Two halves of four warps; a warp may only touch the TMEM lane quadrant (warp % 4), so the
    // halves split a group's planes between them (alternating) and both see every accumulator.
    // Thread m of a quadrant set owns voxel row m of the tile: the channel reduction of the
    // RMSNorm is thread-local.  Per-channel parameters of the current sample sit in shared
    // memory (one copy per half), so the inner loops are LDS.128 + FMA only.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int tid_h = (((warp - 4) & 3) << 5) | lane;  // 0..127 within the half
    const int m = q * 32 + lane;
    const int hl = m >> 3, wl = m & 7;
    const size_t plane_vox = (size_t)p.H * p.W;
    const size_t cgs = (size_t)p.D * plane_vox;
    float* par = s_par + half * 3 * kMaxN;
    EpiCtx ec;
    ec.bias = smem_u32(par); ec.mul = smem_u32(par + kMaxN); ec.add = smem_u32(par + 2 * kMaxN);
    ec.cgs = cgs;
    const int tile_c0 = (int)blockIdx.y * p.N;
    uint32_t gctr = 0;
    int cur_b = -1;
    long long e_wait = 0, e_t0 = clock64();
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const ItemCoord c = decode_item(p, item);
      const int ngroups = (c.lz + p.NZ - 1) / p.NZ;
      const int h = c.h0 + hl, w = c.w0 + wl;
      ec.valid = p.flat || ((hl < p.TH) && (h < p.H) && (w < p.W));
      if (c.b != cur_b) {
        // (bias, mul, add) of sample b; without a norm the bias folds into the affine part
        named_bar_sync(1 + half, 128);
        for (int ch = tid_h; ch < p.N; ch += 128) {
          const int cht = tile_c0 + ch;   // channel within the whole layer (blockIdx.y = N tile)
          float bi = p.bias ? __ldg(p.bias + cht) : 0.f;
          const float mu = p.mul ? __ldg(p.mul + (size_t)c.b * p.mul_stride + cht) : 1.f;
          float ad = p.add ? __ldg(p.add + (size_t)c.b * p.add_stride + cht) : 0.f;
          if (!p.norm) { ad = fmaf(bi, mu, ad); bi = 0.f; }
          par[ch] = bi; par[kMaxN + ch] = mu; par[2 * kMaxN + ch] = ad;
        }
        named_bar_sync(1 + half, 128);
        cur_b = c.b;
      }
      // the N tile's channel-group offset is folded into the per-sample base pointers
      const size_t tile_cg = (size_t)(tile_c0 >> 3) * cgs * 8;
      ec.out_b = p.out + (size_t)c.b * p.out_cgtot * cgs * 8 + tile_cg;
      ec.u_b = p.u_out ? p.u_out + (size_t)c.b * p.out_cgtot * cgs * 8 + tile_cg : nullptr;
      ec.b = c.b;
      ec.res_b = p.resid ? p.resid + (size_t)c.b * p.resid_cgtot * cgs * 8 + tile_cg : nullptr;
      ec.f32_b = p.out_f32 ? p.out_f32 + ((size_t)c.b * p.out_f32_cs + tile_c0) * cgs : nullptr;
      ec.f32_lim = p.out_f32_c - tile_c0;
      for (int g = 0; g < ngroups; ++g, ++gctr) {
        const int nze = min(p.NZ, c.lz - g * p.NZ);
        const uint32_t ab = gctr & 1;
        long long ew0 = 0;
        if (p.dbg) ew0 = clock64();
        mbar_wait_backoff(&acc_full[ab], (gctr >> 1) & 1);
        if (p.dbg) e_wait += clock64() - ew0;
        tc_fence_after();
        for (int zi = (half + g) & 1; zi < nze; zi += 2) {
          const int d = c.d0 + g * p.NZ + zi;
          // voxel index within one (b, cg): flat tiles are 128 consecutive voxels (the last may be ragged)
          const size_t vox = p.flat ? (size_t)d * 128 + m : (size_t)d * plane_vox + (size_t)h * p.W + w;
          if (p.flat) ec.valid = vox < cgs;
          // The residual row is consumed at the very end of a row's epilogue, yet its HBM latency used to be the longest
          // stall of the whole epilogue (~4 000 of 9 000 cycles per row, FTB_CONV_DBG): pull the rows of the NEXT plane
          // this half will finish (in this group or the next) into L2 now, one plane ahead, at no register cost.
          if (ec.res_b) {
            long long nv = -1;
            const size_t vstep = p.flat ? 128 : plane_vox;
            if (zi + 2 < nze) nv = (long long)(vox + 2 * vstep);
            else if (g + 1 < ngroups) {
              const int z2 = (half + g + 1) & 1;
              if (z2 < min(p.NZ, c.lz - (g + 1) * p.NZ)) nv = (long long)(vox + (size_t)(p.NZ - zi + z2) * vstep);
            }
            if (nv >= 0 && (p.flat ? (size_t)nv < cgs : ec.valid)) {
              const bf16* rp = ec.res_b + ((size_t)p.resid_cgoff * cgs + (size_t)nv) * 8;
              for (int cg = 0; cg < (p.N >> 3); ++cg)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)cg * cgs * 8));
            }
          }
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (ab * p.NZ + zi) * p.N;
          float rs = 1.f;
          if (p.ss_in) {
            // fused pre-attention RMSNorm: 1 / max(||x||_2, 1e-12); ||x||^2 came from x's producer
            if (ec.valid) rs = 1.f / fmaxf(sqrtf(__ldg(p.ss_in + (size_t)c.b * cgs + vox)), 1e-12f);
          } else if (p.pre_src && ec.valid) {
            const bf16* src = p.pre_src + (((size_t)c.b * p.pre_cgtot + p.pre_cgoff) * cgs + vox) * 8;
            float ss0 = 0.f, ss1 = 0.f;
            for (int cgi = 0; cgi < p.pre_cg; cgi += 2) {
              float f[8], g8[8];
              const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(src + (size_t)cgi * cgs * 8));
              const uint4 u1 = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(cgi + 1) * cgs * 8));
              unpack_bf16x8(u0, f);
              unpack_bf16x8(u1, g8);
#pragma unroll
              for (int j = 0; j < 8; ++j) { ss0 = fmaf(f[j], f[j], ss0); ss1 = fmaf(g8[j], g8[j], ss1); }
            }
            rs = 1.f / fmaxf(sqrtf(ss0 + ss1), 1e-12f);
          }
          ec.ssq = 0.f;
          if ((p.flags & F_QSOFTMAX) && blockIdx.y == 0) {   // q softmax on the first N tile (q | k | v)
            if (p.q_dh == 32) epi_qsoftmax<32>(p, ec, trow, vox, rs);
            else epi_qsoftmax<16>(p, ec, trow, vox, rs);
          } else if (p.norm) {
            if (p.N == 48) epi_norm_regs<3, kTrain>(p, ec, trow, vox, rs);
            else epi_norm_2pass<kTrain>(p, ec, trow, vox, rs);
          } else if (p.flags & F_PLAIN) {
            epi_plain(p, ec, trow, vox, rs);
          } else {
            epi_affine<kTrain>(p, ec, trow, vox, rs);
          }
          if (p.ss_out && ec.valid) p.ss_out[(size_t)c.b * cgs + vox] = ec.ssq;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ab]);
      }
    }
    if (p.dbg && warp == 4 && lane == 0) {
      long long* o = p.dbg + (size_t)blockIdx.x * 16;
      o[8] = clock64() - e_t0; o[9] = e_wait;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 4-D view of a blocked activation: (W*8, H, D, B*CG) bf16; box = (BW*8, BH, 1, cg)
int make_plane_tmap(CUtensorMap* tm, const Act& a, int BW, int BH, int cg) {
  PFN_encodeTiled enc = get_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.H, (cuuint64_t)a.D,
                        (cuuint64_t)a.B * a.cg()};
  cuuint64_t gstr[3] = {(cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16,
                        (cuuint64_t)a.W * a.H * a.D * 16};
  cuuint32_t box[4] = {(cuuint32_t)BW * 8, (cuuint32_t)BH, 1u, (cuuint32_t)cg};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return 0;
}

long long* g_conv_dbg = nullptr;
constexpr uint32_t kSmemLimit = 227 * 1024 - 128;  // 227 KB opt-in maximum minus the alignment pad

}  // namespace

// debug: per-CTA cycle counters of the last instrumented launch (FTB_CONV_DBG=1): [grid][8]
int conv_debug_read(long long* host, int n) {
  if (!g_conv_dbg) return -1;
  return cudaMemcpy(host, g_conv_dbg, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

int make_voxel_tmap(CUtensorMap* tm, const Act& a, int box_vox, int box_cg, int* vdiv) {
  PFN_encodeTiled enc = get_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t vox = (cuuint64_t)a.voxels();
  // TMA works in units of the box's innermost dimension.  Narrow form: (8 channels, voxels, B*CG), innermost = 16 bytes.
  // Wide form (whenever the voxel count and the box are multiples of 32): (32 voxels x 8 channels, voxels / 32, B*CG),
  // innermost = 512 contiguous bytes; the box lands in shared memory byte for byte like the narrow one
  // ([cg][voxel][8]).  *vdiv = divisor of the voxel coordinate (32 or 1).
  static const bool no_wide = getenv("FTB_TMA_NARROW") != nullptr;
  if (vdiv != nullptr && !no_wide && vox % 32 == 0 && box_vox % 32 == 0) {
    *vdiv = 32;
    cuuint64_t gdim[3] = {256, vox / 32, (cuuint64_t)a.B * a.cg()};
    cuuint64_t gstr[2] = {512, vox * 16};
    cuuint32_t box[3] = {256u, (cuuint32_t)(box_vox / 32), (cuuint32_t)box_cg};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a.p, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wide voxel view) failed (" + std::to_string((int)r) + ")");
    return 0;
  }
  if (vdiv != nullptr) *vdiv = 1;
  cuuint64_t gdim[3] = {8, vox, (cuuint64_t)a.B * a.cg()};
  cuuint64_t gstr[2] = {16, vox * 16};
  cuuint32_t box[3] = {8u, (cuuint32_t)box_vox, (cuuint32_t)box_cg};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, a.p, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (voxel view) failed (" + std::to_string((int)r) + ")");
  return 0;
}

// number of channel chunks (passes) one source of `cg` channel groups is cut into for a K^3 conv with N-wide tiles
// (same rule as the planner below: equal chunks of an even divisor <= 8 groups (16 for K = 1) whose weight chunk of
// one (kh, kw) position stays <= 56 KB); the planner accepts at most kMaxCC chunks over both sources
int conv_chunk_count(int cg, int K, int N) {
  const int cg_cap = K == 1 ? 16 : 8;
  const uint32_t kstep_bytes = (uint32_t)K * N * 32;
  for (int d = cg < cg_cap ? cg : cg_cap; d > 2; d -= 2)
    if (cg % d == 0 && (uint32_t)(d / 2) * kstep_bytes <= 56u * 1024u) return cg / d;
  return cg / 2;
}
int conv_max_chunks() { return kMaxCC; }

int conv_igemm(const ConvSrc& s0, const ConvSrc& s1, const ConvWeights& w, const ConvEpilogue& e,
               Act& out, int out_cgoff, cudaStream_t st) {
  const Act& a0 = *s0.t;
  FTB_CHECK(s0.cg > 0 && (s0.cg % 2) == 0, "conv: src0 channel groups must be a positive multiple of 2");
  FTB_CHECK(s1.cg % 2 == 0, "conv: src1 channel groups must be a multiple of 2");
  FTB_CHECK((s0.cg + s1.cg) * 8 == w.cin, "conv: weight K extent does not match the sources");
  FTB_CHECK(w.n % 16 == 0 && w.n >= 16 && w.n <= 256, "conv: N tile must be a multiple of 16 in [16,256]");
  FTB_CHECK(w.ksize == 1 || w.ksize == 3 || w.ksize == 5 || w.ksize == 7, "conv: ksize");
  FTB_CHECK(out.B == a0.B && out.D == a0.D && out.H == a0.H && out.W == a0.W, "conv: out dims");
  if (s1.t) FTB_CHECK(s1.t->B == a0.B && s1.t->D == a0.D && s1.t->H == a0.H && s1.t->W == a0.W, "conv: src1 dims");
  FTB_CHECK(!e.q_softmax_heads || ((e.q_dim_head == 16 || e.q_dim_head == 32) && w.n % e.q_dim_head == 0),
            "conv: q softmax needs dim_head 16 or 32 dividing N");

  IgemmParams p{};
  p.B = a0.B; p.D = a0.D; p.H = a0.H; p.W = a0.W;
  p.K = w.ksize; p.pad = (w.ksize - 1) / 2;
  p.Kw = w.kw(); p.padw = (p.Kw - 1) / 2; p.taps = p.K * p.Kw;
  FTB_CHECK(p.Kw == p.K || p.Kw == 1, "conv: ksize_w must be ksize or 1");
  p.cg0 = s0.cg; p.cg1 = s1.t ? s1.cg : 0;
  p.KS = (p.cg0 + p.cg1) / 2;
  FTB_CHECK(p.KS <= kMaxKS, "conv: more than 512 input channels");
  p.s0_cgtot = a0.cg(); p.s0_cgoff = s0.cgoff;
  p.s1_cgtot = s1.t ? s1.t->cg() : 0; p.s1_cgoff = s1.cgoff;
  p.N = w.n;
  p.smax = 256 / p.N < 1 ? 1 : 256 / p.N;
  p.BW = 8 + 2 * p.padw;
  // 1x1x1 convs have no halo, so their tiles are free-form: 128 consecutive voxels give 2 KB
  // contiguous runs per channel group in HBM (16 x 8 bricks: 128 B runs) for loads and stores
  p.flat = (p.K == 1 && getenv("FTB_NOFLAT") == nullptr) ? 1 : 0;
  p.Dext = p.flat ? (int)((a0.voxels() + 127) / 128) : a0.D;
  p.nWt = p.flat ? 1 : cdiv(a0.W, 8);
  p.row_pitch = p.BW * 16;
  p.kstep_bytes = (uint32_t)p.K * p.N * 32;
  p.wtap_bytes = (uint32_t)p.KS * p.kstep_bytes;
  // ---- channel chunks: each source is cut into equal chunks of `ccg` channel groups (an even
  // divisor of its group count, <= 8 groups = 64 channels for K > 1, whose weight chunk of one
  // (kh,kw) position stays <= 56 KB).  One source of <= 8 groups is a single chunk: the classic
  // sliding window.
  int cg_cap = p.K == 1 ? 16 : 8;
  const uint32_t kWChunkMax = 56 * 1024;
  auto pick = [&](int cg) {
    for (int d = cg < cg_cap ? cg : cg_cap; d > 2; d -= 2)
      if (cg % d == 0 && (uint32_t)(d / 2) * p.kstep_bytes <= kWChunkMax) return d;
    return 2;
  };
  int max_ccg = 0, nchunks = 0;
  size_t all_w = 0;
  auto build_chunks = [&]() -> int {
    p.ncc = 0;
    max_ccg = 0;
    const int cgs[2] = {p.cg0, p.cg1};
    int ks0 = 0;
    for (int sidx = 0; sidx < 2; ++sidx) {
      if (cgs[sidx] == 0) continue;
      const int ccg = pick(cgs[sidx]);
      for (int off = 0; off < cgs[sidx]; off += ccg) {
        FTB_CHECK(p.ncc < kMaxCC, "conv: too many channel chunks (channel count with no even divisor <= 64/8?)");
        p.cc_src[p.ncc] = sidx; p.cc_cgoff[p.ncc] = off; p.cc_ks[p.ncc] = ccg / 2; p.cc_ks0[p.ncc] = ks0;
        ks0 += ccg / 2;
        ++p.ncc;
      }
      if (ccg > max_ccg) max_ccg = ccg;
    }
    p.wchunk_bytes = (uint32_t)(max_ccg / 2) * p.kstep_bytes;
    nchunks = p.taps * p.ncc;
    all_w = (size_t)nchunks * p.wchunk_bytes;   // resident layout: one ring-sized slot per (chunk, tap)
    return 0;
  };
  const int sms = num_sms();
  const uint32_t bar_bytes =
      (2 * kMaxSlots + 2 * kMaxWSlots + 4) * 8 + 16 + kMaxIss * kMaxEnt * 16 + 2 * 3 * kMaxN * 4;

  // ---- tile height, planes per group (NZ) and ring sizing against the 227 KB shared-memory
  // budget.  The tile is 8 (W) x TH (H) voxels in the 128-row MMA; TH = 16 unless a plane window
  // of that height does not fit (very wide inputs), then rows are traded for capacity.  NZ output
  // planes share one pass over the weights and one TMEM buffer (2 buffers x NZ x N <= 512
  // columns); larger NZ means more depth-tap stacking per MMA and fewer weight passes.
  uint32_t slack = 0, fixed = 0;
  size_t w_region = 0;
  bool fits = false;
  int nz_cap = 512 / (2 * p.N);
  if (nz_cap > 8) nz_cap = 8;
  if (nz_cap > p.Dext) nz_cap = p.Dext;
  {
    // 1x1x1 convs: up to 4 flat 128-voxel tiles per accumulator group (fewer producer/issuer/epilogue hand-offs per
    // byte moved; 2 -> 4 took the 1^3 launches of a Heun step from 2.35 to 2.16 ms)
    static const int k1cap = getenv("FTB_K1_NZ") ? atoi(getenv("FTB_K1_NZ")) : 4;
    if (p.K == 1 && nz_cap > k1cap) nz_cap = k1cap < 1 ? 1 : k1cap;
  }
  if (const char* env = getenv("FTB_NZ")) {   // planner override for experiments
    const int v = atoi(env);
    if (v >= 1 && v < nz_cap) nz_cap = v;
  }
  while (nz_cap > 1 && (nz_cap * p.K > 48 || nz_cap + 2 * p.pad > 32)) --nz_cap;   // tables: <= 8 + 2 * 48 entries
  FTB_CHECK(nz_cap >= 1, "conv: N tile too wide for a double-buffered accumulator");
  // Smaller channel chunks are tried before a shorter tile: a wide kernel (K = 7: window of NZ + 6
  // planes) only fits with 32- or 16-channel chunks, and a full-height tile keeps all 128 MMA rows busy.
  const int th0 = (a0.H >= 16 || p.flat) ? 16 : a0.H;
  for (int attempt = 0; attempt < 4 && !fits; ++attempt) {
  if (attempt > 0 && attempt < 3) { if (cg_cap <= 2) continue; cg_cap = cg_cap > 4 ? 4 : 2; }
  FTB_TRY(build_chunks());
  for (int th = th0; th >= (attempt < 3 ? th0 : 1) && !fits; th = th / 2) {   // last attempt: trade tile rows for capacity
    p.TH = th;
    p.BH = p.TH + 2 * p.pad;
    p.nHt = p.flat ? 1 : cdiv(a0.H, p.TH);
    if (p.flat) p.row_pitch = 128;
    p.cg_pitch = p.BH * p.row_pitch;
    p.slot_stride = (uint32_t)round_up(max_ccg * (int)p.cg_pitch, 128);   // one plane of one channel chunk
    slack = (uint32_t)(16 - p.TH + 2) * p.row_pitch + 512;  // A rows of a partial tile over-read
    fixed = slack + bar_bytes + 256;
    // Candidates (NZ, weight ring): score = shared-memory A reads per output plane
    // ((NZ + halo) / NZ, the stacking efficiency) + the share of a group's planes that cannot be
    // prefetched while the previous group computes (ring slots beyond the window), divided by
    // the fill of the last group of a typical depth segment.  Measured on 48->48 @64^3:
    // (NZ 4, ring 3) 300 us < (5, 4) 347 us < (3, 4) 421 us — the score orders them the same way.
    double best_score = 1e30;
    const int dseg = p.Dext < 16 ? p.Dext : 16;
    int ws_hi = nchunks == 1 ? 2 : (p.wchunk_bytes <= 8192 ? 6 : (p.wchunk_bytes <= 16384 ? 4 : (p.wchunk_bytes <= 32768 ? 3 : 2)));
    int ws_lo = 2;
    if (const char* env = getenv("FTB_WSLOT")) {
      const int v = atoi(env);
      if (v >= 2 && v <= kMaxWSlots) ws_hi = ws_lo = v;
    }
    for (int nz = nz_cap; nz >= 1; --nz) {
      const int win = nz + 2 * p.pad;
      const double fill = (double)dseg / (cdiv(dseg, nz) * nz);
      for (int ws = ws_hi + 1; ws >= ws_lo; --ws) {   // ws_hi + 1 stands for "resident"
        const bool resident = ws == ws_hi + 1;
        if (resident && !(w.batch_stride == 0 && nchunks <= kMaxWSlots)) continue;
        const size_t wbytes = resident ? all_w : (size_t)ws * p.wchunk_bytes;
        if (wbytes + (size_t)(win + 1) * p.slot_stride + fixed > kSmemLimit) continue;
        int ns = (int)((kSmemLimit - fixed - wbytes) / p.slot_stride);
        ns = ns > kMaxSlots ? kMaxSlots : ns;
        const int spare = ns - win;
        // planes that must be fetched per group (sliding window: NZ new ones; chunked passes: a
        // whole window per pass) and cannot be prefetched into spare slots
        const int fetch = p.ncc == 1 ? nz : win;
        double score = (double)win / nz + 0.35 * (spare >= fetch ? 0.0 : (double)(fetch - spare) / fetch);
        score = score / fill - (resident ? 0.1 : 0.0) + (ws == 2 && !resident && nchunks > 1 ? 0.05 : 0.0);
        if (score < best_score - 1e-9) {
          best_score = score;
          p.NZ = nz;
          p.w_resident = resident ? 1 : 0;
          p.wslot = resident ? nchunks : ws;
          fits = true;
        }
      }
    }
    if (th == 1) break;
  }
  }
  FTB_CHECK(fits, "conv: one plane window + weights exceed shared memory (Cin too large)");
  w_region = (size_t)p.wslot * p.wchunk_bytes;
  const int cols = a0.B * p.nHt * p.nWt;
  // Tiny volumes (4^3, 8^3 per sample): with NZ planes per group there are fewer work items than half the SMs and each
  // CTA runs a long serial MMA chain; trade depth-tap stacking for parallelism until the launch fills half the GPU.
  if (getenv("FTB_NO_NZ_SPREAD") == nullptr)
  {
    static const int spread_pct = getenv("FTB_SPREAD_PCT") ? atoi(getenv("FTB_SPREAD_PCT")) : 50;
    while (p.NZ > 1 && (long long)cols * cdiv(p.Dext, p.NZ) * w.ntiles * 100 < (long long)sms * spread_pct) --p.NZ;
  }
  int nslot = (int)((kSmemLimit - fixed - w_region) / p.slot_stride);
  nslot = nslot > kMaxSlots ? kMaxSlots : nslot;
  p.nslot = nslot;
  // segment length along D: minimise (rounds of items over the SMs) x (planes per item, halo
  // planes counted at half weight: they cost loads but no MMAs)
  int lz = p.Dext;
  {
    double best = 1e30;
    for (int cand = p.Dext; cand >= p.NZ; cand = cdiv(cand, 2)) {
      const int cl = round_up(cand, p.NZ) > p.Dext ? p.Dext : round_up(cand, p.NZ);
      const double cost = (double)cdiv(cols * cdiv(p.Dext, cl), sms) * (cl + p.pad);
      if (cost < best - 1e-9) { best = cost; lz = cl; }
      if (cand == 1) break;
    }
  }
  p.LZ = lz;
  p.nSeg = cdiv(p.Dext, lz);
  p.n_items = cols * p.nSeg;
  p.off_w = (uint32_t)round_up((int)(p.nslot * p.slot_stride + slack), 128);
  p.off_bar = (uint32_t)round_up((int)(p.off_w + w_region), 16);
  const uint32_t smem_bytes = p.off_bar + bar_bytes + 128;
  FTB_CHECK(smem_bytes <= kSmemLimit + 128, "conv: smem budget");
  uint32_t cols_needed = 2u * p.NZ * p.N;
  uint32_t tc = 32;
  while (tc < cols_needed) tc <<= 1;
  FTB_CHECK(tc <= 512, "conv: accumulators exceed TMEM");
  p.tmem_cols = tc;

  p.n_iss = p.NZ >= 2 ? kMaxIss : 1;
  if (const char* env = getenv("FTB_NISS")) { const int v = atoi(env); if (v >= 1 && v <= kMaxIss && v <= p.NZ) p.n_iss = v; }
  p.dbg = nullptr;
  static const int no_fast27 = getenv("FTB_CONV_NO_FAST27") ? 1 : 0;
  p.no_fast27 = no_fast27;
  static const int wt_env = getenv("FTB_CONV_WAIT_TEST") ? 1 : 0;   // measured 2.7 % slower here (profiles/r02_wait_test_next.log)
  p.wait_test = wt_env;
  if (getenv("FTB_CONV_DBG")) {
    static long long* dbuf = nullptr;
    if (!dbuf) FTB_CUDA(cudaMalloc(&dbuf, 256 * 16 * sizeof(long long)));
    FTB_CUDA(cudaMemsetAsync(dbuf, 0, 256 * 16 * sizeof(long long), st));
    p.dbg = dbuf;
    g_conv_dbg = dbuf;
  }
  p.wpack = w.w;
  p.w_batch_stride = w.batch_stride;
  p.out = out.p; p.out_cgtot = out.cg();
  p.out_f32 = e.out_f32; p.out_f32_c = e.out_f32_c; p.out_f32_cs = e.out_f32_c;
  p.norm = e.norm ? 1 : 0;
  p.mul = e.mul; p.add = e.add; p.mul_stride = e.mul_stride; p.add_stride = e.add_stride;
  p.resid = e.resid ? e.resid->p : nullptr;
  p.resid_cgtot = e.resid ? e.resid->cg() : 0; p.resid_cgoff = e.resid_cgoff;
  p.ss_in = e.prenorm ? e.prenorm_ss : nullptr;
  p.ss_out = e.sumsq_out;
  if (e.prenorm) { p.pre_src = a0.p; p.pre_cgtot = a0.cg(); p.pre_cgoff = s0.cgoff; p.pre_cg = s0.cg; }
  p.q_dh = e.q_dim_head; p.q_scale = e.q_scale;
  p.u_out = nullptr;
  if (e.pre_out) {
    FTB_CHECK(e.norm && w.ntiles == 1, "conv: the pre-norm side output needs the norm epilogue and a single N tile");
    FTB_CHECK(e.pre_out->C == out.C && e.pre_out->voxels() == out.voxels() && e.pre_out->B == out.B,
              "conv: pre-norm output must be shaped like the output");
    p.u_out = e.pre_out->p;
  }
  p.drop_p = e.drop_p; p.drop_key = e.drop_key;
  p.f32_accum = (e.out_f32 && e.out_f32_accum) ? 1 : 0;

  CUtensorMap tm0, tm1;
  const int box0 = pick(p.cg0), box1 = s1.t ? pick(p.cg1) : 0;   // TMA box = one channel chunk
  if (p.flat) {
    int vd0 = 1, vd1 = 1;
    FTB_TRY(make_voxel_tmap(&tm0, a0, 128, box0, &vd0));
    if (s1.t) FTB_TRY(make_voxel_tmap(&tm1, *s1.t, 128, box1, &vd1)); else { tm1 = tm0; vd1 = vd0; }
    FTB_CHECK(vd0 == vd1, "conv: the two sources of a 1x1x1 conv must have the same voxel count");
    p.vdiv = vd0;
  } else {
    FTB_TRY(make_plane_tmap(&tm0, a0, p.BW, p.BH, box0));
    if (s1.t) FTB_TRY(make_plane_tmap(&tm1, *s1.t, p.BW, p.BH, box1)); else tm1 = tm0;
  }

  static bool attr_set = false;
  if (!attr_set) {
    FTB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kSmemLimit + 128)));
    FTB_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kSmemLimit + 128)));
    attr_set = true;
  }
  const int grid = p.n_items < sms ? p.n_items : sms;
  if (getenv("FTB_CONV_PLAN"))
    fprintf(stderr, "conv plan: K%d cin %d(+%d) N%d @%dx%dx%d B%d -> ncc %d NZ %d iss %d wslot %d%s nslot %d LZ %d TH %d items %d smem %u\n",
            p.K, p.cg0 * 8, p.cg1 * 8, p.N, p.D, p.H, p.W, p.B, p.ncc, p.NZ, p.n_iss, p.wslot,
            p.w_resident ? "(resident)" : "", p.nslot, p.LZ, p.TH, p.n_items, smem_bytes);
  // every N tile of the layer in ONE launch (blockIdx.y = tile): a wide layer on a small volume (192 channels at
  // 4^3: 8-32 items) is bound by each CTA streaming the whole weight tensor through its shared memory, so splitting
  // N over more CTAs divides both that traffic and the MMA width per CTA
  FTB_CHECK(w.ntiles == 1 || (!e.norm && !e.sumsq_out), "conv: channel norm / sum-of-squares epilogues need a single N tile");
  IgemmParams q = p;
  q.wpack = w.w;
  q.w_tile_stride = (long long)w.tile_elems();
  q.bias = e.bias;
  q.out_cgoff = out_cgoff;
  q.flags = (e.silu ? F_SILU : 0) | (e.q_softmax_heads ? F_QSOFTMAX : 0);
  if (!(q.flags & F_QSOFTMAX) && !q.norm && !q.bias && !q.mul && !q.add && !e.silu && !q.resid && !q.out_f32 && !q.ss_out)
    q.flags |= F_PLAIN;
  int prof = -1;
  if (prof_enabled()) {
    // algorithmic work of this launch: real (unpadded) channel counts
    const double vox = (double)a0.B * a0.voxels();
    const double cin = w.cin_real > 0 ? w.cin_real : w.cin;
    const double cout = w.cout_real > 0 ? w.cout_real : (double)w.n * w.ntiles;
    const double flops = 2.0 * vox * cin * cout * p.K * p.K * (w.ksize_w == 1 && p.K > 1 ? 1 : p.K);   // cin_real of an unfolded conv already counts the W taps
    const double bytes = vox * (cin + cout) * 2.0;  // read input once, write output once (bf16)
    prof = prof_begin(st, flops, bytes, w.ksize > 1 ? 0 : 1);
  }
  // the training epilogue (pre-norm side output, dropout) is a separate instantiation: the inference kernel
  // carries none of its code
  {
    static const bool pdl = getenv("FTB_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, w.ntiles); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    if (q.u_out || q.drop_p > 0.f || q.f32_accum) FTB_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true>, tm0, tm1, q));
    else FTB_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false>, tm0, tm1, q));
  }
  prof_end(prof, st);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
