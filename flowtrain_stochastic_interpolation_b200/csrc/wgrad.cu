// Weight gradient of a 3-D convolution on tcgen05 tensor cores (sm_100a):
//     dW[co][ci][kd][kh][kw] = sum over (b, voxel v) of dY[b][co][v] * X[b][ci][v + tap]
// i.e. autograd's conv3d weight gradient for every nn.Conv3d of the Unet3D path (reference
// src/flowtrain/models/unet_attn_3d.py: Block.proj :227, res_conv :263, stem :535, ...), used by the
// training step (project/geodata-3d-unconditional/model_train_inference.py:417-457).
//
// The contraction runs over VOXELS, and both operands sit in HBM as [B][C/8][D][H][W][8] bf16:
// for a fixed channel group, 8 consecutive W voxels x 8 channels are 128 contiguous bytes with the
// channels innermost.  That is the no-swizzle MN-major UMMA core matrix (8 K-elements x 16 B), so
// a TMA box of dY lands in shared memory as the A operand (M = output channels) and a halo box of X
// as the B operand (N = input channels), with a filter tap (kh, kw) a pure start-address shift of
// the B descriptor: no im2col, no transposes.
//   * the X box is requested with dimension order (W, channel group, H): in shared memory the
//     row pitch is then ncg * (channel-group pitch), so the kh taps of a 3^3 conv become 3*ncg
//     uniformly strided 8-channel groups and ONE MMA with N = K*Cin_chunk covers all kh
//     (an M=128 MMA costs max(N/2, 32 + N/4) cycles: N = 48 alone wastes 45 % on the fixed part).
//   * M is always 128: rows beyond Cout read whatever follows in shared memory and produce
//     accumulator rows nobody reads (an M=128 MMA costs the same as M=64).
//   * accumulators stay in TMEM for the CTA's whole life (split-K over voxel tiles); a CTA owns one
//     "class" = (kd, up to 512/N (kh-group, kw) accumulators, one Cin chunk, one 128-row block of
//     Cout) and a slice of the (b, d, tile) items; at the end 4 warps add the rows into the fp32
//     gradient with atomics.
//   warp 0: TMA producer | warp 1: MMA issuer | warps 2-5: TMEM -> global (atomicAdd)
#include <stdlib.h>

#include "ops.h"

namespace ftb {

namespace {

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 8;
constexpr uint32_t kWgSmemLimit = 227 * 1024 - 128;

struct WgradParams {
  int B, D, H, W;
  int K, Kw, pad, padw;
  int stack;            // kh taps stacked along N (1 or K)
  int ncg;              // channel groups per X chunk
  int Nacc;             // columns of one accumulator = stack * ncg * 8
  int nacc;             // accumulators per class
  int ngrp;             // accumulator groups per kd = (K / stack) * Kw
  int ncls;             // classes per kd
  int nchunk, nmb, nsplit, per_batch;
  int kdp, nkg;         // depth taps per CTA (2: the dY tiles of kd0 and kd0+1 share the M = 128 rows), kd groups
  uint32_t du_box_bytes;
  int nHt, nWt;
  long long items;      // items per sample (per_batch) or over the whole batch
  int x_cgtot, x_cgoff, du_cgtot, du_cgoff;
  int nstage;
  uint32_t du_bytes, x_bytes;
  uint32_t b_lbo, b_sbo, b_kh, b_kstep;   // X operand: K-direction / N-direction core-matrix strides, bytes per kh row, per k-step
  uint32_t off_x, off_bar, tmem_cols;
  float* dw;
  long long dw_bstride;
  int cout_real, cin_tot, ci_base, ci_real;
  int unfold_cin;       // > 0: X is W-unfolded (channel c' = kw * unfold_cin + ci), Kw == 1
  int Kww;              // W extent of the weight tensor
};

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tm_du, const __grid_constant__ CUtensorMap tm_x,
             const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* done = bars + 2 * kWgMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x -> (split, class, kd, chunk, mb [, b])
  int r = blockIdx.x;
  const int split = r % p.nsplit; r /= p.nsplit;
  const int cls = r % p.ncls; r /= p.ncls;
  const int kd = (r % p.nkg) * p.kdp; r /= p.nkg;   // first depth tap of this CTA
  const int chunk = r % p.nchunk; r /= p.nchunk;
  const int mb = r % p.nmb; r /= p.nmb;
  const int bfix = r;   // per_batch: sample index
  const int g0 = cls * p.nacc;
  const int na = min(p.nacc, p.ngrp - g0);
  const long long per = (p.items + p.nsplit - 1) / p.nsplit;
  const long long i_lo = (long long)split * per;
  const long long i_hi = min(p.items, i_lo + per);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.nstage; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
    prefetch_tmap(&tm_du);
    prefetch_tmap(&tm_x);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_pp = p.nHt * p.nWt;   // tiles per plane

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (long long it = i_lo; it < i_hi; ++it) {
        int t = (int)(it % tiles_pp);
        long long q = it / tiles_pp;
        const int xd = (int)(q % p.D);     // items walk the X planes; the dY plane of depth tap kd is xd - kd + pad
        const int b = p.per_batch ? bfix : (int)(q / p.D);
        const int d0 = xd - kd + p.pad, d1 = d0 - 1;
        const bool use0 = d0 >= 0 && d0 < p.D;
        const bool use1 = p.kdp == 2 && kd + 1 < p.K && d1 >= 0 && d1 < p.D;
        if (!use0 && !use1) continue;      // only zero padding under these taps
        const int h0 = (t / p.nWt) * 16, w0 = (t % p.nWt) * 8;
        const int s = n % p.nstage;
        mbar_wait(&empty[s], ((n / p.nstage) & 1) ^ 1);
        mbar_expect_tx(&full[s], p.du_box_bytes * (p.kdp == 2 ? 2u : 1u) + p.x_bytes);
        // out-of-range dY planes are zero-filled by TMA (they multiply real X data)
        tma_load_4d(smem + (size_t)s * p.du_bytes, &tm_du, &full[s], w0 * 8, h0, d0,
                    b * p.du_cgtot + p.du_cgoff + mb * 16);
        if (p.kdp == 2)
          tma_load_4d(smem + (size_t)s * p.du_bytes + 16384, &tm_du, &full[s], w0 * 8, h0, kd + 1 < p.K ? d1 : -1,
                      b * p.du_cgtot + p.du_cgoff);
        if (p.stack > 1)
          tma_load_4d(smem + p.off_x + (size_t)s * p.x_bytes, &tm_x, &full[s], (w0 - p.padw) * 8,
                      b * p.x_cgtot + p.x_cgoff + chunk * p.ncg, h0 - p.pad, xd);
        else
          tma_load_4d(smem + p.off_x + (size_t)s * p.x_bytes, &tm_x, &full[s], (w0 - p.padw) * 8, h0 - p.pad, xd,
                      b * p.x_cgtot + p.x_cgoff + chunk * p.ncg);
        ++n;
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // MN-major A and B (bits 15, 16), bf16 x bf16 -> f32, M = 128, N = Nacc
      const uint32_t idesc = umma_idesc_bf16_f32(128, p.Nacc) | (1u << 15) | (1u << 16);
      // A (dY tile [cg][16 h][8 w][8 ch]): LBO = 128 B (next 8 voxels = next h row), SBO = 2048 B (next channel group)
      const uint32_t a_hi = (2048u >> 4) | (1u << 14);
      const uint32_t a_lbo = (128u >> 4) << 16;
      // B (X halo tile): LBO = halo-row pitch (next 8 voxels along K), SBO = channel-group pitch
      const uint32_t b_hi = (p.b_sbo >> 4) | (1u << 14);
      const uint32_t b_lbo = (p.b_lbo >> 4) << 16;
      const uint32_t b_kstep = p.b_kstep >> 4;   // two h rows per k-step of 16 voxels
      int n = 0;
      for (long long it = i_lo; it < i_hi; ++it) {
        const int xd = (int)((it / tiles_pp) % p.D);
        const int d0 = xd - kd + p.pad, d1 = d0 - 1;
        if (!(d0 >= 0 && d0 < p.D) && !(p.kdp == 2 && kd + 1 < p.K && d1 >= 0 && d1 < p.D)) continue;
        const int s = n % p.nstage;
        mbar_wait(&full[s], (n / p.nstage) & 1);
        tc_fence_after();
        const uint32_t a0 = (smem_u32(smem + (size_t)s * p.du_bytes) >> 4) | a_lbo;
        const uint32_t xb = smem_u32(smem + p.off_x + (size_t)s * p.x_bytes);
        for (int a = 0; a < na; ++a) {
          const int g = g0 + a;
          const int kh0 = (g / p.Kw) * p.stack, kw = g % p.Kw;
          const uint32_t b0 = ((xb + (uint32_t)kh0 * p.b_kh + (uint32_t)kw * 16u) >> 4) | b_lbo;
          const uint32_t dcol = tmem_base + (uint32_t)(a * p.Nacc);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16_lohi(dcol, a0 + ks * 16, a_hi, b0 + ks * b_kstep, b_hi, idesc, (n | ks) != 0);
        }
        umma_commit(&empty[s]);
        ++n;
      }
      umma_commit(done);
    }
    __syncwarp();
  } else {
    // ---- TMEM -> fp32 gradient (atomicAdd): thread = accumulator row = output channel
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = p.kdp == 2 ? (row & 63) : mb * 128 + row;
    const int kd_row = p.kdp == 2 ? kd + (row >> 6) : kd;
    // any item with an in-range depth tap?  (same predicate as the producer / issuer loops)
    bool any = false;
    for (long long it = i_lo; it < i_hi && !any; ++it) {
      const int d0 = (int)((it / tiles_pp) % p.D) - kd + p.pad, d1 = d0 - 1;
      any = (d0 >= 0 && d0 < p.D) || (p.kdp == 2 && kd + 1 < p.K && d1 >= 0 && d1 < p.D);
    }
    if (any) {
      mbar_wait(done, 0);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      float* dwb = p.dw + (p.per_batch ? (long long)bfix * p.dw_bstride : 0);
      const int nci = p.ncg * 8;
      for (int a = 0; a < na; ++a) {
        const int g = g0 + a;
        const int kh0 = (g / p.Kw) * p.stack, kw = g % p.Kw;
        for (int c0 = 0; c0 < p.Nacc; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(trow + a * p.Nacc + c0, v);
          tmem_ld_wait();
          if (co < p.cout_real && kd_row < p.K) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = c0 + j;
              const int kh = kh0 + n / nci;
              int ci = chunk * nci + n % nci, kww = kw;
              bool ok;
              if (p.unfold_cin > 0) {
                kww = ci / p.unfold_cin;
                ci = ci % p.unfold_cin;
                ok = kww < p.Kww;
              } else {
                ok = ci < p.ci_real;
              }
              if (ok) {
                const size_t idx = ((((size_t)co * p.cin_tot + p.ci_base + ci) * p.K + kd_row) * p.K + kh) * p.Kww + kww;
                atomicAdd(dwb + idx, __uint_as_float(v[j]));
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled wg_encode() {
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

// dims (W*8, H, D, B*CG), box (64, 16, 1, cg): lands as [cg][16][8][8]
int tmap_du(CUtensorMap* tm, const Act& a, int cg) {
  PFN_encodeTiled enc = wg_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B * a.cg()};
  cuuint64_t gstr[3] = {(cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16, (cuuint64_t)a.W * a.H * a.D * 16};
  cuuint32_t box[4] = {64u, 16u, 1u, (cuuint32_t)cg};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad dY) failed (" + std::to_string((int)r) + ")");
  return 0;
}
// stacked: dims (W*8, B*CG, H, D), box (BW*8, ncg, BH, 1): lands as [BH][ncg][BW][8]
// natural: dims (W*8, H, D, B*CG), box (BW*8, BH, 1, ncg): lands as [ncg][BH][BW][8]
int tmap_x(CUtensorMap* tm, const Act& a, int BW, int BH, int ncg, bool stacked) {
  PFN_encodeTiled enc = wg_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  if (!stacked) {
    cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B * a.cg()};
    cuuint64_t gstr[3] = {(cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16, (cuuint64_t)a.W * a.H * a.D * 16};
    cuuint32_t box[4] = {(cuuint32_t)BW * 8, (cuuint32_t)BH, 1u, (cuuint32_t)ncg};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad X, natural) failed (" + std::to_string((int)r) + ")");
    return 0;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.B * a.cg(), (cuuint64_t)a.H, (cuuint64_t)a.D};
  cuuint64_t gstr[3] = {(cuuint64_t)a.W * a.H * a.D * 16, (cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16};
  cuuint32_t box[4] = {(cuuint32_t)BW * 8, (cuuint32_t)ncg, (cuuint32_t)BH, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad X) failed (" + std::to_string((int)r) + ")");
  return 0;
}

}  // namespace

// dw (fp32, [cout][cin_tot][K][K][Kww], pre-zeroed or holding a running sum) += the gradient
// contribution of source `x` (channel groups [x_cgoff, x_cgoff + x_cg)), which occupies the weight's
// input channels [ci_base, ci_base + ci_real).  per_batch: dw has one slab per sample (dw_bstride).
int conv_wgrad(const Act& x, int x_cgoff, int x_cg, const Act& dy, int dy_cgoff, int cout_real, int ksize,
               int unfold_cin, float* dw, int cin_tot, int ci_base, int ci_real, long long dw_bstride,
               cudaStream_t st) {
  FTB_CHECK(x.B == dy.B && x.D == dy.D && x.H == dy.H && x.W == dy.W, "wgrad: dims");
  FTB_CHECK(ksize == 1 || ksize == 3 || ksize == 5 || ksize == 7, "wgrad: ksize");
  FTB_CHECK(x_cg > 0 && x_cg % 2 == 0, "wgrad: channel groups must be a positive multiple of 2");
  WgradParams p{};
  p.B = x.B; p.D = x.D; p.H = x.H; p.W = x.W;
  p.K = ksize; p.pad = (ksize - 1) / 2;
  p.Kww = ksize;
  p.unfold_cin = unfold_cin;
  p.Kw = unfold_cin > 0 ? 1 : ksize;
  p.padw = (p.Kw - 1) / 2;
  // stack the kh taps along N when a chunk of >= 16 channels keeps K*chunk <= 256 columns
  p.stack = (ksize > 1 && unfold_cin == 0 && getenv("FTB_WGRAD_NOSTACK") == nullptr) ? ksize : 1;
  const int cap = 256 / (8 * p.stack);   // channel groups per chunk
  p.ncg = 0;
  for (int d = x_cg < cap ? x_cg : cap; d >= 2; --d)
    if (x_cg % d == 0 && d % 2 == 0) { p.ncg = d; break; }
  FTB_CHECK(p.ncg > 0, "wgrad: no valid channel chunk");
  p.nchunk = x_cg / p.ncg;
  p.Nacc = p.stack * p.ncg * 8;
  p.ngrp = (p.K / p.stack) * p.Kw;
  p.nacc = 512 / p.Nacc < p.ngrp ? 512 / p.Nacc : p.ngrp;
  p.ncls = cdiv(p.ngrp, p.nacc);
  uint32_t tc = 32;
  while (tc < (uint32_t)(p.nacc * p.Nacc)) tc <<= 1;
  p.tmem_cols = tc;
  const int cout_cg = cdiv(cout_real, 8);
  p.nmb = cdiv(cout_cg, 16);
  const int du_box_cg = cout_cg < 16 ? cout_cg : 16;
  p.nHt = cdiv(x.H, 16); p.nWt = cdiv(x.W, 8);
  p.per_batch = dw_bstride != 0 ? 1 : 0;
  p.items = (long long)p.D * p.nHt * p.nWt * (p.per_batch ? 1 : p.B);
  const int BH = 16 + p.K - 1, BW = 8 + p.Kw - 1;
  const uint32_t P = BW * 16;
  if (p.stack > 1) {   // [BH][ncg][BW][8]
    p.b_sbo = P; p.b_lbo = p.ncg * P; p.b_kh = p.ncg * P; p.b_kstep = 2 * p.ncg * P;
  } else {             // [ncg][BH][BW][8]
    p.b_sbo = BH * P; p.b_lbo = P; p.b_kh = P; p.b_kstep = 2 * P;
  }
  // Cout <= 64: two depth taps per CTA share the 128 accumulator rows (rows 0-63: kd0, rows 64-127: kd0+1)
  p.kdp = (p.K > 1 && p.nmb == 1 && cout_cg <= 8 && getenv("FTB_WGRAD_NOPAIR") == nullptr) ? 2 : 1;
  p.nkg = cdiv(p.K, p.kdp);
  p.du_box_bytes = du_box_cg * 2048;
  p.du_bytes = p.kdp == 2 ? 32768 : du_box_cg * 2048;
  p.x_bytes = (uint32_t)BH * p.ncg * P;
  FTB_CHECK(p.x_bytes % 128 == 0, "wgrad: X box bytes must be a multiple of 128");
  // stages: all dY tiles first, then all X tiles, then 32 KB of readable slack for the M=128 over-read
  int ns = kWgMaxStages;
  const size_t slack = p.kdp == 2 ? 0 : 32768;   // paired taps: the 128 rows are exactly the stage
  auto total = [&](int n) { return (size_t)n * p.du_bytes + (size_t)n * p.x_bytes + slack + 256 + 128; };
  while (ns > 2 && total(ns) > kWgSmemLimit) --ns;
  FTB_CHECK(total(ns) <= kWgSmemLimit, "wgrad: stage does not fit shared memory");
  p.nstage = ns;
  p.off_x = (uint32_t)ns * p.du_bytes;
  p.off_bar = (uint32_t)round_up((int)(p.off_x + ns * p.x_bytes + slack), 16);
  const uint32_t smem_bytes = p.off_bar + 256 + 128;
  p.x_cgtot = x.cg(); p.x_cgoff = x_cgoff;
  p.du_cgtot = dy.cg(); p.du_cgoff = dy_cgoff;
  p.dw = dw; p.dw_bstride = dw_bstride;
  p.cout_real = cout_real; p.cin_tot = cin_tot; p.ci_base = ci_base; p.ci_real = ci_real;
  // split the item range so that the grid is about one wave
  const int fixed = p.ncls * p.nkg * p.nchunk * p.nmb * (p.per_batch ? p.B : 1);
  int nsplit = num_sms() / fixed;   // one wave: a CTA holds its accumulators for its whole life
  const long long max_split = (p.items + 3) / 4;   // at least ~4 tiles per CTA
  if (nsplit > max_split) nsplit = (int)max_split;
  if (nsplit < 1) nsplit = 1;
  p.nsplit = nsplit;

  CUtensorMap tmd, tmx;
  FTB_TRY(tmap_du(&tmd, dy, du_box_cg));
  FTB_TRY(tmap_x(&tmx, x, BW, BH, p.ncg, p.stack > 1));
  static bool attr_set = false;
  if (!attr_set) {
    FTB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWgSmemLimit + 128)));
    attr_set = true;
  }
  if (getenv("FTB_CONV_PLAN"))
    fprintf(stderr, "wgrad plan: K%d cin %d cout %d @%dx%dx%d B%d -> kdp %d stack %d ncg %d N %d nacc %d cls %d chunks %d mb %d split %d stages %d smem %u\n",
            p.K, x_cg * 8, cout_real, p.D, p.H, p.W, p.B, p.kdp, p.stack, p.ncg, p.Nacc, p.nacc, p.ncls, p.nchunk, p.nmb,
            p.nsplit, p.nstage, smem_bytes);
  int prof = -1;
  if (prof_enabled()) {
    const double flops = 2.0 * x.B * (double)x.voxels() * ci_real * cout_real * ksize * ksize * ksize;
    prof = prof_begin(st, flops, (double)x.B * x.voxels() * (ci_real + cout_real) * 2.0, 2);
  }
  wgrad_kernel<<<fixed * nsplit, kWgThreads, smem_bytes, st>>>(tmd, tmx, p);
  prof_end(prof, st);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
