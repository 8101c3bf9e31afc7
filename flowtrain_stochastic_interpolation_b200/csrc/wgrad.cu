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
//     Cout) and a slice of the (b, d, tile) items; at the end 4 warps store the rows as one fp32
//     partial per CTA (64 contiguous bytes per thread and tcgen05.ld) and wgrad_reduce_kernel sums
//     the partials of a class into the gradient, one owner thread per element.  (The first version
//     added the rows with atomics: thread = output channel means the 32 lanes of every atomic hit
//     32 different lines 5 KB apart, 8 M single-lane L2 atomics per launch on 62 K addresses, which
//     cost 80-130 us per launch whatever the layer size; kept as the fallback without a scratch buffer.)
//   warp 0: TMA producer | warp 1: MMA issuer | warps 2-5: TMEM -> global
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
  uint32_t b_lbo, b_sbo, b_kh, b_kstep;   // X operand: K-direction / N-direction core-matrix strides, bytes per kh row, per k-step row
  int map4;             // 4-D tensor maps (inner dimension = a whole box row of TW voxels x 8 channels)
  int wait_test;        // the stage wait also tests the next stage's barrier (FTB_NO_WAIT_TEST=1 turns it off)
  int dbg;              // FTB_WGRAD_DBG: block 0 prints issuer / producer cycle counters
  int TH, TW, kpr;      // voxel tile (TH x TW = 128), k-steps per tile row (1 when TW == 8: a k-step is two h rows)
  uint32_t off_x, off_bar, tmem_cols;
  float* dw;
  long long dw_bstride;
  int cout_real, cin_tot, ci_base, ci_real;
  int unfold_cin;       // > 0: X is W-unfolded (channel c' = kw * unfold_cin + ci), Kw == 1
  float* partials;      // [grid][128][pcols] per-CTA accumulators (summed by wgrad_reduce_kernel), or null: atomics
  int pcols;            // nacc * Nacc
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
  // contiguous item range of this CTA; the loops below walk it with (tile, plane, sample) counters, no division
  const int per = (int)((p.items + p.nsplit - 1) / p.nsplit);
  const int i_lo = split * per;
  const int i_hi = min((int)p.items, i_lo + per);

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
      uint32_t s = 0, ph = 0;
      long long pw = 0, pt0 = clock64();
      // items walk the X planes; the dY plane of depth tap kd is xd - kd + pad
      int t = i_lo % tiles_pp, xd = (i_lo / tiles_pp) % p.D, bb = (i_lo / tiles_pp) / p.D;
      for (int it = i_lo; it < i_hi; ++it, ++t) {
        if (t == tiles_pp) { t = 0; if (++xd == p.D) { xd = 0; ++bb; } }
        const int b = p.per_batch ? bfix : bb;
        const int d0 = xd - kd + p.pad, d1 = d0 - 1;
        const bool use0 = d0 >= 0 && d0 < p.D;
        const bool use1 = p.kdp == 2 && kd + 1 < p.K && d1 >= 0 && d1 < p.D;
        if (!use0 && !use1) continue;      // only zero padding under these taps
        const int ht = t / p.nWt;
        const int h0 = ht * p.TH, w0 = (t - ht * p.nWt) * p.TW;
        long long pq = 0;
        if (p.dbg) pq = clock64();
        mbar_wait_backoff(&empty[s], ph ^ 1);
        if (p.dbg) pw += clock64() - pq;
        if (p.dbg == 2) {   // timing experiment: no loads at all (results are garbage), the MMA stream alone
          mbar_arrive(&full[s]);
          ++n;
          if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; }
          continue;
        }
        mbar_expect_tx(&full[s], p.du_box_bytes * (p.kdp == 2 ? 2u : 1u) + p.x_bytes);
        // out-of-range dY planes are zero-filled by TMA (they multiply real X data)
        if (p.map4) {
          tma_load_4d(smem + (size_t)s * p.du_bytes, &tm_du, &full[s], w0 * 8, h0, d0, b * p.du_cgtot + p.du_cgoff + mb * 16);
          if (p.kdp == 2)
            tma_load_4d(smem + (size_t)s * p.du_bytes + 16384, &tm_du, &full[s], w0 * 8, h0, kd + 1 < p.K ? d1 : -1,
                        b * p.du_cgtot + p.du_cgoff);
          if (p.stack > 1)
            tma_load_4d(smem + p.off_x + (size_t)s * p.x_bytes, &tm_x, &full[s], (w0 - p.padw) * 8,
                        b * p.x_cgtot + p.x_cgoff + chunk * p.ncg, h0 - p.pad, xd);
          else
            tma_load_4d(smem + p.off_x + (size_t)s * p.x_bytes, &tm_x, &full[s], (w0 - p.padw) * 8, h0 - p.pad, xd,
                        b * p.x_cgtot + p.x_cgoff + chunk * p.ncg);
        } else {
        tma_load_5d(smem + (size_t)s * p.du_bytes, &tm_du, &full[s], 0, w0, h0, d0,
                      b * p.du_cgtot + p.du_cgoff + mb * 16);
          if (p.kdp == 2)
            tma_load_5d(smem + (size_t)s * p.du_bytes + 16384, &tm_du, &full[s], 0, w0, h0, kd + 1 < p.K ? d1 : -1,
                        b * p.du_cgtot + p.du_cgoff);
          if (p.stack > 1)
            tma_load_5d(smem + p.off_x + (size_t)s * p.x_bytes, &tm_x, &full[s], 0, w0 - p.padw,
                        b * p.x_cgtot + p.x_cgoff + chunk * p.ncg, h0 - p.pad, xd);
          else
            tma_load_5d(smem + p.off_x + (size_t)s * p.x_bytes, &tm_x, &full[s], 0, w0 - p.padw, h0 - p.pad, xd,
                        b * p.x_cgtot + p.x_cgoff + chunk * p.ncg);
        }
        ++n;
        if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; }
      }
      if (p.dbg && blockIdx.x == 0) printf("wgrad dbg producer: total %lld wait_empty %lld stages %d\n", clock64() - pt0, pw, n);
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
      // The issuer is ONE thread and its instruction stream is on the critical path (an M=128 x N=144 MMA takes 72
      // cycles; a dependent integer instruction of a single warp ~6): everything that does not change between stages
      // is computed here, once - per accumulator the B start offset of its (kh group, kw) and its TMEM column - and
      // the stage / phase cursors advance incrementally (no division in the loop).
      constexpr int kUnrollAcc = 4;
      uint32_t boff[kUnrollAcc], dcol[kUnrollAcc];
#pragma unroll
      for (int a = 0; a < kUnrollAcc; ++a) {
        const int g = g0 + (a < na ? a : 0);
        const int kh0 = (g / p.Kw) * p.stack, kw = g % p.Kw;
        boff[a] = ((uint32_t)kh0 * p.b_kh + (uint32_t)kw * 16u) >> 4;
        dcol[a] = tmem_base + (uint32_t)(a * p.Nacc);
      }
      uint32_t kso[8];   // B offset of k-step ks: tile rows of kpr k-steps (TW == 8: one k-step = two h rows)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) kso[ks] = (uint32_t)(ks / p.kpr) * b_kstep + (uint32_t)(ks % p.kpr) * 16u;
      const uint32_t du_enc = p.du_bytes >> 4, x_enc = p.x_bytes >> 4;
      const uint32_t a_base = (smem_u32(smem) >> 4) | a_lbo;
      const uint32_t x_base = (smem_u32(smem + p.off_x) >> 4) | b_lbo;
      int n = 0;
      uint32_t s = 0, ph = 0;
      bool ready = false;   // the stage at the cursor was already seen full
      long long iw = 0, it0 = clock64();
      int t = i_lo % tiles_pp, xd = (i_lo / tiles_pp) % p.D;
      for (int it = i_lo; it < i_hi; ++it, ++t) {
        if (t == tiles_pp) { t = 0; if (++xd == p.D) xd = 0; }
        const int d0 = xd - kd + p.pad, d1 = d0 - 1;
        if (!(d0 >= 0 && d0 < p.D) && !(p.kdp == 2 && kd + 1 < p.K && d1 >= 0 && d1 < p.D)) continue;
        long long iq = 0;
        if (p.dbg) iq = clock64();
        if (!ready) {   // also tests the next stage's barrier: when that is already full its wait is skipped
          uint32_t s1 = s + 1, ph1 = ph;
          if (s1 == (uint32_t)p.nstage) { s1 = 0; ph1 ^= 1; }
          ready = mbar_wait_test_next(&full[s], ph, &full[s1], ph1, p.wait_test != 0);
        } else {
          ready = false;
        }
        if (p.dbg) iw += clock64() - iq;
        tc_fence_after();
        const uint32_t a0 = a_base + s * du_enc;
        const uint32_t xb = x_base + s * x_enc;
        const uint32_t first = n != 0 ? 1u : 0u;
#pragma unroll
        for (int a = 0; a < kUnrollAcc; ++a) {
          if (a < na) {
            const uint32_t b0 = xb + boff[a];
            umma_bf16_lohi(dcol[a], a0, a_hi, b0, b_hi, idesc, first);
#pragma unroll
            for (int ks = 1; ks < 8; ++ks)
              umma_bf16_lohi(dcol[a], a0 + ks * 16, a_hi, b0 + kso[ks], b_hi, idesc, 1u);
          }
        }
        for (int a = kUnrollAcc; a < na; ++a) {
          const int g = g0 + a;
          const int kh0 = (g / p.Kw) * p.stack, kw = g % p.Kw;
          const uint32_t b0 = xb + (((uint32_t)kh0 * p.b_kh + (uint32_t)kw * 16u) >> 4);
          const uint32_t dc = tmem_base + (uint32_t)(a * p.Nacc);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16_lohi(dc, a0 + ks * 16, a_hi, b0 + kso[ks], b_hi, idesc, (n | ks) != 0);
        }
        umma_commit(&empty[s]);
        ++n;
        if (++s == (uint32_t)p.nstage) { s = 0; ph ^= 1; }
      }
      umma_commit(done);
      const long long it1 = clock64();
      mbar_wait(done, 0);
      if (p.dbg && blockIdx.x == 0)
        printf("wgrad dbg issuer: loop %lld wait_full %lld drain %lld stages %d (MMA time %lld)\n", it1 - it0, iw, clock64() - it1, n,
               (long long)n * na * 8 * (p.Nacc / 2 > 32 + p.Nacc / 4 ? p.Nacc / 2 : 32 + p.Nacc / 4));
    }
    __syncwarp();
    tc_fence_before();
    // the epilogue warps sleep on a hardware barrier instead of polling the mbarrier (shared-memory cycles the
    // tensor core needs for its operand fetches)
    asm volatile("bar.sync 1, 160;" ::: "memory");
  } else {
    // ---- TMEM -> fp32 gradient (atomicAdd): thread = accumulator row = output channel
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int co = p.kdp == 2 ? (row & 63) : mb * 128 + row;
    const int kd_row = p.kdp == 2 ? kd + (row >> 6) : kd;
    // any item with an in-range depth tap?  (same predicate as the producer / issuer loops)
    bool any = false;
    for (int it = i_lo; it < i_hi && !any; ++it) {
      const int d0 = (it / tiles_pp) % p.D - kd + p.pad, d1 = d0 - 1;
      any = (d0 >= 0 && d0 < p.D) || (p.kdp == 2 && kd + 1 < p.K && d1 >= 0 && d1 < p.D);
    }
    // all MMAs of this CTA have completed once the issuer warp arrives here
    asm volatile("bar.sync 1, 160;" ::: "memory");
    tc_fence_after();
    if (p.partials) {
      // ---- per-CTA partial: [row][pcols] fp32, plain 16-byte stores (zeros when this CTA had no item)
      float* out = p.partials + ((size_t)blockIdx.x * 128 + row) * p.pcols;
      const bool row_ok = co < p.cout_real && kd_row < p.K;
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int a = 0; a < na; ++a)
        for (int c0 = 0; c0 < p.Nacc; c0 += 16) {
          uint32_t v[16];
          if (any) {
            tmem_ld16(trow + a * p.Nacc + c0, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0u;
          }
          if (row_ok) {
            uint4* o4 = reinterpret_cast<uint4*>(out + a * p.Nacc + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) o4[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
    } else
    if (any) {
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

// dw[idx(class, row, col)] += sum over the nsplit partials of the class; one thread per (class, row, col), consecutive
// threads = consecutive columns (coalesced reads of every partial), 16 independent loads in flight per thread (the
// pass is latency-bound: a 74-deep chain of dependent adds took 13-15 us whatever the grid); the index mapping is
// the epilogue's; each element has exactly one owner, so no atomics.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const WgradParams p, int fixed) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long per_cls = (long long)128 * p.pcols;
  if (e >= (long long)fixed * per_cls) return;
  const int cidx = (int)(e / per_cls);
  const int rem = (int)(e % per_cls);
  const int row = rem / p.pcols, col = rem % p.pcols;
  int r = cidx;
  const int cls = r % p.ncls; r /= p.ncls;
  const int kd = (r % p.nkg) * p.kdp; r /= p.nkg;
  const int chunk = r % p.nchunk; r /= p.nchunk;
  const int mb = r % p.nmb; r /= p.nmb;
  const int bfix = r;
  const int g0 = cls * p.nacc;
  const int na = min(p.nacc, p.ngrp - g0);
  const int a = col / p.Nacc, n = col % p.Nacc;
  if (a >= na) return;
  const int co = p.kdp == 2 ? (row & 63) : mb * 128 + row;
  const int kd_row = p.kdp == 2 ? kd + (row >> 6) : kd;
  if (co >= p.cout_real || kd_row >= p.K) return;
  const int g = g0 + a;
  const int kh0 = (g / p.Kw) * p.stack, kw = g % p.Kw;
  const int nci = p.ncg * 8;
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
  if (!ok) return;
  float* dwb = p.dw + (p.per_batch ? (long long)bfix * p.dw_bstride : 0);
  const size_t idx = ((((size_t)co * p.cin_tot + p.ci_base + ci) * p.K + kd_row) * p.K + kh) * p.Kww + kww;
  const float old = dwb[idx];   // in flight with the partial loads
  const float* src = p.partials + ((size_t)cidx * p.nsplit * 128 + row) * p.pcols + col;
  const size_t sstride = (size_t)128 * p.pcols;
  float acc = 0.f;
  int sp = 0;
  for (; sp + 16 <= p.nsplit; sp += 16) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = __ldcs(src + (size_t)(sp + k) * sstride);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += v[k + 8];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += v[k + 4];
    acc += (v[0] + v[2]) + (v[1] + v[3]);
  }
  if (sp < p.nsplit) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = sp + k < p.nsplit ? __ldcs(src + (size_t)(sp + k) * sstride) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += v[k + 8];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] += v[k + 4];
    acc += (v[0] + v[2]) + (v[1] + v[3]);
  }
  dwb[idx] = old + acc;
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

// dims (8, W, H, D, B*CG), box (8, TW, TH, 1, cg): lands as [cg][TH][TW][8] = [cg][16 blocks of 8 voxels][8]
int tmap_du(CUtensorMap* tm, const Act& a, int cg, int TH, int TW) {
  PFN_encodeTiled enc = wg_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5] = {8, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B * a.cg()};
  cuuint64_t gstr[4] = {16, (cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16, (cuuint64_t)a.W * a.H * a.D * 16};
  cuuint32_t box[5] = {8u, (cuuint32_t)TW, (cuuint32_t)TH, 1u, (cuuint32_t)cg};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad dY) failed (" + std::to_string((int)r) + ")");
  return 0;
}
// stacked: dims (8, W, B*CG, H, D), box (8, BW, ncg, BH, 1): lands as [BH][ncg][BW][8]
// natural: dims (8, W, H, D, B*CG), box (8, BW, BH, 1, ncg): lands as [ncg][BH][BW][8]
int tmap_x(CUtensorMap* tm, const Act& a, int BW, int BH, int ncg, bool stacked) {
  PFN_encodeTiled enc = wg_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (!stacked) {
    cuuint64_t gdim[5] = {8, (cuuint64_t)a.W, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B * a.cg()};
    cuuint64_t gstr[4] = {16, (cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16, (cuuint64_t)a.W * a.H * a.D * 16};
    cuuint32_t box[5] = {8u, (cuuint32_t)BW, (cuuint32_t)BH, 1u, (cuuint32_t)ncg};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad X, natural) failed (" + std::to_string((int)r) + ")");
    return 0;
  }
  cuuint64_t gdim[5] = {8, (cuuint64_t)a.W, (cuuint64_t)a.B * a.cg(), (cuuint64_t)a.H, (cuuint64_t)a.D};
  cuuint64_t gstr[4] = {16, (cuuint64_t)a.W * a.H * a.D * 16, (cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16};
  cuuint32_t box[5] = {8u, (cuuint32_t)BW, (cuuint32_t)ncg, (cuuint32_t)BH, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad X) failed (" + std::to_string((int)r) + ")");
  return 0;
}

// 4-D variants: dims (W*8, H, D, B*CG) resp. (W*8, B*CG, H, D); the innermost box dimension is a whole tile row
int tmap_du4(CUtensorMap* tm, const Act& a, int cg, int TH, int TW) {
  PFN_encodeTiled enc = wg_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B * a.cg()};
  cuuint64_t gstr[3] = {(cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16, (cuuint64_t)a.W * a.H * a.D * 16};
  cuuint32_t box[4] = {(cuuint32_t)TW * 8, (cuuint32_t)TH, 1u, (cuuint32_t)cg};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad dY, 4-D) failed (" + std::to_string((int)r) + ")");
  return 0;
}
int tmap_x4(CUtensorMap* tm, const Act& a, int BW, int BH, int ncg, bool stacked) {
  PFN_encodeTiled enc = wg_encode();
  FTB_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r;
  if (!stacked) {
    cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.H, (cuuint64_t)a.D, (cuuint64_t)a.B * a.cg()};
    cuuint64_t gstr[3] = {(cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16, (cuuint64_t)a.W * a.H * a.D * 16};
    cuuint32_t box[4] = {(cuuint32_t)BW * 8, (cuuint32_t)BH, 1u, (cuuint32_t)ncg};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t gdim[4] = {(cuuint64_t)a.W * 8, (cuuint64_t)a.B * a.cg(), (cuuint64_t)a.H, (cuuint64_t)a.D};
    cuuint64_t gstr[3] = {(cuuint64_t)a.W * a.H * a.D * 16, (cuuint64_t)a.W * 16, (cuuint64_t)a.W * a.H * 16};
    cuuint32_t box[4] = {(cuuint32_t)BW * 8, (cuuint32_t)ncg, (cuuint32_t)BH, 1u};
    r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a.p, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, tma_promo(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  FTB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wgrad X, 4-D) failed (" + std::to_string((int)r) + ")");
  return 0;
}

}  // namespace

size_t conv_wgrad_partial_bytes() { return (size_t)(num_sms() + 12) * 128 * 512 * sizeof(float); }

// dw (fp32, [cout][cin_tot][K][K][Kww], pre-zeroed or holding a running sum) += the gradient
// contribution of source `x` (channel groups [x_cgoff, x_cgoff + x_cg)), which occupies the weight's
// input channels [ci_base, ci_base + ci_real).  per_batch: dw has one slab per sample (dw_bstride).
int conv_wgrad(const Act& x, int x_cgoff, int x_cg, const Act& dy, int dy_cgoff, int cout_real, int ksize,
               int unfold_cin, float* dw, int cin_tot, int ci_base, int ci_real, long long dw_bstride,
               cudaStream_t st, float* partials, size_t partial_bytes) {
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
  // voxel tile TH x TW = 128 voxels: 16 x 8, 8 x 16 or 4 x 32.  The stage ring is bound by TMA delivery, and TMA works
  // in units of the box's INNERMOST dimension: with 5-D maps (innermost = 8 channels = 16 bytes) a stage took 2 630
  // cycles against 1 730 of tensor time; 4-D maps whose innermost dimension is a whole tile row (TW x 8 channels:
  // 256-byte dY rows and 288-byte X rows at TW = 16) brought 48 -> 48 @64^3 from 397 to 315 us, the 7^3 stem from
  // 2.27 to 1.43 ms and the 1^3 48 -> 384 layer from 540 to 281 us (profiles/r02_wgrad_tma_maps.log).  TW = 32 needs
  // the 5-D form (a 34-voxel halo row exceeds the 256-element box limit) and is slower.
  static const int tw_env = getenv("FTB_WGRAD_TW") ? atoi(getenv("FTB_WGRAD_TW")) : 16;
  p.TW = tw_env;
  while (p.TW > 8 && p.TW / 2 >= x.W) p.TW /= 2;
  FTB_CHECK(p.TW == 8 || p.TW == 16 || p.TW == 32, "wgrad: tile width");
  p.TH = 128 / p.TW;
  p.kpr = p.TW == 8 ? 1 : p.TW / 16;
  p.nHt = cdiv(x.H, p.TH); p.nWt = cdiv(x.W, p.TW);
  p.per_batch = dw_bstride != 0 ? 1 : 0;
  p.items = (long long)p.D * p.nHt * p.nWt * (p.per_batch ? 1 : p.B);
  const int BH = p.TH + p.K - 1, BW = p.TW + p.Kw - 1;
  const uint32_t P = BW * 16;
  if (p.stack > 1) {   // [BH][ncg][BW][8]
    p.b_sbo = P; p.b_kh = p.ncg * P;
  } else {             // [ncg][BH][BW][8]
    p.b_sbo = BH * P; p.b_kh = P;
  }
  // a k-step is 16 voxels = two 8-voxel blocks: h-adjacent rows when TW == 8, w-adjacent blocks otherwise
  p.b_lbo = p.TW == 8 ? p.b_kh : 128;
  p.b_kstep = p.TW == 8 ? 2 * p.b_kh : p.b_kh;
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
  static const int dbg_env = getenv("FTB_WGRAD_DBG") ? atoi(getenv("FTB_WGRAD_DBG")) : 0;
  p.dbg = dbg_env;
  static const int wt_env = getenv("FTB_NO_WAIT_TEST") ? 0 : 1;
  p.wait_test = wt_env;

  CUtensorMap tmd, tmx;
  // 4-D maps whenever a box row fits the 256-element limit of a box dimension (TW <= 16 with a W halo)
  static const int map4_env = getenv("FTB_WGRAD_MAP4") ? atoi(getenv("FTB_WGRAD_MAP4")) : 1;
  p.map4 = (map4_env && BW * 8 <= 256 && p.TW * 8 <= 256) ? 1 : 0;
  if (p.map4) {
    FTB_TRY(tmap_du4(&tmd, dy, du_box_cg, p.TH, p.TW));
    FTB_TRY(tmap_x4(&tmx, x, BW, BH, p.ncg, p.stack > 1));
  } else {
    FTB_TRY(tmap_du(&tmd, dy, du_box_cg, p.TH, p.TW));
    FTB_TRY(tmap_x(&tmx, x, BW, BH, p.ncg, p.stack > 1));
  }
  static bool attr_set = false;
  if (!attr_set) {
    FTB_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kWgSmemLimit + 128)));
    attr_set = true;
  }
  if (getenv("FTB_CONV_PLAN"))
    fprintf(stderr, "wgrad plan: K%d cin %d cout %d @%dx%dx%d B%d -> tile %dx%d kdp %d stack %d ncg %d N %d nacc %d cls %d chunks %d mb %d split %d stages %d smem %u\n",
            p.K, x_cg * 8, cout_real, p.D, p.H, p.W, p.B, p.TH, p.TW, p.kdp, p.stack, p.ncg, p.Nacc, p.nacc, p.ncls, p.nchunk, p.nmb,
            p.nsplit, p.nstage, smem_bytes);
  int prof = -1;
  if (prof_enabled()) {
    const double flops = 2.0 * x.B * (double)x.voxels() * ci_real * cout_real * ksize * ksize * ksize;
    prof = prof_begin(st, flops, (double)x.B * x.voxels() * (ci_real + cout_real) * 2.0, 2);
  }
  // per-CTA partials + a reduce pass when the caller's scratch holds them; else the atomic epilogue
  p.pcols = p.nacc * p.Nacc;
  const size_t need = (size_t)fixed * nsplit * 128 * p.pcols * sizeof(float);
  static const bool no_partials = getenv("FTB_WGRAD_ATOMIC") != nullptr;
  p.partials = (partials != nullptr && need <= partial_bytes && !no_partials) ? partials : nullptr;
  wgrad_kernel<<<fixed * nsplit, kWgThreads, smem_bytes, st>>>(tmd, tmx, p);
  FTB_LAUNCH_OK();
  if (p.partials) {
    const long long total = (long long)fixed * 128 * p.pcols;
    wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p, fixed);
    FTB_LAUNCH_OK();
  }
  prof_end(prof, st);
  return 0;
}

}  // namespace ftb
