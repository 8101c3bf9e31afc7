// Shared device/host helpers for the flowtrain-b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>

namespace ftb {

// ------------------------------------------------------------------------------------
// error plumbing: every C-ABI entry returns int (0 ok, <0 error); message is thread-local
// ------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(const char* file, int line, const std::string& msg);

#define FTB_FAIL(msg) return ::ftb::fail(__FILE__, __LINE__, (msg))
#define FTB_CHECK(cond, msg)                                   \
  do {                                                         \
    if (!(cond)) return ::ftb::fail(__FILE__, __LINE__, (msg)); \
  } while (0)
#define FTB_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::ftb::fail(__FILE__, __LINE__, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
// after a kernel launch: surface launch errors and count the launch (bench.py "gpu_launches")
#define FTB_LAUNCH_OK()                \
  do {                                 \
    FTB_CUDA(cudaGetLastError());      \
    ::ftb::count_launch(1);            \
  } while (0)
#define FTB_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != 0) return _r;    \
  } while (0)

// ------------------------------------------------------------------------------------
// Activation layout in HBM ("blocked8"): [B][CG][D][H][W][8] bf16, CG = C/8 channel groups.
// A voxel's 8-channel group is one 16-byte unit; W is the fastest spatial dim.  This is the
// no-swizzle K-major UMMA core-matrix layout (8 rows x 16 B) laid out along W, so a TMA box
// lands in shared memory already in operand form and a conv tap is a pure address shift.
// ------------------------------------------------------------------------------------
struct Act {
  __nv_bfloat16* p = nullptr;
  int B = 0, C = 0, D = 0, H = 0, W = 0;  // C is padded to a multiple of 16
  int cg() const { return C / 8; }
  size_t voxels() const { return (size_t)D * H * W; }
  size_t elems() const { return (size_t)B * C * voxels(); }
  size_t bytes() const { return elems() * 2; }
};

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline size_t round_up_sz(size_t x, size_t m) { return (x + m - 1) / m * m; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

int num_sms();
// L2 promotion of every activation tensor map (FTB_TMA_PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B; default 128 B)
static inline CUtensorMapL2promotion tma_promo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FTB_TMA_PROMO");
    v = e ? atoi(e) : 2;
  }
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}
void count_launch(int n);
// optional per-launch CUDA-event timing of the conv kernel (bench.py roofline leg)
bool prof_enabled();
int prof_begin(cudaStream_t st, double flops, double bytes, int kind);
void prof_end(int idx, cudaStream_t st);

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must never hang the GPU box (a hang is a strike); after
// ~2 s of waiting the kernel traps, which surfaces as a CUDA error on the host.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 20000000u) {
      printf("ftb: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// Wait for up to four barriers at once: the four try_wait round trips overlap instead of adding up (a satisfied wait
// still costs its shared-memory latency on the issuing thread's critical path).  Unused entries repeat the last one.
__device__ __forceinline__ bool mbar_try_wait4(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1, uint32_t a2, uint32_t p2,
                                               uint32_t a3, uint32_t p3) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred q0, q1, q2, q3;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q3, [%7], %8;\n\t"
      "and.pred q0, q0, q1;\n\t"
      "and.pred q2, q2, q3;\n\t"
      "and.pred q0, q0, q2;\n\t"
      "selp.u32 %0, 1, 0, q0;\n\t}"
      : "=r"(ok)
      : "r"(a0), "r"(p0), "r"(a1), "r"(p1), "r"(a2), "r"(p2), "r"(a3), "r"(p3)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait4(uint32_t a0, uint32_t p0, uint32_t a1, uint32_t p1, uint32_t a2, uint32_t p2,
                                           uint32_t a3, uint32_t p3) {
  uint32_t spins = 0;
  while (!mbar_try_wait4(a0, p0, a1, p1, a2, p2, a3, p3)) {
    if (++spins > 20000000u) {
      printf("ftb: mbarrier (x4) timeout block %d thread %d bar %u\n", blockIdx.x, threadIdx.x, a0);
      __trap();
    }
  }
}

// Six barriers per round trip (the group start of the conv issuer: accumulator buffer + new planes).
__device__ __forceinline__ void mbar_wait6(const uint32_t (&a)[6], const uint32_t (&p)[6]) {
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred q0, q1, q2, q3, q4, q5;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q3, [%7], %8;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q4, [%9], %10;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q5, [%11], %12;\n\t"
        "and.pred q0, q0, q1;\n\t"
        "and.pred q2, q2, q3;\n\t"
        "and.pred q4, q4, q5;\n\t"
        "and.pred q0, q0, q2;\n\t"
        "and.pred q0, q0, q4;\n\t"
        "selp.u32 %0, 1, 0, q0;\n\t}"
        : "=r"(ok)
        : "r"(a[0]), "r"(p[0]), "r"(a[1]), "r"(p[1]), "r"(a[2]), "r"(p[2]), "r"(a[3]), "r"(p[3]), "r"(a[4]), "r"(p[4]),
          "r"(a[5]), "r"(p[5])
        : "memory");
    if (ok) return;
    if (++spins > 20000000u) {
      printf("ftb: mbarrier (x6) timeout block %d thread %d bar %u\n", blockIdx.x, threadIdx.x, a[0]);
      __trap();
    }
  }
}

// Wait for `bar` and, in the same shared-memory round trip, TEST (non-blocking) whether the barrier the caller will
// need next has completed too: when it has, the caller skips that wait altogether.  Returns the test's result.
__device__ __forceinline__ bool mbar_wait_test_next(uint64_t* bar, uint32_t parity, uint64_t* next, uint32_t next_parity,
                                                    bool test = true) {
  if (!test) { mbar_wait(bar, parity); return false; }
  uint32_t spins = 0;
  while (true) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred q0, q1;\n\t.reg .u32 r0, r1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
        "selp.u32 r0, 1, 0, q0;\n\t"
        "selp.u32 r1, 2, 0, q1;\n\t"
        "or.b32 %0, r0, r1;\n\t}"
        : "=r"(r)
        : "r"(smem_u32(bar)), "r"(parity), "r"(smem_u32(next)), "r"(next_parity)
        : "memory");
    if (r & 1u) return (r & 2u) != 0;
    if (++spins > 20000000u) {
      printf("ftb: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// The same wait for warps that are not on the critical path (producers waiting for a free slot, epilogue warps
// waiting for an accumulator): sleep between polls, so the polling does not take shared-memory cycles from the
// tensor core's operand fetches.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > 20000000u) {
      printf("ftb: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
             threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// TMA: 4-D tiled tensor load, global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, 16-B aligned both sides)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, one CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// one lane of a converged warp (the same lane every call)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// same MMA with the 64-bit descriptors given as (lo, hi) register pairs: the issue loop only
// adds to the 14-bit address field in `lo`
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi,
                                               uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      :
      : "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// UMMA shared-memory matrix descriptor, K-major, no swizzle ("interleave"):
// element (row r, k) lives at  start + (r/8)*SBO + (k/8)*LBO + (r%8)*16 + (k%8)*2  bytes.
__device__ __forceinline__ uint64_t umma_desc_kmajor_noswz(uint32_t saddr, uint32_t lbo_bytes,
                                                           uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // layout_type (bits 61..63) = 0: SWIZZLE_NONE
}
// instruction descriptor: bf16 A/B (K-major both), fp32 accumulate, M=128, N
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_f32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= 1u << 7;                      // a_format = BF16
  d |= 1u << 10;                     // b_format = BF16
  d |= (uint32_t)(N >> 3) << 17;     // n_dim
  d |= (uint32_t)(M >> 4) << 24;     // m_dim
  return d;
}

// Dropout keep-mask, counter-based so the backward regenerates it instead of storing it: one 64-bit hash per
// (voxel, 8-channel group) gives 8 bits per element; element j is dropped when its byte < t = round(256 p), and
// kept values are scaled by 256/(256 - t) (the exact inverse keep probability of this mask).  `group` is the index
// of the 16-byte unit inside the blocked tensor: (b*CG + cg)*voxels + v.
struct DropMask {
  unsigned long long bits;
  unsigned t;
  float scale;
  __device__ __forceinline__ float operator()(int j) const {
    return ((unsigned)(bits >> (8 * j)) & 255u) < t ? 0.f : scale;
  }
};
__device__ __forceinline__ unsigned mix32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ DropMask drop_mask8(float drop_p, unsigned long long key, size_t group) {
  DropMask m;
  m.t = (unsigned)(drop_p * 256.f + 0.5f);
  m.scale = 256.f / (256.f - (float)m.t);
  const unsigned g = (unsigned)group ^ ((unsigned)(group >> 32) * 0x85ebca6bu);
  const unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
  const unsigned h0 = mix32(g ^ k0), h1 = mix32((g + 0x9E3779B9u) ^ k1 ^ h0);
  m.bits = ((unsigned long long)h1 << 32) | h0;
  return m;
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16x2(f[0], f[1]);
  q.y = pack_bf16x2(f[2], f[3]);
  q.z = pack_bf16x2(f[4], f[5]);
  q.w = pack_bf16x2(f[6], f[7]);
  return q;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__

}  // namespace ftb
