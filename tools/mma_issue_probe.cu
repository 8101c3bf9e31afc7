// Microbenchmark: can ONE thread keep the tensor pipe fed with the conv kernel's MMA stream (NZ = 4, N = 48, 3^3:
// per tap the window planes issue N = 48, 96, 144, 144, 96, 48 at three k-steps: 1 032 cycles of tensor time) when
// the whole warp computes the descriptors on the uniform datapath (values from kernel parameters and loop counters,
// only the MMA predicated on the elected lane) instead of one diverged lane using vector registers + R2UR?
// Prints cycles per tap.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_issue_probe tools/mma_issue_probe.cu -I flowtrain_stochastic_interpolation_b200/csrc
#include <cstdio>
#include "ftb_common.cuh"
using namespace ftb;

struct P {
  uint32_t slot_enc, nslot, kinc, kstep, a_hi, b_hi, wchunk_enc, rowp_enc, nb_enc, idesc0, idesc_n;
  int nze, N, ngroups, mode;
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ P p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1) {
    const uint32_t planes_enc = (smem_u32(smem) >> 4) | (((2880u >> 4) & 0x3FFFu) << 16);
    const uint32_t w_enc = (smem_u32(smem + 176 * 1024) >> 4) | ((128u >> 4) << 16);
    uint32_t slot_w0 = 0;
    const long long t0 = clock64();
    if (p.mode == 0) {
      // uniform: all 32 lanes run the loop, the MMA alone is predicated
      for (int g = 0; g < p.ngroups; ++g) {
        const int win = p.nze + 2;
        const uint32_t acc0 = tmem_base + (uint32_t)((g & 1) * p.nze * p.N);
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t aoff = kh * p.rowp_enc + kw;
            const uint32_t wb = w_enc + ((kh * 3 + kw) % 3) * p.wchunk_enc;
            for (int q = 0; q < win; ++q) {
              const int kd_lo = q - (p.nze - 1) > 0 ? q - (p.nze - 1) : 0, kd_hi = q < 2 ? q : 2;
              const int cnt = kd_hi - kd_lo + 1;
              uint32_t slot = slot_w0 + q;
              if (slot >= p.nslot) slot -= p.nslot;
              const uint32_t a = planes_enc + slot * p.slot_enc + aoff;
              const uint32_t b = (uint32_t)(2 - kd_hi) * p.nb_enc + wb;
              const uint32_t d = acc0 + (uint32_t)((q - kd_hi) * p.N);
              const uint32_t idesc = p.idesc0 + (uint32_t)cnt * p.idesc_n;
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) umma_bf16_lohi(d, a + ks * p.kinc, p.a_hi, b + ks * p.kstep, p.b_hi, idesc, 1u);
              }
            }
          }
        slot_w0 += p.nze;
        if (slot_w0 >= p.nslot) slot_w0 -= p.nslot;
      }
    } else if (elect_one()) {
      // one diverged lane, same arithmetic (vector registers + R2UR)
      for (int g = 0; g < p.ngroups; ++g) {
        const int win = p.nze + 2;
        const uint32_t acc0 = tmem_base + (uint32_t)((g & 1) * p.nze * p.N);
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw) {
            const uint32_t aoff = kh * p.rowp_enc + kw;
            const uint32_t wb = w_enc + ((kh * 3 + kw) % 3) * p.wchunk_enc;
            for (int q = 0; q < win; ++q) {
              const int kd_lo = q - (p.nze - 1) > 0 ? q - (p.nze - 1) : 0, kd_hi = q < 2 ? q : 2;
              const int cnt = kd_hi - kd_lo + 1;
              uint32_t slot = slot_w0 + q;
              if (slot >= p.nslot) slot -= p.nslot;
              const uint32_t a = planes_enc + slot * p.slot_enc + aoff;
              const uint32_t b = (uint32_t)(2 - kd_hi) * p.nb_enc + wb;
              const uint32_t d = acc0 + (uint32_t)((q - kd_hi) * p.N);
              const uint32_t idesc = p.idesc0 + (uint32_t)cnt * p.idesc_n;
#pragma unroll
              for (int ks = 0; ks < 3; ++ks) umma_bf16_lohi(d, a + ks * p.kinc, p.a_hi, b + ks * p.kstep, p.b_hi, idesc, 1u);
            }
          }
        slot_w0 += p.nze;
        if (slot_w0 >= p.nslot) slot_w0 -= p.nslot;
      }
    }
    if (elect_one()) {
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      if (blockIdx.x == 0) out[0] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    P p{};
    p.slot_enc = 17280 >> 4; p.nslot = 10; p.kinc = (2 * 2880) >> 4; p.kstep = (3 * 48 * 32) >> 4;
    p.a_hi = ((160u >> 4) & 0x3FFF) | (1u << 14); p.b_hi = (256u >> 4) | (1u << 14);
    p.wchunk_enc = 13824 >> 4; p.rowp_enc = 160 >> 4; p.nb_enc = 48 * 2;
    p.idesc0 = umma_idesc_bf16_f32(128, 0); p.idesc_n = (uint32_t)(48 >> 3) << 17;
    p.nze = 4; p.N = 48; p.ngroups = 200; p.mode = mode;
    probe<<<148, 128, 226 * 1024>>>(p, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long cyc; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("mode %d (%s): %.1f cycles per tap (tensor time 1032), %.1f per MMA\n", mode, mode ? "one diverged lane" : "uniform warp",
           (double)cyc / (p.ngroups * 9.0), (double)cyc / (p.ngroups * 9.0 * 18.0));
  }
  return 0;
}
