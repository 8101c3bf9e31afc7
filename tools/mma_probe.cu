// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16, no-swizzle K-major operands) as a
// function of N, accumulator reuse pattern and operand layout pitch.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu -I flowtrain_stochastic_interpolation_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ftb_common.cuh"
using namespace ftb;

struct Args { int N, nacc, reps, a_sbo, a_lbo, nA, a_off; };

__global__ void __launch_bounds__(128, 1) probe(Args a, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_ptr;
  if (warp == 1 && elect_one()) {
    const uint32_t a_hi = ((a.a_sbo >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t a_lo0 = (smem_u32(smem) >> 4) | (((a.a_lbo >> 4) & 0x3FFF) << 16);
    const uint32_t b_hi = (256u >> 4) | (1u << 14);
    const uint32_t b_lo0 = (smem_u32(smem + 128 * 1024) >> 4) | ((128u >> 4) << 16);
    const uint32_t idesc = umma_idesc_bf16_f32(128, a.N);
    long long t0 = clock64();
    const uint32_t dmask = (uint32_t)a.nacc - 1;
    for (int r = 0; r < a.reps; r += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint32_t d = tmem_base + ((uint32_t)u & dmask) * a.N;
        umma_bf16_lohi(d, a_lo0 + (u & 3) * 64 + a.a_off * ((u % 3)), a_hi, b_lo0 + u * 32, b_hi, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d_out; cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int Ns[] = {16, 48, 96, 144, 192, 240, 256};
  printf("N  nacc  a_sbo a_lbo  cyc/mma  (math N/2)\n");
  for (int layout = 1; layout < 4; ++layout)
    for (int nacc : {1})
      for (int N : Ns) {
        if (nacc * N > 512) continue;
        // layout 0: dense core matrices (SBO 128... a GEMM-like tile: LBO = 128 rows*16 = 2048, SBO = 128)
        // layout 1: conv plane layout: SBO = row pitch 160 (10 voxels), LBO = cg pitch 2880 (18 rows)
        Args a{N, nacc, 4000, layout ? 160 : 128, layout ? 2880 : 2048, 4, layout == 2 ? 1 : (layout == 3 ? 11 : 0)};
        probe<<<148, 128, 210 * 1024>>>(a, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long cyc; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
        printf("%3d  %d  %4d %4d off%d  %7.1f   (%d)\n", N, nacc, a.a_sbo, a.a_lbo, a.a_off, (double)cyc / a.reps, N / 2);
      }
  return 0;
}
