// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16) with MN-major no-swizzle operands at the strides
// wgrad.cu uses (A = dY tile [cg][16][8][8], B = X halo tile, kh-stacked or natural), against K-major operands.
// mode 0: K-major A and B (reference); mode 1: MN-major A only; mode 2: MN-major B only; mode 3: both MN-major.
// A second group times the k-loop shape of wgrad (8 k-steps x nacc accumulators per stage).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe_mn tools/mma_probe_mn.cu -I flowtrain_stochastic_interpolation_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ftb_common.cuh"
using namespace ftb;

struct Args { int N, M, nacc, reps, mode, a_sbo, a_lbo, b_sbo, b_lbo, a_kstep, b_kstep, b_acc_off; };

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
    const uint32_t b_hi = ((a.b_sbo >> 4) & 0x3FFF) | (1u << 14);
    const uint32_t b_lo0 = (smem_u32(smem + 96 * 1024) >> 4) | (((a.b_lbo >> 4) & 0x3FFF) << 16);
    uint32_t idesc = umma_idesc_bf16_f32(a.M, a.N);
    if (a.mode & 1) idesc |= 1u << 15;
    if (a.mode & 2) idesc |= 1u << 16;
    long long t0 = clock64();
    for (int r = 0; r < a.reps; ++r) {
      for (int acc = 0; acc < a.nacc; ++acc) {
        const uint32_t d = tmem_base + (uint32_t)acc * a.N;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          umma_bf16_lohi(d, a_lo0 + ks * (a.a_kstep >> 4), a_hi, b_lo0 + acc * (a.b_acc_off >> 4) + ks * (a.b_kstep >> 4), b_hi,
                         idesc, 1u);
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
  auto run = [&](const char* what, Args a) {
    probe<<<148, 128, 210 * 1024>>>(a, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s (%s)\n", cudaGetErrorString(e), what); exit(1); }
    long long cyc; cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-46s M%3d N%3d nacc%d mode%d: %7.1f cyc/mma (math %d, fetch %d)\n", what, a.M, a.N, a.nacc, a.mode,
           (double)cyc / ((double)a.reps * a.nacc * 8), a.N / 2, 32 + a.N / 4);
  };
  const int reps = 1000;
  for (int N : {48, 96, 144, 192, 256}) {
    const int nacc = 512 / N < 3 ? 512 / N : 3;
    // K-major reference: A rows 16 B apart in 128-B core matrices: SBO 128 (next 8 rows), LBO 2048 (next 8 k), k-step 4096
    run("K-major dense", Args{N, 128, nacc, reps, 0, 256, 128, 256, 128, 4096, 8192, 16});
    // wgrad 3^3 stacked: A [cg][16h][8w][8]: LBO 128, SBO 2048, kstep 256; B [18][6][10][8]: SBO 160, LBO 960, kstep 1920, kw shift 16
    run("MN A only (wgrad dY tile)", Args{N, 128, nacc, reps, 1, 2048, 128, 128, 4096, 256, 8192, 16});
    run("MN B only (wgrad X stacked, 160/960)", Args{N, 128, nacc, reps, 2, 128, 2048, 160, 960, 4096, 1920, 16});
    run("MN both (wgrad stacked)", Args{N, 128, nacc, reps, 3, 2048, 128, 160, 960, 256, 1920, 16});
    run("MN both, B dense (SBO 128, LBO 4096)", Args{N, 128, nacc, reps, 3, 2048, 128, 128, 4096, 256, 8192, 0});
    run("MN both, B pitch 144 (SBO 144, LBO 864)", Args{N, 128, nacc, reps, 3, 2048, 128, 144, 864, 256, 1728, 16});
    run("MN both, B natural (SBO 2880, LBO 160)", Args{N, 128, nacc, reps, 3, 2048, 128, 2880, 160, 256, 320, 16});
    run("MN both, A dense (SBO 128, LBO 2048)", Args{N, 128, nacc, reps, 3, 128, 2048, 160, 960, 4096, 1920, 16});
    run("MN both M=64", Args{N, 64, nacc, reps, 3, 2048, 128, 160, 960, 256, 1920, 16});
  }
  return 0;
}
