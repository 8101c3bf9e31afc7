// Time path: learned-Fourier embedding + time MLP + all per-block FiLM MLPs (fp32 throughout).
// Reference: unet_attn_3d.py:203-208 (y = sqrt2*cos(t*f + phi)), :551-556 (Linear, GELU(erf),
// Linear), :255-257/:269-271 (SiLU -> Linear(time_dim -> 2*Cout) per ResnetBlock).
// |t*f| reaches ~3e3 rad (bandwidth 1000), so the product and the phase add are kept as two
// separately rounded fp32 ops (no FMA) and cosf() is the full-range-reduction version.
#include "ops.h"

namespace ftb {

namespace {

__global__ void __launch_bounds__(1024)
time_embed_kernel(TimeMlpParams p, const float* __restrict__ t, float* __restrict__ temb,
                  float* __restrict__ temb_silu, float* __restrict__ save) {
  extern __shared__ float sm[];
  float* y = sm;                 // [time_res]
  float* h1 = sm + p.time_res;   // [time_dim]
  const int b = blockIdx.x;
  const float tv = __ldg(t + b);
  for (int i = threadIdx.x; i < p.time_res; i += blockDim.x) {
    const float arg = __fadd_rn(__fmul_rn(tv, __ldg(p.freqs + i)), __ldg(p.phases + i));
    y[i] = __fmul_rn(cosf(arg), 1.41421356237309504880f);
    if (save) save[(size_t)b * (p.time_res + 2 * p.time_dim) + i] = y[i];   // training: y | pre-GELU | GELU
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int r = warp; r < p.time_dim; r += nw) {
    const float* wr = p.w1 + (size_t)r * p.time_res;
    float s = 0.f;
    for (int i = lane; i < p.time_res; i += 32) s += __ldg(wr + i) * y[i];
    s = warp_sum(s);
    if (lane == 0) {
      const float v = s + __ldg(p.b1 + r);
      h1[r] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));  // exact GELU
      if (save) {
        save[(size_t)b * (p.time_res + 2 * p.time_dim) + p.time_res + r] = v;
        save[(size_t)b * (p.time_res + 2 * p.time_dim) + p.time_res + p.time_dim + r] = h1[r];
      }
    }
  }
  __syncthreads();
  for (int r = warp; r < p.time_dim; r += nw) {
    const float* wr = p.w2 + (size_t)r * p.time_dim;
    float s = 0.f;
    for (int i = lane; i < p.time_dim; i += 32) s += __ldg(wr + i) * h1[i];
    s = warp_sum(s);
    if (lane == 0) {
      const float v = s + __ldg(p.b2 + r);
      temb[(size_t)b * p.time_dim + r] = v;
      temb_silu[(size_t)b * p.time_dim + r] = v / (1.f + expf(-v));
    }
  }
}

// one warp per output row (over all FiLM MLPs), looping over the batch
__global__ void __launch_bounds__(256)
film_kernel(FilmTable ft, const float* __restrict__ x, int B, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= ft.total_rows) return;
  int lo = 0, hi = ft.nblk;  // row_off[lo] <= row < row_off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(ft.row_off + mid) <= row) lo = mid; else hi = mid;
  }
  const int r = row - __ldg(ft.row_off + lo);
  const float* wr = ft.w[lo] + (size_t)r * ft.time_dim;
  const float bias = __ldg(ft.b[lo] + r);
  const int half = (__ldg(ft.row_off + lo + 1) - __ldg(ft.row_off + lo)) >> 1;
  const float* gsp = ft.gs ? ft.gs[lo] : nullptr;
  const bool is_scale = r < half;
  const float gain = (is_scale && gsp) ? __ldg(gsp + r) : 1.f;
  for (int b = 0; b < B; ++b) {
    const float* xb = x + (size_t)b * ft.time_dim;
    float s = 0.f;
    for (int i = lane; i < ft.time_dim; i += 32) s += __ldg(wr + i) * __ldg(xb + i);
    s = warp_sum(s);
    if (lane == 0) out[(size_t)b * ft.total_rows + row] = is_scale ? (s + bias + 1.f) * gain : s + bias;
  }
}

}  // namespace

int time_embed(const TimeMlpParams& p, const float* t, int B, float* temb, float* temb_silu,
               cudaStream_t st, float* save) {
  const size_t smem = (size_t)(p.time_res + p.time_dim) * sizeof(float);
  FTB_CHECK(smem <= 48 * 1024, "time_embed: time_resolution + time_dim too large for shared memory");
  time_embed_kernel<<<B, 1024, smem, st>>>(p, t, temb, temb_silu, save);   // 32 warps: the row loops are latency-bound
  FTB_LAUNCH_OK();
  return 0;
}

int film_mlps(const FilmTable& ft, const float* temb_silu, int B, float* out, cudaStream_t st) {
  const int warps_per_block = 8;
  film_kernel<<<cdiv(ft.total_rows, warps_per_block), 256, 0, st>>>(ft, temb_silu, B, out);
  FTB_LAUNCH_OK();
  return 0;
}

}  // namespace ftb
