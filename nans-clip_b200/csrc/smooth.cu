// smooth.cu — label-smoothed variant of the contrastive loss (train_lora.py:95-110:
// F.cross_entropy(logits, arange, label_smoothing=eps) in both directions).
//
// With z = s * I T^T, N classes and smoothing eps, per row
//     CE_eps_i = lse_i - (1 - eps) z_ii - (eps / N) sum_j z_ij = CE_i + eps z_ii - (eps / N) s I_i . Tsum
// so the smoothed loss is the plain fused loss plus O(N D) terms:
//     loss_eps = loss + eps s / N * sum_i I_i.T_i  -  eps s / N^2 * (Isum . Tsum)
//     dI_i    += eps s / N * T_i - eps s / N^2 * Tsum          (dT_j likewise with I)
//     dloss/ds+= eps / N * sum_i I_i.T_i - eps / N^2 * (Isum . Tsum)
// Two HBM-bound kernels: the column sums + diagonal sum, and the gradient fix-up.
// Algorithmic bytes: stats 2 * rows * D * 4 read; fix-up rows * D * 4 * (2 read + 2 read-modify-write).
#include "common.cuh"

namespace nans {
namespace {

constexpr int ROWS_PER_BLOCK = 32;

// stats[0..D) += sum_r I[r], stats[D..2D) += sum_r T[r], stats[2D] += sum_r I[r].T[r]
__global__ void __launch_bounds__(256) smooth_stats_kernel(const float* __restrict__ I, const float* __restrict__ T,
                                                           long long ld, long long rows, int D,
                                                           float* __restrict__ stats) {
  const long long r0 = static_cast<long long>(blockIdx.x) * ROWS_PER_BLOCK;
  const long long r1 = min(r0 + ROWS_PER_BLOCK, rows);
  float dot = 0.f;
  for (int c = threadIdx.x * 4; c < D; c += 256 * 4) {
    float4 si = make_float4(0.f, 0.f, 0.f, 0.f), st = si;
    for (long long r = r0; r < r1; ++r) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(I + r * ld + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(T + r * ld + c));
      si.x += a.x; si.y += a.y; si.z += a.z; si.w += a.w;
      st.x += b.x; st.y += b.y; st.z += b.z; st.w += b.w;
      dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    atomicAdd(stats + c, si.x); atomicAdd(stats + c + 1, si.y);
    atomicAdd(stats + c + 2, si.z); atomicAdd(stats + c + 3, si.w);
    atomicAdd(stats + D + c, st.x); atomicAdd(stats + D + c + 1, st.y);
    atomicAdd(stats + D + c + 2, st.z); atomicAdd(stats + D + c + 3, st.w);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(stats + 2 * D, t);
  }
}

// dI[r] += a T[r] - b Tsum,  dT[r] += a I[r] - b Isum,  a = g s coef, b = a / N
__global__ void __launch_bounds__(256) smooth_bwd_kernel(float* __restrict__ dI, float* __restrict__ dT,
                                                         const float* __restrict__ I, const float* __restrict__ T,
                                                         long long ld, long long rows, int D,
                                                         const float* __restrict__ stats,
                                                         const float* __restrict__ s_dev,
                                                         const float* __restrict__ g_dev, float coef, float inv_n) {
  const float a = __ldg(g_dev) * __ldg(s_dev) * coef;
  const float b = a * inv_n;
  const long long nvec = rows * (D / 4);
  for (long long v = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * 256) {
    const long long r = v / (D / 4);
    const int c = static_cast<int>(v - r * (D / 4)) * 4;
    const float4 ti = __ldg(reinterpret_cast<const float4*>(T + r * ld + c));
    const float4 ii = __ldg(reinterpret_cast<const float4*>(I + r * ld + c));
    const float4 ts = __ldg(reinterpret_cast<const float4*>(stats + D + c));
    const float4 is = __ldg(reinterpret_cast<const float4*>(stats + c));
    float4* pi = reinterpret_cast<float4*>(dI + r * D + c);
    float4* pt = reinterpret_cast<float4*>(dT + r * D + c);
    float4 gi = *pi, gt = *pt;
    gi.x += a * ti.x - b * ts.x; gi.y += a * ti.y - b * ts.y; gi.z += a * ti.z - b * ts.z; gi.w += a * ti.w - b * ts.w;
    gt.x += a * ii.x - b * is.x; gt.y += a * ii.y - b * is.y; gt.z += a * ii.z - b * is.z; gt.w += a * ii.w - b * is.w;
    *pi = gi;
    *pt = gt;
  }
}

}  // namespace
}  // namespace nans

using namespace nans;

extern "C" int nans_label_smooth_stats(const float* I, const float* T, int64_t ld, int64_t rows, int64_t D,
                                       float* stats, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(rows >= 0 && D > 0 && D % 4 == 0 && D <= (1 << 20), "label_smooth_stats: D must be a positive multiple of 4");
  NANS_REQUIRE(stats != nullptr, "label_smooth_stats: null output");
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(stats) & 15) == 0, "label_smooth_stats: stats must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NANS_CUDA_OK(cudaMemsetAsync(stats, 0, (2 * static_cast<size_t>(D) + 1) * sizeof(float), st));
  if (rows == 0) return NANS_OK;
  NANS_REQUIRE(I && T, "label_smooth_stats: null pointer");
  NANS_REQUIRE(ld >= D && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(I) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(T) & 15) == 0,
               "label_smooth_stats: rows must be 16-byte aligned");
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, ROWS_PER_BLOCK));
  smooth_stats_kernel<<<grid, 256, 0, st>>>(I, T, ld, rows, static_cast<int>(D), stats);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_label_smooth_bwd(float* dI, float* dT, const float* I_rows, const float* T_rows, int64_t ld,
                                     int64_t rows, int64_t D, const float* stats, const float* s_dev,
                                     const float* grad_out_dev, float coef, float inv_n, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(rows >= 0 && D > 0 && D % 4 == 0, "label_smooth_bwd: D must be a positive multiple of 4");
  if (rows == 0) return NANS_OK;
  NANS_REQUIRE(dI && dT && I_rows && T_rows && stats && s_dev && grad_out_dev, "label_smooth_bwd: null pointer");
  NANS_REQUIRE(ld >= D && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(I_rows) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(T_rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(dI) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dT) & 15) == 0 && (reinterpret_cast<uintptr_t>(stats) & 15) == 0,
               "label_smooth_bwd: pointers must be 16-byte aligned");
  const long long nvec = rows * (D / 4);
  const unsigned grid = static_cast<unsigned>(std::min<long long>(ceil_div(nvec, 256), 148 * 16));
  smooth_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dI, dT, I_rows, T_rows, ld, rows, static_cast<int>(D), stats, s_dev, grad_out_dev, coef, inv_n);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}
