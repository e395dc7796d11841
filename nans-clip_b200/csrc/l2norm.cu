// l2norm.cu — kernel (1): row L2-normalise + 16-bit cast, and its backward.
// Replaces cn_clip/clip/model.py:412-413.  HBM-bound: one warp per row, 16-byte accesses, the row
// lives in registers between the sum-of-squares and the scaled store (single pass over HBM).
// Algorithmic bytes per row: D*(sizeof(in) + 2 [+4 if y32] ) (+4 if inv_norm).
#include "common.cuh"

namespace nans {

template <typename T>
struct Vec16;  // 16-byte vector of T, unpacked to floats
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&f)[4]) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
};
template <>
struct Vec16<__half> {
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float (&f)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __bfloat1622float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
};

template <int N>
__device__ __forceinline__ void store16(void* y16, int y_dtype, int64_t off, const float (&f)[N]) {
  // N consecutive elements starting at element offset `off` (N = 4 or 8)
  uint32_t w[N / 2];
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    if (y_dtype == NANS_BF16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
  }
  uint16_t* base = reinterpret_cast<uint16_t*>(y16) + off;
  if constexpr (N == 4) {
    *reinterpret_cast<uint2*>(base) = make_uint2(w[0], w[1]);
  } else {
    *reinterpret_cast<uint4*>(base) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// NV = 16-byte vectors held per lane (row fits in registers when vecs_per_row <= 32*NV).
template <typename T, int NV>
__global__ void __launch_bounds__(256) l2norm_cast_kernel(const T* __restrict__ x, int64_t rows,
                                                          int D, int64_t ld, void* __restrict__ y16,
                                                          int y_dtype, float* __restrict__ y32,
                                                          float* __restrict__ inv_norm,
                                                          int normalize) {
  constexpr int E = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = D / E;
  const T* xr = x + row * ld;
  float f[NV][E];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      Vec16<T>::load(xr + v * E, f[i]);
#pragma unroll
      for (int e = 0; e < E; ++e) ss = fmaf(f[i][e], f[i][e], ss);
    }
  }
  ss = warp_sum(ss);
  // the reference divides by sqrt(sum of squares); 1/sqrt then multiply differs by <= 1 ulp
  const float inv = normalize ? 1.0f / sqrtf(ss) : 1.0f;
  if (inv_norm != nullptr && lane == 0) inv_norm[row] = 1.0f / sqrtf(ss);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 32 * i;
    if (v < nvec) {
      float g[E];
#pragma unroll
      for (int e = 0; e < E; ++e) g[e] = f[i][e] * inv;
      const int64_t off = row * D + static_cast<int64_t>(v) * E;
      if (y16 != nullptr) store16<E>(y16, y_dtype, off, g);
      if (y32 != nullptr) {
#pragma unroll
        for (int e = 0; e < E; e += 4)
          *reinterpret_cast<float4*>(y32 + off + e) = make_float4(g[e], g[e + 1], g[e + 2], g[e + 3]);
      }
    }
  }
}

// Wide rows: two passes over the row (the second one hits L1/L2).
template <typename T>
__global__ void __launch_bounds__(256) l2norm_cast_wide_kernel(const T* __restrict__ x,
                                                               int64_t rows, int D, int64_t ld,
                                                               void* __restrict__ y16, int y_dtype,
                                                               float* __restrict__ y32,
                                                               float* __restrict__ inv_norm,
                                                               int normalize) {
  constexpr int E = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = D / E;
  const T* xr = x + row * ld;
  float ss = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[E];
    Vec16<T>::load(xr + v * E, f);
#pragma unroll
    for (int e = 0; e < E; ++e) ss = fmaf(f[e], f[e], ss);
  }
  ss = warp_sum(ss);
  const float inv = normalize ? 1.0f / sqrtf(ss) : 1.0f;
  if (inv_norm != nullptr && lane == 0) inv_norm[row] = 1.0f / sqrtf(ss);
  for (int v = lane; v < nvec; v += 32) {
    float f[E];
    Vec16<T>::load(xr + v * E, f);
#pragma unroll
    for (int e = 0; e < E; ++e) f[e] *= inv;
    const int64_t off = row * D + static_cast<int64_t>(v) * E;
    if (y16 != nullptr) store16<E>(y16, y_dtype, off, f);
    if (y32 != nullptr) {
#pragma unroll
      for (int e = 0; e < E; e += 4)
        *reinterpret_cast<float4*>(y32 + off + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
    }
  }
}

// dx = inv * (dy - y <dy, y>), y = x * inv
template <typename T>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const T* __restrict__ x, int64_t ld,
                                                         const float* __restrict__ inv_norm,
                                                         const float* __restrict__ dy,
                                                         int64_t rows, int D,
                                                         float* __restrict__ dx) {
  constexpr int E = Vec16<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = D / E;
  const T* xr = x + row * ld;
  const float* dyr = dy + row * D;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    float f[E];
    Vec16<T>::load(xr + v * E, f);
#pragma unroll
    for (int e = 0; e < E; e += 4) {
      float4 g = __ldg(reinterpret_cast<const float4*>(dyr + v * E + e));
      dot = fmaf(f[e] * inv, g.x, dot);
      dot = fmaf(f[e + 1] * inv, g.y, dot);
      dot = fmaf(f[e + 2] * inv, g.z, dot);
      dot = fmaf(f[e + 3] * inv, g.w, dot);
    }
  }
  dot = warp_sum(dot);
  for (int v = lane; v < nvec; v += 32) {
    float f[E];
    Vec16<T>::load(xr + v * E, f);
#pragma unroll
    for (int e = 0; e < E; e += 4) {
      float4 g = __ldg(reinterpret_cast<const float4*>(dyr + v * E + e));
      float4 o;
      o.x = inv * (g.x - f[e] * inv * dot);
      o.y = inv * (g.y - f[e + 1] * inv * dot);
      o.z = inv * (g.z - f[e + 2] * inv * dot);
      o.w = inv * (g.w - f[e + 3] * inv * dot);
      *reinterpret_cast<float4*>(dx + row * D + v * E + e) = o;
    }
  }
}

template <typename T>
static int launch_l2norm(const void* x, int64_t rows, int64_t D, int64_t ld, void* y16,
                         int y16_dtype, float* y32, float* inv_norm, int normalize,
                         cudaStream_t st) {
  constexpr int E = Vec16<T>::N;
  const int nvec = static_cast<int>(D / E);
  const int warps = 8;
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, warps));
  const T* xp = static_cast<const T*>(x);
#define NANS_L2_LAUNCH(NV)                                                                  \
  l2norm_cast_kernel<T, NV><<<grid, warps * 32, 0, st>>>(xp, rows, (int)D, ld, y16, y16_dtype, \
                                                         y32, inv_norm, normalize)
  if (nvec <= 32) NANS_L2_LAUNCH(1);
  else if (nvec <= 64) NANS_L2_LAUNCH(2);
  else if (nvec <= 128) NANS_L2_LAUNCH(4);
  else if (nvec <= 256) NANS_L2_LAUNCH(8);
  else
    l2norm_cast_wide_kernel<T><<<grid, warps * 32, 0, st>>>(xp, rows, (int)D, ld, y16, y16_dtype,
                                                            y32, inv_norm, normalize);
#undef NANS_L2_LAUNCH
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

}  // namespace nans

using namespace nans;

extern "C" int nans_l2norm_cast(const void* x, int x_dtype, int64_t rows, int64_t D, int64_t ld_x,
                                void* y16, int y16_dtype, float* y32, float* inv_norm,
                                int normalize, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(rows >= 0 && D > 0 && D % 8 == 0, "l2norm_cast: D=%lld must be a positive multiple of 8",
               (long long)D);
  NANS_REQUIRE(rows < (1ll << 34), "l2norm_cast: too many rows");
  if (rows == 0) return NANS_OK;
  NANS_REQUIRE(x != nullptr, "l2norm_cast: x is null");
  NANS_REQUIRE(y16 != nullptr || y32 != nullptr || inv_norm != nullptr, "l2norm_cast: no output");
  NANS_REQUIRE(y16 == nullptr || y16_dtype == NANS_F16 || y16_dtype == NANS_BF16,
               "l2norm_cast: y16_dtype must be NANS_F16 or NANS_BF16");
  NANS_REQUIRE(ld_x >= D, "l2norm_cast: ld_x < D");
  const int esz = x_dtype == NANS_F32 ? 4 : 2;
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ld_x * esz) % 16 == 0,
               "l2norm_cast: x must be 16-byte aligned with a 16-byte row pitch");
  NANS_REQUIRE(y16 == nullptr || (reinterpret_cast<uintptr_t>(y16) & 15) == 0, "l2norm_cast: y16 unaligned");
  NANS_REQUIRE(y32 == nullptr || (reinterpret_cast<uintptr_t>(y32) & 15) == 0, "l2norm_cast: y32 unaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (x_dtype) {
    case NANS_F32: return launch_l2norm<float>(x, rows, D, ld_x, y16, y16_dtype, y32, inv_norm, normalize, st);
    case NANS_F16: return launch_l2norm<__half>(x, rows, D, ld_x, y16, y16_dtype, y32, inv_norm, normalize, st);
    case NANS_BF16: return launch_l2norm<__nv_bfloat16>(x, rows, D, ld_x, y16, y16_dtype, y32, inv_norm, normalize, st);
    default: set_error("l2norm_cast: unknown x_dtype %d", x_dtype); return NANS_ERR_ARG;
  }
}

extern "C" int nans_l2norm_bwd(const void* x, int x_dtype, int64_t ld_x, const float* inv_norm,
                               const float* dy, int64_t rows, int64_t D, float* dx, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(rows >= 0 && D > 0 && D % 8 == 0, "l2norm_bwd: D must be a positive multiple of 8");
  if (rows == 0) return NANS_OK;
  NANS_REQUIRE(x && inv_norm && dy && dx, "l2norm_bwd: null pointer");
  const int esz = x_dtype == NANS_F32 ? 4 : 2;
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (ld_x * esz) % 16 == 0 &&
                   (reinterpret_cast<uintptr_t>(dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0,
               "l2norm_bwd: pointers must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(ceil_div(rows, 8));
  switch (x_dtype) {
    case NANS_F32:
      l2norm_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), ld_x, inv_norm, dy, rows, (int)D, dx);
      break;
    case NANS_F16:
      l2norm_bwd_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(x), ld_x, inv_norm, dy, rows, (int)D, dx);
      break;
    case NANS_BF16:
      l2norm_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ld_x, inv_norm, dy, rows, (int)D, dx);
      break;
    default: set_error("l2norm_bwd: unknown x_dtype %d", x_dtype); return NANS_ERR_ARG;
  }
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}
