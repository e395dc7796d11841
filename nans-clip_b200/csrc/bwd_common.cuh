// bwd_common.cuh — constants, launch parameters and the softmax-gradient inner loop shared by the
// backward kernels (included by strip_bwd.cu only; everything lives in that TU's anonymous namespace).
#pragma once
constexpr int BM = 128;
constexpr int KT = 128;
constexpr int BK = 64;
constexpr int SLICE = 256;
constexpr int A_CHUNK = BM * BK * 2;        // 16 KB
constexpr int B_CHUNK = KT * BK * 2;        // 16 KB
constexpr int PAIR = 2;                     // chunks per ring stage
constexpr int MAX_NR = 6;
constexpr int SM_WARPS = 8;  // softmax-gradient warps: 4 lane groups x 2 column halves
constexpr int NUM_THREADS = 128 + SM_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int TMEM_S = 256;  // two 128-column S buffers; G aliases the first 64 columns of each
constexpr int BAR_BYTES = 1536;  // mbarriers + TMEM pointer (256 B) + two 128-float column-factor buffers
constexpr int CF_OFF = 256;
constexpr size_t SMEM_CAP = 227 * 1024;
constexpr float kFactorRange = 100.0f;  // max spread of base-2 lse for the one-exp formulation
constexpr float kGShiftLog2 = 12.0f;  // G is carried as fp16 scaled by 2^12 (|G| <= 2 -> 8192)

struct BwdParams {
  int row_begin, row_end;  // rows of the local block that receive gradient
  int ncols, D, kchunks;
  int nrb, npass, nsplit, ntiles;
  int nr;
  uint32_t idesc1_fmt;  // operand format bits (0 = f16, 1 = bf16)
  uint32_t g_fmt;       // format G is written in (0 = f16, 1 = bf16)
  int label_shift;      // label column of local row r = r + label_shift
  const float* s_dev;
  const float* grad_out_dev;
  float coef_host;  // grad_mult / (2 N) / 2^12
  const float* lse_row[2];  // per strip: base-2 lse of the local rows (indexed by local row)
  const float* lse_col[2];  // per strip: base-2 lse of all columns
  float* out[2];            // per strip: fp32 [row_end-row_begin, D]
  int accumulate;           // 1: red.add into out (column splits), 0: plain stores
  const int* lse_minmax;    // [2] order-preserving int encodings of min / max of all lse values
  int debug;                // bring-up experiments (NANS_BWD_DEBUG): 1 = no exp in the softmax warps
  int n_full, ns_tail;      // narrow pairs: units [0, n_full) sweep all columns, the rest are split ns_tail ways
  long long total_tiles;    // persistent kernel: 2 * nrb * ntiles tile steps shared out over npairs CTA pairs
  int npairs;
  int npp_units, npp_t1;    // persistent kernel, helper mode (npp_t1 > 0): pairs [0, npp_units) sweep tiles [0, npp_t1) of
                            // their own unit, the remaining pairs share the tiles [npp_t1, ntiles) of all units
  // exchange mode (nans_clip_loss_bwd_xchg): the column operands are the double-buffered gathered tensors
  // [2 slots * ncols, D]; column tile coordinates are shifted by (*col_step & 1) * ncols rows.  Narrow kernels only.
  const uint32_t* col_step;
};
__device__ __forceinline__ int col_slot_offset(const BwdParams& p) {
  return p.col_step != nullptr ? static_cast<int>(*p.col_step & 1u) * p.ncols : 0;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}


// G for 32 columns of one row: g = 2^12 (p_row + p_col) (- 2^13 at the label), packed to 16 bit.
// Everything that is uniform over the tile is a template parameter: with run-time branches inside
// the element loop the compiler predicated both exp formulations, both conversions and the label
// test into every element (~13 issue slots per element instead of ~5).
template <bool FACTORED, bool BF16, bool LABEL>
__device__ __forceinline__ void softmax_grad32(const uint32_t (&r)[32], const float* __restrict__ cf,
                                               float c, float lr2, float a_i, int label_rel,
                                               uint32_t* __restrict__ g16) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 f = *reinterpret_cast<const float4*>(cf + 4 * q);  // smem broadcast
    const float cfv[4] = {f.x, f.y, f.z, f.w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float cosv = __uint_as_float(r[4 * q + e]);
      const float e1 = fast_exp2(fmaf(cosv, c, -lr2));
      if (FACTORED) g[e] = e1 * fmaf(a_i, cfv[e], 1.0f);
      else g[e] = e1 + fast_exp2(fmaf(cosv, c, -cfv[e]));
      if (LABEL) g[e] = (4 * q + e == label_rel) ? g[e] - 8192.0f : g[e];  // 2 * 2^12
    }
#pragma unroll
    for (int e = 0; e < 4; e += 2) {
      if (BF16) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(g[e], g[e + 1]);
        g16[2 * q + (e >> 1)] = *reinterpret_cast<const uint32_t*>(&hh);
      } else {
        const __half2 hh = __floats2half2_rn(g[e], g[e + 1]);
        g16[2 * q + (e >> 1)] = *reinterpret_cast<const uint32_t*>(&hh);
      }
    }
  }
}

template <bool FACTORED, bool BF16>
__device__ __forceinline__ void softmax_grad32_dispatch(bool has_label, const uint32_t (&r)[32],
                                                        const float* __restrict__ cf, float c, float lr2,
                                                        float a_i, int label_rel, uint32_t* __restrict__ g16) {
  if (has_label) softmax_grad32<FACTORED, BF16, true>(r, cf, c, lr2, a_i, label_rel, g16);
  else softmax_grad32<FACTORED, BF16, false>(r, cf, c, lr2, a_i, label_rel, g16);
}
