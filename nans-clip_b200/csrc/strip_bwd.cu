// strip_bwd.cu — kernel (3): fused contrastive backward.
//
// Replaces the autograd backward of cn_clip/training/train.py:87-115.  For a block of 128 local
// rows A (image or text features) and every 128-column tile B_j of the gathered other modality:
//   MMA1 : S_j = A * B_j^T                      (recomputed logits / s, fp32 in TMEM; SS, N = 128)
//   warps: G_j = 2^12 * ( exp(s S - lse_row) + exp(s S - lse_col[j]) - 2 delta_label )  -> 16 bit,
//          written back into TMEM over S_j (tcgen05.st), never to shared or global memory.
//          One MUFU per element instead of two: exp(sS-lr) + exp(sS-lc) = exp(sS-lr) * (1 + a_i b_j),
//          a_i = 2^(lr_i - mu), b_j = 2^(mu - lc_j); b_j is produced once per tile by warp 3.
//          (Used when all lse lie within 2^+-50 of mu; otherwise the two-exp form runs.)
//   MMA2 : dA[:, slice] += G_j * B_j[:, slice]  (A operand from TMEM; B_j streamed a second time
//          from L2 and read as an MN-major operand: rows of B_j are the K dimension)
// The fp32 dA accumulator for 128 rows x D does not fit TMEM next to S (D = 512 alone is all 512
// columns), so a unit (= one CTA) owns a 256-feature slice of the output and the logits are
// recomputed once per slice ("pass").  unit = (strip, row block, pass, column split).
//
// Shapes follow the measured cost of tcgen05.mma on B200 (tools/mma_bench.cu, M = 128, K = 16):
//   SS (A from smem): 68.5 clk for N <= 64, 74.8 for N = 128, 138.8 for N = 256 (the A tile is
//   read from shared memory at 64 B/clk);  TS (A from TMEM): 44.3 / 64.0 / 128.0.
// so S uses N = 128 tiles and the gradient contraction runs as TS with N = 128.
//
// TMEM: [0,256) dA slice | [256,384) S buffer 0 | [384,512) S buffer 1; G_j overwrites columns
//       [0,64) of its own S buffer (two 16-bit values per column).
// SMEM: A block resident (D <= 512), a ring of stages each holding a PAIR of 64-feature chunks of
//       a B tile (2 x 16 KB; + 2 x 16 KB of A when A is streamed).  One mbarrier per stage = 8
//       MMAs per wait, which keeps the single issuing thread off the critical path.
//
// Roofline: tensor cores.  Algorithmic flops = 4 * rows * N * D per strip (S recompute excluded).
#include <stdlib.h>

#include "common.cuh"

namespace nans {
namespace {

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

struct BwdPlan {
  bool a_resident;
  int nr;           // ring stages
  size_t bytes;
};

BwdPlan plan_bwd(int kchunks) {
  BwdPlan p;
  const size_t cap = SMEM_CAP - 1024 - BAR_BYTES;
  const size_t a_res = static_cast<size_t>(kchunks) * A_CHUNK;
  const size_t stage_b = static_cast<size_t>(PAIR) * B_CHUNK;
  const size_t stage_ab = static_cast<size_t>(PAIR) * (A_CHUNK + B_CHUNK);
  if (a_res + 2 * stage_b <= cap) {
    p.a_resident = true;
    p.nr = static_cast<int>((cap - a_res) / stage_b);
    if (p.nr > MAX_NR) p.nr = MAX_NR;
    p.bytes = a_res + static_cast<size_t>(p.nr) * stage_b + BAR_BYTES + 1024;
  } else {
    p.a_resident = false;
    p.nr = static_cast<int>(cap / stage_ab);
    if (p.nr > MAX_NR) p.nr = MAX_NR;
    p.bytes = static_cast<size_t>(p.nr) * stage_ab + BAR_BYTES + 1024;
  }
  return p;
}

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
};

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

template <bool A_RES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
clip_bwd_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const BwdParams p) {
  constexpr int STAGE = A_RES ? PAIR * B_CHUNK : PAIR * (A_CHUNK + B_CHUNK);
  constexpr int STAGE_B_OFF = A_RES ? 0 : PAIR * A_CHUNK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;  // resident block (A_RES only)
  uint8_t* smR = smA + (A_RES ? static_cast<size_t>(p.kchunks) * A_CHUNK : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smR + static_cast<size_t>(p.nr) * STAGE);
  uint64_t* fullR = bars;
  uint64_t* emptyR = fullR + MAX_NR;
  uint64_t* a_full = emptyR + MAX_NR;
  uint64_t* s_full = a_full + 1;   // [2]
  uint64_t* g_ready = s_full + 2;  // [2]
  uint64_t* da_full = g_ready + 2;
  uint64_t* b_full = da_full + 1;   // [2] column factors of a tile are in shared memory
  uint64_t* b_empty = b_full + 2;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_empty + 2);
  float* cfbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + CF_OFF);  // [2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- unit decode ----
  int unit = blockIdx.x;
  const int split = unit % p.nsplit;
  unit /= p.nsplit;
  const int pass = unit % p.npass;
  unit /= p.npass;
  const int rb = unit % p.nrb;
  const int strip = unit / p.nrb;

  const CUtensorMap* tmA = strip == 0 ? &tmA0 : &tmA1;
  const CUtensorMap* tmB = strip == 0 ? &tmB0 : &tmB1;
  const int row0 = p.row_begin + rb * BM;
  const int tile_begin = static_cast<int>(static_cast<long long>(split) * p.ntiles / p.nsplit);
  const int tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / p.nsplit);
  const int ntiles = tile_end - tile_begin;
  const int npairs = (p.kchunks + PAIR - 1) / PAIR;            // ring stages per MMA1 sweep
  const int slice_c0 = pass * (SLICE / BK);                    // first chunk of the slice
  const int slice_nc = min(SLICE / BK, p.kchunks - slice_c0);  // chunks in the slice (1..4)
  const int slice_np = (slice_nc + PAIR - 1) / PAIR;           // ring stages per MMA2 sweep
  const int slice_w = slice_nc * BK;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(tmA);
      tma_prefetch_desc(tmB);
      for (int i = 0; i < MAX_NR; ++i) { mbar_init(&fullR[i], 1); mbar_init(&emptyR[i], 1); }
      mbar_init(a_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&g_ready[i], SM_WARPS); }
      mbar_init(da_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], SM_WARPS); }
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // one-exp formulation if every lse lies within +-kFactorRange/2 of mu (uniform over the grid)
  float lse_mu;
  bool factored;
  {
    int lo = __ldg(p.lse_minmax), hi = __ldg(p.lse_minmax + 1);
    lo = lo >= 0 ? lo : lo ^ 0x7fffffff;
    hi = hi >= 0 ? hi : hi ^ 0x7fffffff;
    const float fmin = __int_as_float(lo), fmax = __int_as_float(hi);
    factored = (fmax - fmin) < kFactorRange;
    lse_mu = 0.5f * (fmax + fmin);
  }

  // Both the producer and the MMA issuer walk the same schedule:
  //   step tau = 0 .. ntiles :  [tau < ntiles] MMA1 stages of tile tau ; [tau >= 1] MMA2 stages of tile tau-1
  if (warp == 0) {
    // ---------------- TMA producer (warp-uniform loop, one elected lane issues) ----------------
    if (A_RES) {
      if (elect_one()) {
        mbar_arrive_expect_tx(a_full, static_cast<uint32_t>(p.kchunks) * A_CHUNK);
        for (int c = 0; c < p.kchunks; ++c)
          tma_load_2d(smA + static_cast<size_t>(c) * A_CHUNK, tmA, a_full, c * BK, row0);
      }
      __syncwarp();
    }
    int sr = 0;
    uint32_t pr = 0;
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const int col0 = (tile_begin + tau) * KT;
        for (int j = 0; j < npairs; ++j) {
          const int nck = min(PAIR, p.kchunks - j * PAIR);
          mbar_wait(&emptyR[sr], pr ^ 1u);
          if (elect_one()) {
            uint8_t* st = smR + static_cast<size_t>(sr) * STAGE;
            mbar_arrive_expect_tx(&fullR[sr], static_cast<uint32_t>(nck) * (B_CHUNK + (A_RES ? 0 : A_CHUNK)));
            for (int ci = 0; ci < nck; ++ci) {
              const int f = (j * PAIR + ci) * BK;
              if (!A_RES) tma_load_2d(st + ci * A_CHUNK, tmA, &fullR[sr], f, row0);
              tma_load_2d(st + STAGE_B_OFF + ci * B_CHUNK, tmB, &fullR[sr], f, col0);
            }
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1 && !(p.debug & 2)) {
        const int col0 = (tile_begin + tau - 1) * KT;
        for (int j = 0; j < slice_np; ++j) {
          const int nck = min(PAIR, slice_nc - j * PAIR);
          mbar_wait(&emptyR[sr], pr ^ 1u);
          if (elect_one()) {
            uint8_t* st = smR + static_cast<size_t>(sr) * STAGE;
            mbar_arrive_expect_tx(&fullR[sr], static_cast<uint32_t>(nck) * B_CHUNK);
            for (int ci = 0; ci < nck; ++ci)
              tma_load_2d(st + STAGE_B_OFF + ci * B_CHUNK, tmB, &fullR[sr], (slice_c0 + j * PAIR + ci) * BK, col0);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (warp-uniform loop, one elected lane issues) ----------------
    const uint32_t fmt = p.idesc1_fmt;
    const uint32_t idesc1 = make_idesc(fmt, fmt, 0, 0, BM, KT);
    const uint32_t smA_addr = smem_u32(smA), smR_addr = smem_u32(smR);
    if (A_RES) {
      mbar_wait(a_full, 0);
      tc_fence_after();
    }
    int sr = 0;
    uint32_t pr = 0;
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const uint32_t d_S = tmem_base + TMEM_S + (tau & 1) * KT;
        for (int j = 0; j < npairs; ++j) {
          const int nck = min(PAIR, p.kchunks - j * PAIR);
          mbar_wait(&fullR[sr], pr);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st = smR_addr + static_cast<uint32_t>(sr) * STAGE;
            const uint64_t ad0 = make_smem_desc(A_RES ? smA_addr + static_cast<uint32_t>(j * PAIR) * A_CHUNK : st, 16, 1024);
            const uint64_t bd0 = make_smem_desc(st + STAGE_B_OFF, 16, 1024);
            for (int ci = 0; ci < nck; ++ci) {
              const uint64_t ad = ad0 + static_cast<uint64_t>(ci * (A_CHUNK >> 4));
              const uint64_t bd = bd0 + static_cast<uint64_t>(ci * (B_CHUNK >> 4));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                mma_ss(d_S, ad + 2 * k, bd + 2 * k, idesc1, (j | ci | k) != 0 ? 1u : 0u);
            }
            tc_commit(&emptyR[sr]);
            if (j == npairs - 1) tc_commit(&s_full[tau & 1]);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1) {
        const int u = tau - 1;
        const int gb = u & 1;
        mbar_wait(&g_ready[gb], static_cast<uint32_t>(u >> 1) & 1u);
        tc_fence_after();
        const uint32_t g_tmem = tmem_base + TMEM_S + gb * KT;
        if (p.debug & 2) {
          if (tau == ntiles) { if (elect_one()) tc_commit(da_full); __syncwarp(); }
          continue;
        }
        for (int j = 0; j < slice_np; ++j) {
          const int nck = min(PAIR, slice_nc - j * PAIR);
          mbar_wait(&fullR[sr], pr);
          tc_fence_after();
          if (elect_one()) {
            // B_j[:, 64-feature blocks] as an MN-major operand: 128-byte lines = 64 features of one
            // row, lines step K (rows), 8-row groups 1024 B apart, feature blocks one chunk apart
            const uint32_t idesc2 = make_idesc(p.g_fmt, fmt, 0, 1, BM, nck * BK);
            const uint64_t bd = make_smem_desc(smR_addr + static_cast<uint32_t>(sr) * STAGE + STAGE_B_OFF, B_CHUNK, 1024);
            const uint32_t d_dA = tmem_base + static_cast<uint32_t>(j * PAIR * BK);
#pragma unroll
            for (int kk = 0; kk < KT / 16; ++kk)  // 16 rows x 128 B = 2048 B (>> 4 = 128) per K step
              mma_ts(d_dA, g_tmem + kk * 8, bd + 128 * kk, idesc2, (u > 0 || kk > 0) ? 1u : 0u);
            tc_commit(&emptyR[sr]);
            if (tau == ntiles && j == slice_np - 1) tc_commit(da_full);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ---------------- column-factor warp: per tile, b_j = 2^(mu - lc_j) (or lc_j - 12) -> smem ----
    const float* lse_col = p.lse_col[strip];
    for (int t = 0; t < ntiles; ++t) {
      const int bb = t & 1;
      mbar_wait(&b_empty[bb], (static_cast<uint32_t>(t >> 1) & 1u) ^ 1u);
      const int cb = (tile_begin + t) * KT + lane * 4;
      float v[4];
      if (cb + 4 <= p.ncols) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(lse_col + cb));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(lse_col + max(min(cb + k, p.ncols - 1), 0));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = factored ? fast_exp2(lse_mu - v[k]) : v[k] - kGShiftLog2;
      *reinterpret_cast<float4*>(cfbuf + bb * KT + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[bb]);  // mbarrier arrive has release semantics (cta scope)
    }
  } else if (warp >= 4) {
    // ---------------- softmax-gradient warps: thread = (row, 64-column half of the tile) ---------
    const int wq = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const int row = row0 + wq * 32 + lane;
    const bool valid = row < p.row_end;
    const float s = __ldg(p.s_dev);
    const float c = s * kLog2e;
    // 2^12 * exp(s cos - lse) = 2^(cos * c - (lse2 - 12)), lse2 = base-2 lse from the forward
    const float lr2 = valid ? __ldg(p.lse_row[strip] + row) - kGShiftLog2 : INFINITY;
    // a_i = 2^(lr_i - mu); rows past the end get 0 (their e1 is 0 as well)
    const float a_i = valid ? fast_exp2(__ldg(p.lse_row[strip] + row) - lse_mu) : 0.f;
    const int label = row + p.label_shift;
    const int warp_label_lo = label - lane;
    const bool g_bf16 = p.g_fmt != 0;

    for (int t = 0; t < ntiles; ++t) {
      const int sb = t & 1;
      const int cb0 = (tile_begin + t) * KT + half * 64;
      uint32_t g16[32];
      mbar_wait(&b_full[sb], static_cast<uint32_t>(t >> 1) & 1u);
      mbar_wait(&s_full[sb], static_cast<uint32_t>(t >> 1) & 1u);
      tc_fence_after();
      const float* cf = cfbuf + sb * KT + half * 64;
      if (p.debug & 4) {
        __syncwarp();
        if (lane == 0) { mbar_arrive(&b_empty[sb]); }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_ready[sb]);
        continue;
      }
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int cb = cb0 + sub * 32;
        const bool has_label = (warp_label_lo < cb + 32) && (warp_label_lo + 31 >= cb);
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + TMEM_S + sb * KT + half * 64 + sub * 32, r);
        tmem_wait_ld();
        const float* cfs = cf + sub * 32;
        const int label_rel = label - cb;  // in [0,32) only for the thread whose label is here
        uint32_t* go = g16 + sub * 16;
        if (p.debug & 1) {
#pragma unroll
          for (int k = 0; k < 16; ++k) go[k] = r[2 * k];
        } else if (factored) {
          if (g_bf16) softmax_grad32_dispatch<true, true>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
          else softmax_grad32_dispatch<true, false>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
        } else {
          if (g_bf16) softmax_grad32_dispatch<false, true>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
          else softmax_grad32_dispatch<false, false>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[sb]);
      // G (32 columns per half) overwrites S columns [0,64) of this buffer; the half-1 thread of a
      // row writes columns [32,64), which the half-0 thread of the same row has just read as S:
      // the two warps of a lane group meet before any of them stores.
      asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
      {
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo[i] = g16[i]; hi[i] = g16[16 + i]; }
        const uint32_t g_addr = tmem_base + lane_base + TMEM_S + sb * KT + half * 32;
        tmem_st16(g_addr, lo);
        tmem_st16(g_addr + 16, hi);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&g_ready[sb]);
    }

    // ---- write the dA slice: the two halves take alternate 32-column chunks ----
    mbar_wait(da_full, 0);
    tc_fence_after();
    const float coef = __ldg(p.grad_out_dev) * s * p.coef_host;
    float* out = p.out[strip] + static_cast<long long>(row - p.row_begin) * p.D + pass * SLICE;
    for (int ch = half; ch < slice_w / 32; ch += 2) {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_base + ch * 32, r);
      tmem_wait_ld();
      if (valid) {
        const int f0 = pass * SLICE + ch * 32;
        if (f0 + 32 <= p.D) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float a0 = __uint_as_float(r[k]) * coef, a1 = __uint_as_float(r[k + 1]) * coef;
            const float a2 = __uint_as_float(r[k + 2]) * coef, a3 = __uint_as_float(r[k + 3]) * coef;
            if (p.accumulate) red_add_v4(out + ch * 32 + k, a0, a1, a2, a3);
            else *reinterpret_cast<float4*>(out + ch * 32 + k) = make_float4(a0, a1, a2, a3);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (f0 + k < p.D) {
              if (p.accumulate) atomicAdd(out + ch * 32 + k, __uint_as_float(r[k]) * coef);
              else out[ch * 32 + k] = __uint_as_float(r[k]) * coef;
            }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================
// CTA-pair version (cta_group::2).  Two CTAs of a cluster own 256 consecutive rows (128 each) and
// share every tcgen05.mma (M = 256): each CTA stages only HALF of each B tile, which halves the
// shared-memory traffic per MMA.  That traffic is what bounds the single-CTA kernel: an SS MMA with
// N = 128 reads 4 KB of A + 4 KB of B and TMA writes another 4 KB per MMA against 128 B/clk of
// shared memory (96 clk for 64 clk of math; 115 measured).  In the pair a CTA moves 4 + 2 + 2 KB.
// Everything per-CTA is as in the single-CTA kernel (own A block, own TMEM with the dA slice, the S
// buffers and G, own softmax and column-factor warps); only the leader (cluster rank 0) issues MMAs:
//   * TMA loads of both CTAs credit the LEADER's full barrier (tma_load_2d_pair);
//   * the leader's commits are multicast to both CTAs' empty / s_full / da_full barriers;
//   * all 16 softmax warps arrive on the leader's g_ready barrier (remote mbarrier arrive).
// MMA1: B K-major, N = 128 = 64 tile rows from each CTA -> S columns [0,64) | [64,128).
// MMA2: B MN-major, N = 128 features per stage = 64 from each CTA.
constexpr int P_BHALF = (KT / 2) * BK * 2;  // 8 KB: this CTA's half of a 64-feature chunk of a B tile
constexpr int P_STAGE_B = 2 * P_BHALF;      // 16 KB: MMA1: two chunk halves | MMA2: one [128 x 64] chunk

struct PairPlan {
  bool a_resident;
  int nr;
  size_t bytes;
};

PairPlan plan_pair(int kchunks) {
  PairPlan p;
  const size_t cap = SMEM_CAP - 1024 - BAR_BYTES;
  const size_t a_res = static_cast<size_t>(kchunks) * A_CHUNK;
  if (a_res + 3 * static_cast<size_t>(P_STAGE_B) <= cap) {
    p.a_resident = true;
    p.nr = static_cast<int>((cap - a_res) / P_STAGE_B);
    if (p.nr > MAX_NR) p.nr = MAX_NR;
    p.bytes = a_res + static_cast<size_t>(p.nr) * P_STAGE_B + BAR_BYTES + 1024;
  } else {
    const size_t st = static_cast<size_t>(PAIR) * A_CHUNK + P_STAGE_B;
    p.a_resident = false;
    p.nr = static_cast<int>(cap / st);
    if (p.nr > MAX_NR) p.nr = MAX_NR;
    p.bytes = static_cast<size_t>(p.nr) * st + BAR_BYTES + 1024;
  }
  return p;
}

template <bool A_RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
clip_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmBk0,
                     const __grid_constant__ CUtensorMap tmBm0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmBk1, const __grid_constant__ CUtensorMap tmBm1,
                     const BwdParams p) {
  constexpr int STAGE = A_RES ? P_STAGE_B : (PAIR * A_CHUNK + P_STAGE_B);
  constexpr int STAGE_B_OFF = A_RES ? 0 : PAIR * A_CHUNK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;
  uint8_t* smR = smA + (A_RES ? static_cast<size_t>(p.kchunks) * A_CHUNK : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smR + static_cast<size_t>(p.nr) * STAGE);
  uint64_t* fullR = bars;           // used in the leader only
  uint64_t* emptyR = fullR + MAX_NR;
  uint64_t* a_full = emptyR + MAX_NR;  // leader only
  uint64_t* s_full = a_full + 1;       // [2]
  uint64_t* g_ready = s_full + 2;      // [2] leader only, 16 arrivals
  uint64_t* da_full = g_ready + 2;
  uint64_t* b_full = da_full + 1;      // [2]
  uint64_t* b_empty = b_full + 2;      // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_empty + 2);
  float* cfbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + CF_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  // ---- unit decode (one unit per cluster) ----
  int unit = blockIdx.x >> 1;
  const int split = unit % p.nsplit;
  unit /= p.nsplit;
  const int pass = unit % p.npass;
  unit /= p.npass;
  const int rb = unit % p.nrb;
  const int strip = unit / p.nrb;

  const CUtensorMap* tmA = strip == 0 ? &tmA0 : &tmA1;
  const CUtensorMap* tmBk = strip == 0 ? &tmBk0 : &tmBk1;  // box [64 rows, 64 features]
  const CUtensorMap* tmBm = strip == 0 ? &tmBm0 : &tmBm1;  // box [128 rows, 64 features]
  const int row0 = p.row_begin + rb * 2 * BM + static_cast<int>(rank) * BM;
  const int tile_begin = static_cast<int>(static_cast<long long>(split) * p.ntiles / p.nsplit);
  const int tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / p.nsplit);
  const int ntiles = tile_end - tile_begin;
  const int npairs = (p.kchunks + PAIR - 1) / PAIR;
  const int slice_c0 = pass * (SLICE / BK);
  const int slice_nc = min(SLICE / BK, p.kchunks - slice_c0);
  const int slice_np = (slice_nc + PAIR - 1) / PAIR;
  const int slice_w = slice_nc * BK;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(tmA);
      tma_prefetch_desc(tmBk);
      tma_prefetch_desc(tmBm);
      for (int i = 0; i < MAX_NR; ++i) { mbar_init(&fullR[i], 1); mbar_init(&emptyR[i], 1); }
      mbar_init(a_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&g_ready[i], 2 * SM_WARPS); }
      mbar_init(da_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], SM_WARPS); }
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc_pair(tmem_ptr, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  float lse_mu;
  bool factored;
  {
    int lo = __ldg(p.lse_minmax), hi = __ldg(p.lse_minmax + 1);
    lo = lo >= 0 ? lo : lo ^ 0x7fffffff;
    hi = hi >= 0 ? hi : hi ^ 0x7fffffff;
    const float fmin = __int_as_float(lo), fmax = __int_as_float(hi);
    factored = (fmax - fmin) < kFactorRange;
    lse_mu = 0.5f * (fmax + fmin);
  }

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs, same schedule) ----------------
    if (A_RES) {
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(a_full, 2u * static_cast<uint32_t>(p.kchunks) * A_CHUNK);
        for (int c = 0; c < p.kchunks; ++c)
          tma_load_2d_pair(smA + static_cast<size_t>(c) * A_CHUNK, tmA, a_full, c * BK, row0);
      }
      __syncwarp();
    }
    int sr = 0;
    uint32_t pr = 0;
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const int col0 = (tile_begin + tau) * KT + static_cast<int>(rank) * (KT / 2);
        for (int j = 0; j < npairs; ++j) {
          const int nck = min(PAIR, p.kchunks - j * PAIR);
          mbar_wait(&emptyR[sr], pr ^ 1u);
          if (elect_one()) {
            uint8_t* st = smR + static_cast<size_t>(sr) * STAGE;
            if (leader)
              mbar_arrive_expect_tx(&fullR[sr], 2u * static_cast<uint32_t>(nck) * (P_BHALF + (A_RES ? 0 : A_CHUNK)));
            for (int ci = 0; ci < nck; ++ci) {
              const int f = (j * PAIR + ci) * BK;
              if (!A_RES) tma_load_2d_pair(st + ci * A_CHUNK, tmA, &fullR[sr], f, row0);
              tma_load_2d_pair(st + STAGE_B_OFF + ci * P_BHALF, tmBk, &fullR[sr], f, col0);
            }
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1 && !(p.debug & 2)) {
        const int col0 = (tile_begin + tau - 1) * KT;
        for (int j = 0; j < slice_np; ++j) {
          mbar_wait(&emptyR[sr], pr ^ 1u);
          if (elect_one()) {
            uint8_t* st = smR + static_cast<size_t>(sr) * STAGE;
            if (leader) mbar_arrive_expect_tx(&fullR[sr], 2u * P_STAGE_B);
            // this CTA's 64 of the stage's 128 features (a chunk past D is zero-filled by TMA)
            tma_load_2d_pair(st + STAGE_B_OFF, tmBm, &fullR[sr], (slice_c0 + j * PAIR + static_cast<int>(rank)) * BK, col0);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ---------------- MMA issuer (leader CTA only) ----------------
    const uint32_t fmt = p.idesc1_fmt;
    const uint32_t idesc1 = make_idesc(fmt, fmt, 0, 0, 2 * BM, KT);
    const uint32_t idesc2 = make_idesc(p.g_fmt, fmt, 0, 1, 2 * BM, PAIR * BK);
    const uint32_t smA_addr = smem_u32(smA), smR_addr = smem_u32(smR);
    if (A_RES) {
      mbar_wait(a_full, 0);
      tc_fence_after();
    }
    int sr = 0;
    uint32_t pr = 0;
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const uint32_t d_S = tmem_base + TMEM_S + (tau & 1) * KT;
        for (int j = 0; j < npairs; ++j) {
          const int nck = min(PAIR, p.kchunks - j * PAIR);
          mbar_wait(&fullR[sr], pr);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st = smR_addr + static_cast<uint32_t>(sr) * STAGE;
            const uint64_t ad0 = make_smem_desc(A_RES ? smA_addr + static_cast<uint32_t>(j * PAIR) * A_CHUNK : st, 16, 1024);
            const uint64_t bd0 = make_smem_desc(st + STAGE_B_OFF, 16, 1024);
            for (int ci = 0; ci < nck; ++ci) {
              const uint64_t ad = ad0 + static_cast<uint64_t>(ci * (A_CHUNK >> 4));
              const uint64_t bd = bd0 + static_cast<uint64_t>(ci * (P_BHALF >> 4));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                mma_ss_pair(d_S, ad + 2 * k, bd + 2 * k, idesc1, (j | ci | k) != 0 ? 1u : 0u);
            }
            tc_commit_pair(&emptyR[sr], 3);
            if (j == npairs - 1) tc_commit_pair(&s_full[tau & 1], 3);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1) {
        const int u = tau - 1;
        const int gb = u & 1;
        mbar_wait(&g_ready[gb], static_cast<uint32_t>(u >> 1) & 1u);
        tc_fence_after();
        const uint32_t g_tmem = tmem_base + TMEM_S + gb * KT;
        if (p.debug & 2) {
          if (tau == ntiles) { if (elect_one()) tc_commit_pair(da_full, 3); __syncwarp(); }
          continue;
        }
        for (int j = 0; j < slice_np; ++j) {
          mbar_wait(&fullR[sr], pr);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bd = make_smem_desc(smR_addr + static_cast<uint32_t>(sr) * STAGE + STAGE_B_OFF, B_CHUNK, 1024);
            const uint32_t d_dA = tmem_base + static_cast<uint32_t>(j * PAIR * BK);
#pragma unroll
            for (int kk = 0; kk < KT / 16; ++kk)
              mma_ts_pair(d_dA, g_tmem + kk * 8, bd + 128 * kk, idesc2, (u > 0 || kk > 0) ? 1u : 0u);
            tc_commit_pair(&emptyR[sr], 3);
            if (tau == ntiles && j == slice_np - 1) tc_commit_pair(da_full, 3);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ---------------- column-factor warp (each CTA for itself) ----------------
    const float* lse_col = p.lse_col[strip];
    for (int t = 0; t < ntiles; ++t) {
      const int bb = t & 1;
      mbar_wait(&b_empty[bb], (static_cast<uint32_t>(t >> 1) & 1u) ^ 1u);
      const int cb = (tile_begin + t) * KT + lane * 4;
      float v[4];
      if (cb + 4 <= p.ncols) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(lse_col + cb));
        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = __ldg(lse_col + max(min(cb + k, p.ncols - 1), 0));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = factored ? fast_exp2(lse_mu - v[k]) : v[k] - kGShiftLog2;
      *reinterpret_cast<float4*>(cfbuf + bb * KT + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[bb]);
    }
  } else if (warp >= 4) {
    // ---------------- softmax-gradient warps: thread = (row, 64-column half of the tile) ---------
    const int wq = warp & 3;
    const int half = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(wq * 32) << 16;
    const int row = row0 + wq * 32 + lane;
    const bool valid = row < p.row_end;
    const float s = __ldg(p.s_dev);
    const float c = s * kLog2e;
    const float lr2 = valid ? __ldg(p.lse_row[strip] + row) - kGShiftLog2 : INFINITY;
    const float a_i = valid ? fast_exp2(__ldg(p.lse_row[strip] + row) - lse_mu) : 0.f;
    const int label = row + p.label_shift;
    const int warp_label_lo = label - lane;
    const bool g_bf16 = p.g_fmt != 0;

    for (int t = 0; t < ntiles; ++t) {
      const int sb = t & 1;
      const int cb0 = (tile_begin + t) * KT + half * 64;
      uint32_t g16[32];
      mbar_wait(&b_full[sb], static_cast<uint32_t>(t >> 1) & 1u);
      mbar_wait(&s_full[sb], static_cast<uint32_t>(t >> 1) & 1u);
      tc_fence_after();
      const float* cf = cfbuf + sb * KT + half * 64;
      if (p.debug & 4) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_empty[sb]);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&g_ready[sb], 0);
        continue;
      }
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        const int cb = cb0 + sub * 32;
        const bool has_label = (warp_label_lo < cb + 32) && (warp_label_lo + 31 >= cb);
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + TMEM_S + sb * KT + half * 64 + sub * 32, r);
        tmem_wait_ld();
        const float* cfs = cf + sub * 32;
        const int label_rel = label - cb;  // in [0,32) only for the thread whose label is here
        uint32_t* go = g16 + sub * 16;
        if (p.debug & 1) {
#pragma unroll
          for (int k = 0; k < 16; ++k) go[k] = r[2 * k];
        } else if (factored) {
          if (g_bf16) softmax_grad32_dispatch<true, true>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
          else softmax_grad32_dispatch<true, false>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
        } else {
          if (g_bf16) softmax_grad32_dispatch<false, true>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
          else softmax_grad32_dispatch<false, false>(has_label, r, cfs, c, lr2, a_i, label_rel, go);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[sb]);
      asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
      {
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo[i] = g16[i]; hi[i] = g16[16 + i]; }
        const uint32_t g_addr = tmem_base + lane_base + TMEM_S + sb * KT + half * 32;
        tmem_st16(g_addr, lo);
        tmem_st16(g_addr + 16, hi);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&g_ready[sb], 0);  // the leader issues MMA2 for both CTAs
    }

    mbar_wait(da_full, 0);
    tc_fence_after();
    const float coef = __ldg(p.grad_out_dev) * s * p.coef_host;
    float* out = p.out[strip] + static_cast<long long>(row - p.row_begin) * p.D + pass * SLICE;
    for (int ch = half; ch < slice_w / 32; ch += 2) {
      uint32_t r[32];
      tmem_ld32(tmem_base + lane_base + ch * 32, r);
      tmem_wait_ld();
      if (valid) {
        const int f0 = pass * SLICE + ch * 32;
        if (f0 + 32 <= p.D) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float a0 = __uint_as_float(r[k]) * coef, a1 = __uint_as_float(r[k + 1]) * coef;
            const float a2 = __uint_as_float(r[k + 2]) * coef, a3 = __uint_as_float(r[k + 3]) * coef;
            if (p.accumulate) red_add_v4(out + ch * 32 + k, a0, a1, a2, a3);
            else *reinterpret_cast<float4*>(out + ch * 32 + k) = make_float4(a0, a1, a2, a3);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (f0 + k < p.D) {
              if (p.accumulate) atomicAdd(out + ch * 32 + k, __uint_as_float(r[k]) * coef);
              else out[ch * 32 + k] = __uint_as_float(r[k]) * coef;
            }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while its partner may still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================
// Narrow CTA-pair version (D <= 512): the pair owns 128 rows, 64 per CTA (cta_group::2, M = 128).
// With 64 rows per CTA the fp32 dA for ALL D <= 512 features takes 256 TMEM columns (the M = 128
// pair layout puts the two N halves of an accumulator on lanes 0-63 / 64-127), which leaves room for
// two S buffers of a 256-column tile: the logits are recomputed ONCE per tile instead of once per
// 256-feature slice.  The kernel is bound by the shared-memory data pipe (tensor-core operand reads
// + the softmax warps' TMEM / shared traffic, ncu: 98 % with 128-column tiles), so MMA1 uses the
// widest N (256: A is re-read once per 256 columns) and the softmax warps use ld/st.shared.
//   TMEM : [0,128) dA features 0-255 | [128,256) dA features 256-511 | [256,512) S buffers 0-1
//          S / dA rows 0-63 of the CTA sit on lanes 0-63 (first N half) and 64-127 (second N half).
//   SMEM : A (64 rows, resident) | 2 G buffers (64 rows x 256 K, 16-bit, K-major swizzled: G goes
//          through shared memory because the TMEM-A form of a pair MMA wants a duplicated layout)
//          | ring of 32 KB stages: MMA1 = 2 chunks of [128 tile rows x 64 features] of this CTA's
//          half of the tile; MMA2 = [128 tile rows x 64 features] x 2 of this CTA's 128 features of
//          a 256-feature block.  MMA2 of tile t is issued after MMA1 of tile t + 1.
constexpr int NP_ROWS = 64;
constexpr int NP_KT = 256;
constexpr int NP_ACH = NP_ROWS * BK * 2;      // 8 KB  : one 64-feature chunk of this CTA's A rows / of G
constexpr int NP_STAGE = 2 * B_CHUNK;         // 32 KB
constexpr int NP_STAGE_STREAM = 3 * B_CHUNK;  // 48 KB: two MMA1 chunks of [64 A rows | 128 tile rows] x 64 features
constexpr int NP_GBUF = NP_ROWS * NP_KT * 2;  // 32 KB
constexpr int NP_NG = 2;
constexpr int NP_NS = 2;                      // S buffers (128 TMEM columns each)
constexpr int NP_MAXR = 6;
constexpr int NP_BAR_BYTES = 2560;            // mbarriers + TMEM pointer (256 B) + two 256-float column-factor buffers
constexpr int NP_CF_OFF = 256;

struct NpPlan {
  int nr;
  size_t bytes;
};
NpPlan plan_np(int kchunks, int tk = NP_KT, bool a_resident = true) {
  NpPlan p;
  const size_t NP_STAGE = a_resident ? nans::NP_STAGE : nans::NP_STAGE_STREAM;  // shadows the constant below
  const size_t fixed = (a_resident ? static_cast<size_t>(kchunks) * NP_ACH : 0) +
                       static_cast<size_t>(NP_NG) * NP_ROWS * tk * 2 + NP_BAR_BYTES;
  // the 1 KB of alignment slack is dropped when it would cost a ring stage (D = 512): the kernel
  // checks that its aligned carve-up fits and traps otherwise
  size_t pad = 1024;
  p.nr = static_cast<int>((SMEM_CAP - pad - fixed) / NP_STAGE);
  if (p.nr < 3 && (SMEM_CAP - fixed) / NP_STAGE >= 3) {
    p.nr = 3;
    pad = SMEM_CAP - fixed - 3 * static_cast<size_t>(NP_STAGE);
  }
  if (p.nr > NP_MAXR) p.nr = NP_MAXR;
  p.bytes = fixed + static_cast<size_t>(p.nr) * NP_STAGE + pad;
  return p;
}

__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// softmax_grad32 for the narrow-pair kernels (column factors through one load + warp shuffles)
template <bool FACTORED, bool BF16, bool LABEL>
__device__ __forceinline__ void np_grad32(const uint32_t (&r)[32], uint32_t cf_addr, float c, float lr2,
                                          float a_i, int label_rel, uint32_t* __restrict__ g16) {
  // the 32 column factors of this group: ONE lane-distributed load + 32 shuffles.  Eight broadcast
  // LDS.128 per thread cost two wavefronts each on the shared-memory data pipe, which is what bounds
  // this kernel (tensor-core operand reads share it); the shuffles do not (3.11 -> 2.98 ms).
  float cfl;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(cfl) : "r"(cf_addr + 4 * (threadIdx.x & 31)));
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float cfv[4] = {__shfl_sync(0xffffffffu, cfl, 4 * q), __shfl_sync(0xffffffffu, cfl, 4 * q + 1),
                          __shfl_sync(0xffffffffu, cfl, 4 * q + 2), __shfl_sync(0xffffffffu, cfl, 4 * q + 3)};
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
__device__ __forceinline__ void np_grad32_dispatch(bool has_label, const uint32_t (&r)[32], uint32_t cf_addr,
                                                   float c, float lr2, float a_i, int label_rel,
                                                   uint32_t* __restrict__ g16) {
  if (has_label) np_grad32<FACTORED, BF16, true>(r, cf_addr, c, lr2, a_i, label_rel, g16);
  else np_grad32<FACTORED, BF16, false>(r, cf_addr, c, lr2, a_i, label_rel, g16);
}

// TK = tile width in columns: 256 for D <= 512 (two 128-column S buffers beside 256 columns of dA),
// 128 for 512 < D <= 768 (dA takes 384 TMEM columns, two 64-column S buffers are left).
//   tmS* : column operand for MMA1, box [TK/2 rows, 64 features] (this CTA's half of the tile)
//   tmBm*: column operand for MMA2, box [128 rows, 64 features]
// ARES = false (768 < D <= 1024): A does not stay resident (128 KB) — its chunks travel through the
// ring with the column chunks of MMA1 — and dA is produced in two passes of 512 features (256 TMEM
// columns each); a unit is then (strip, row block, pass) and S is recomputed once per pass.
template <int TK, bool ARES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
clip_bwd_np_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmS0,
                   const __grid_constant__ CUtensorMap tmBm0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmS1, const __grid_constant__ CUtensorMap tmBm1,
                   const BwdParams p) {
  constexpr int SW = TK / 2;                       // TMEM columns of one S buffer
  constexpr int GBUF = NP_ROWS * TK * 2;           // bytes of one G buffer (TK / 64 chunks of 8 KB)
  constexpr int S_CH = (TK / 2) * BK * 2;          // MMA1: this CTA's rows of a tile, one 64-feature chunk
  constexpr int STG = ARES ? NP_STAGE : NP_STAGE_STREAM;  // ring stage bytes
  constexpr int CPS = ARES ? STG / S_CH : STG / (S_CH + NP_ACH);  // MMA1 chunks per ring stage
  constexpr int MCH = ARES ? S_CH : S_CH + NP_ACH;  // bytes of one MMA1 chunk in a stage ([A rows |] tile rows)
  constexpr int KH = TK / 128;                     // MMA2: 128-row K halves per tile
  constexpr int TPT = TK / 4;                      // tile columns per softmax thread (64 or 32)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;
  uint8_t* smG = smA + (ARES ? static_cast<size_t>(p.kchunks) * NP_ACH : 0);
  uint8_t* smR = smG + static_cast<size_t>(NP_NG) * GBUF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smR + static_cast<size_t>(p.nr) * STG);
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (reinterpret_cast<uint8_t*>(bars) + NP_BAR_BYTES > smem_raw + dyn) __trap();  // carve-up does not fit
  }
  uint64_t* fullR = bars;                 // leader only
  uint64_t* emptyR = fullR + NP_MAXR;
  uint64_t* a_full = emptyR + NP_MAXR;    // leader only
  uint64_t* s_full = a_full + 1;          // [NP_NS]
  uint64_t* g_ready = s_full + NP_NS;     // [NP_NS] leader only, 16 arrivals
  uint64_t* g_empty = g_ready + NP_NS;    // [NP_NG]
  uint64_t* da_full = g_empty + NP_NG;
  uint64_t* b_full = da_full + 1;         // [2]
  uint64_t* b_empty = b_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_empty + 2);
  float* cfbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + NP_CF_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  // units [0, n_full) fill whole waves of CTA pairs and sweep all columns with plain stores; the
  // remaining units (a partial wave) are split ns_tail ways by columns and accumulate with red.add
  int unit = blockIdx.x >> 1;
  int split = 0, nsplit_u = 1;
  if (unit >= p.n_full) {
    const int j = unit - p.n_full;
    unit = p.n_full + j / p.ns_tail;
    split = j % p.ns_tail;
    nsplit_u = p.ns_tail;
  }
  const bool accumulate = nsplit_u > 1;
  const int pass = unit % p.npass;
  unit /= p.npass;
  const int rb = unit % p.nrb;
  const int strip = unit / p.nrb;
  const int f_off = pass * 2 * SLICE;  // first feature of this pass's 512-feature slice of dA

  const CUtensorMap* tmA = strip == 0 ? &tmA0 : &tmA1;     // box [64 rows, 64 features]
  const CUtensorMap* tmS = strip == 0 ? &tmS0 : &tmS1;
  const CUtensorMap* tmBm = strip == 0 ? &tmBm0 : &tmBm1;
  const int row0 = p.row_begin + rb * 2 * NP_ROWS + static_cast<int>(rank) * NP_ROWS;
  const int tile_begin = static_cast<int>(static_cast<long long>(split) * p.ntiles / nsplit_u);
  const int tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / nsplit_u);
  const int ntiles = tile_end - tile_begin;
  const int n1 = (p.kchunks + CPS - 1) / CPS;  // MMA1 stages per tile
  const uint32_t tmem_s = static_cast<uint32_t>(512 - NP_NS * SW);  // S buffers sit at the top of TMEM
  const int nfb = min((p.D - f_off + SLICE - 1) / SLICE, ARES ? 3 : 2);  // 256-feature blocks of dA in this pass

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(tmA);
      tma_prefetch_desc(tmBm);
      tma_prefetch_desc(tmS);
      for (int i = 0; i < NP_MAXR; ++i) { mbar_init(&fullR[i], 1); mbar_init(&emptyR[i], 1); }
      mbar_init(a_full, 1);
      for (int i = 0; i < NP_NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&g_ready[i], 2 * SM_WARPS); }
      for (int i = 0; i < NP_NG; ++i) mbar_init(&g_empty[i], 1);
      mbar_init(da_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], SM_WARPS); }
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc_pair(tmem_ptr, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  float lse_mu;
  bool factored;
  {
    int lo = __ldg(p.lse_minmax), hi = __ldg(p.lse_minmax + 1);
    lo = lo >= 0 ? lo : lo ^ 0x7fffffff;
    hi = hi >= 0 ? hi : hi ^ 0x7fffffff;
    const float fmin = __int_as_float(lo), fmax = __int_as_float(hi);
    factored = (fmax - fmin) < kFactorRange;
    lse_mu = 0.5f * (fmax + fmin);
  }

  // schedule shared by producer and issuer: step tau: [tau < ntiles] MMA1(tau); [tau >= 1] MMA2(tau - 1)
  if (warp == 0) {
    if (ARES) {
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(a_full, 2u * static_cast<uint32_t>(p.kchunks) * NP_ACH);
        for (int c = 0; c < p.kchunks; ++c)
          tma_load_2d_pair(smA + static_cast<size_t>(c) * NP_ACH, tmA, a_full, c * BK, row0);
      }
      __syncwarp();
    }
    int sr = 0;
    uint32_t pr = 0;
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const int col0 = (tile_begin + tau) * TK + static_cast<int>(rank) * (TK / 2);
        for (int j = 0; j < n1; ++j) {
          const int nck = min(CPS, p.kchunks - CPS * j);
          mbar_wait_parked(&emptyR[sr], pr ^ 1u);
          if (elect_one()) {
            uint8_t* st = smR + static_cast<size_t>(sr) * STG;
            if (leader) mbar_arrive_expect_tx(&fullR[sr], 2u * static_cast<uint32_t>(nck) * MCH);
            for (int ci = 0; ci < nck; ++ci) {
              if (!ARES) tma_load_2d_pair(st + ci * MCH, tmA, &fullR[sr], (CPS * j + ci) * BK, row0);
              tma_load_2d_pair(st + ci * MCH + (ARES ? 0 : NP_ACH), tmS, &fullR[sr], (CPS * j + ci) * BK, col0);
            }
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1) {
        const int col0 = (tile_begin + tau - 1) * TK;
        for (int kh = 0; kh < KH; ++kh) {
          for (int fb = 0; fb < nfb; ++fb) {
            mbar_wait_parked(&emptyR[sr], pr ^ 1u);
            if (elect_one()) {
              uint8_t* st = smR + static_cast<size_t>(sr) * STG;
              if (leader) mbar_arrive_expect_tx(&fullR[sr], 2u * 2u * B_CHUNK);
              // this CTA's 128 of the block's 256 features: chunks 4 fb + 2 rank, + 1 (zero fill past D)
              for (int ci = 0; ci < 2; ++ci)
                tma_load_2d_pair(st + ci * B_CHUNK, tmBm, &fullR[sr],
                                 f_off + (4 * fb + 2 * static_cast<int>(rank) + ci) * BK, col0 + kh * 128);
            }
            __syncwarp();
            if (++sr == p.nr) { sr = 0; pr ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    const uint32_t fmt = p.idesc1_fmt;
    const uint32_t idesc1 = make_idesc(fmt, fmt, 0, 0, 2 * NP_ROWS, TK);
    const uint32_t idesc2 = make_idesc(p.g_fmt, fmt, 0, 1, 2 * NP_ROWS, SLICE);
    const uint32_t smA_addr = smem_u32(smA), smR_addr = smem_u32(smR), smG_addr = smem_u32(smG);
    if (ARES) {
      mbar_wait(a_full, 0);
      tc_fence_after();
    }
    int sr = 0;
    uint32_t pr = 0;
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const uint32_t d_S = tmem_base + tmem_s + static_cast<uint32_t>((tau % NP_NS) * SW);
        for (int j = 0; j < n1; ++j) {
          const int nck = min(CPS, p.kchunks - CPS * j);
          mbar_wait(&fullR[sr], pr);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t st_addr = smR_addr + static_cast<uint32_t>(sr) * STG;
            const uint64_t ad0 = ARES ? make_smem_desc(smA_addr + static_cast<uint32_t>(CPS * j) * NP_ACH, 16, 1024)
                                      : make_smem_desc(st_addr, 16, 1024);
            const uint64_t bd0 = make_smem_desc(st_addr + (ARES ? 0 : NP_ACH), 16, 1024);
            for (int ci = 0; ci < nck; ++ci) {
              const uint64_t ad = ad0 + static_cast<uint64_t>(ci * ((ARES ? NP_ACH : MCH) >> 4));
              const uint64_t bd = bd0 + static_cast<uint64_t>(ci * (MCH >> 4));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                mma_ss_pair(d_S, ad + 2 * k, bd + 2 * k, idesc1, (j | ci | k) != 0 ? 1u : 0u);
            }
            tc_commit_pair(&emptyR[sr], 3);
            if (j == n1 - 1) tc_commit_pair(&s_full[tau % NP_NS], 3);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1) {
        const int u = tau - 1;
        mbar_wait(&g_ready[u % NP_NS], static_cast<uint32_t>(u / NP_NS) & 1u);
        tc_fence_after();
        const uint32_t g_addr = smG_addr + static_cast<uint32_t>(u % NP_NG) * GBUF;
        for (int kh = 0; kh < KH; ++kh) {
          for (int fb = 0; fb < nfb; ++fb) {
            mbar_wait(&fullR[sr], pr);
            tc_fence_after();
            if (elect_one()) {
              // A = G from shared memory (K-major, 64 rows per CTA, 64-wide K chunks of 8 KB)
              const uint64_t gd0 = make_smem_desc(g_addr + static_cast<uint32_t>(2 * kh) * NP_ACH, 16, 1024);
              // B = 128 tile rows as K, 2 x 64 features of this CTA as MN blocks 16 KB apart
              const uint64_t bd = make_smem_desc(smR_addr + static_cast<uint32_t>(sr) * STG, B_CHUNK, 1024);
              const uint32_t d_dA = tmem_base + static_cast<uint32_t>(fb * (SLICE / 2));
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t gd = gd0 + static_cast<uint64_t>((kk >> 2) * (NP_ACH >> 4) + (kk & 3) * 2);
                mma_ss_pair(d_dA, gd, bd + 128 * kk, idesc2, (u > 0 || kh > 0 || kk > 0) ? 1u : 0u);
              }
              tc_commit_pair(&emptyR[sr], 3);
              if (kh == KH - 1 && fb == nfb - 1) {
                tc_commit_pair(&g_empty[u % NP_NG], 3);
                if (u == ntiles - 1) tc_commit_pair(da_full, 3);
              }
            }
            __syncwarp();
            if (++sr == p.nr) { sr = 0; pr ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 3) {
    const float* lse_col = p.lse_col[strip];
    for (int t = 0; t < ntiles; ++t) {
      const int bb = t & 1;
      mbar_wait_parked(&b_empty[bb], (static_cast<uint32_t>(t >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int hh = 0; hh < KH; ++hh) {
        const int cb = (tile_begin + t) * TK + hh * 128 + lane * 4;
        float v[4];
        if (cb + 4 <= p.ncols) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(lse_col + cb));
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldg(lse_col + max(min(cb + k, p.ncols - 1), 0));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = factored ? fast_exp2(lse_mu - v[k]) : v[k] - kGShiftLog2;
        *reinterpret_cast<float4*>(cfbuf + bb * TK + hh * 128 + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[bb]);
    }
  } else if (warp >= 4) {
    // softmax-gradient warps.  Lane group q = warp % 4 sits on TMEM lanes 32q..32q+31:
    //   row of the CTA = (q & 1) * 32 + lane,  tile columns (q >> 1) * 128 + [0,128) in TMEM columns [0,128);
    // the two warps of a lane group split those 128 columns (h = 0 / 1 -> 64 columns each = one
    // 64-wide K chunk of G: the thread writes one whole swizzled 128-byte row of that chunk).
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int rloc = (q & 1) * 32 + lane;
    const int row = row0 + rloc;
    const bool valid = row < p.row_end;
    const int ctile = (q >> 1) * (TK / 2) + h * TPT;  // first tile column of this thread
    const float s = __ldg(p.s_dev);
    const float c = s * kLog2e;
    const float lr2 = valid ? __ldg(p.lse_row[strip] + row) - kGShiftLog2 : INFINITY;
    const float a_i = valid ? fast_exp2(__ldg(p.lse_row[strip] + row) - lse_mu) : 0.f;
    const int label = row + p.label_shift;
    const int warp_label_lo = label - lane;
    const bool g_bf16 = p.g_fmt != 0;
    const uint32_t cf_addr0 = smem_u32(cfbuf) + static_cast<uint32_t>(ctile) * 4;
    // this thread's TPT columns of G: 64-wide K chunk ctile / 64, 16-byte units (ctile % 64) / 8 onwards
    const uint32_t g_row_addr = smem_u32(smG) + static_cast<uint32_t>(ctile / 64) * NP_ACH +
                                static_cast<uint32_t>(rloc) * 128;
    constexpr int G_UNITS = TPT / 8;
    const int g_unit0 = (ctile % 64) / 8;

    for (int t = 0; t < ntiles; ++t) {
      const int sb = t % NP_NS;
      const int bb = t & 1;
      const int cb = (tile_begin + t) * TK + ctile;
      mbar_wait_parked(&b_full[bb], static_cast<uint32_t>(t >> 1) & 1u);
      mbar_wait_parked(&s_full[sb], static_cast<uint32_t>(t / NP_NS) & 1u);
      tc_fence_after();
      uint32_t go[TPT / 2];
#pragma unroll
      for (int hf = 0; hf < TPT / 32; ++hf) {
        const int cbh = cb + 32 * hf;
        const bool has_label = (warp_label_lo < cbh + 32) && (warp_label_lo + 31 >= cbh);
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + tmem_s + sb * SW + h * TPT + hf * 32, r);
        tmem_wait_ld();
        const uint32_t cfa = cf_addr0 + static_cast<uint32_t>(bb * TK + 32 * hf) * 4;
        const int label_rel = label - cbh;
        if (factored) {
          if (g_bf16) np_grad32_dispatch<true, true>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
          else np_grad32_dispatch<true, false>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
        } else {
          if (g_bf16) np_grad32_dispatch<false, true>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
          else np_grad32_dispatch<false, false>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[bb]);
      // G(t) -> shared memory buffer t % 2 once MMA2(t - 2) has drained it
      const int gbi = t % NP_NG;
      mbar_wait_parked(&g_empty[gbi], (static_cast<uint32_t>(t / NP_NG) & 1u) ^ 1u);
      {
        const uint32_t grow = g_row_addr + static_cast<uint32_t>(gbi) * GBUF;
#pragma unroll
        for (int j = 0; j < G_UNITS; ++j)
          sts_v4(grow + static_cast<uint32_t>(((g_unit0 + j) ^ (rloc & 7)) * 16), go[4 * j], go[4 * j + 1],
                 go[4 * j + 2], go[4 * j + 3]);
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
      tc_fence_before();         // the S reads above are ordered before the hand-over as well
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&g_ready[sb], 0);
    }

    // ---- dA: lanes 0-63 hold features [0,128) of each 256-feature block, lanes 64-127 [128,256) ----
    mbar_wait_parked(da_full, 0);
    tc_fence_after();
    const float coef = __ldg(p.grad_out_dev) * s * p.coef_host;
    float* out = p.out[strip] + static_cast<long long>(row - p.row_begin) * p.D;
    for (int fb = 0; fb < nfb; ++fb) {
      for (int ch = h; ch < 4; ch += 2) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + fb * (SLICE / 2) + ch * 32, r);
        tmem_wait_ld();
        const int f0 = f_off + fb * SLICE + (q >> 1) * (SLICE / 2) + ch * 32;
        if (valid && f0 < p.D) {
          if (f0 + 32 <= p.D) {
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              const float a0 = __uint_as_float(r[k]) * coef, a1 = __uint_as_float(r[k + 1]) * coef;
              const float a2 = __uint_as_float(r[k + 2]) * coef, a3 = __uint_as_float(r[k + 3]) * coef;
              if (accumulate) red_add_v4(out + f0 + k, a0, a1, a2, a3);
              else *reinterpret_cast<float4*>(out + f0 + k) = make_float4(a0, a1, a2, a3);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (f0 + k < p.D) {
                if (accumulate) atomicAdd(out + f0 + k, __uint_as_float(r[k]) * coef);
                else out[f0 + k] = __uint_as_float(r[k]) * coef;
              }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ================================================================================================
// Persistent, load-balanced forms of the narrow-pair kernel (see the dispatch for when each is used).  The (strip, row block, tile)
// space is linearised (row-block major, tiles inside) and cut into `npairs` equal contiguous ranges,
// one per resident CTA pair, so that the SMs stay busy whatever 2 * nrb is (with one unit per CTA
// pair, 64 units on 74 pairs leave 14 % of the machine idle at n_loc = 4096).  A pair's range
// crosses row-block boundaries: each piece is a SEGMENT with its own A block, its own dA
// accumulation and a red.add write-out; the MMA1 / softmax / MMA2 pipeline runs straight through
// segment boundaries (MMA1 of the next segment's first tile is issued before MMA2 of the previous
// segment's last tile).  Extra barriers: a_empty (A may be overwritten), da_empty (dA drained).
struct NppCursor {
  int strip, rb, tile, seg;
  int tile_lo, tile_hi;  // the window of column tiles this pair sweeps in every unit it touches
  bool first, last;      // first / last tile of its segment
};
// This pair's share: `nt` tile steps starting at linear index g0 of the space (unit, tile in window).
__device__ __forceinline__ void npp_range(long long pair, const BwdParams& p, long long& g0, long long& nt,
                                          int& tile_lo, int& tile_hi) {
  if (p.npp_t1 > 0) {
    if (pair < p.npp_units) {  // main pair: the first npp_t1 tiles of its own unit
      tile_lo = 0; tile_hi = p.npp_t1;
      g0 = pair * p.npp_t1; nt = p.npp_t1;
    } else {                   // helper pair: an equal share of the last tiles of ALL units
      tile_lo = p.npp_t1; tile_hi = p.ntiles;
      const long long tot = static_cast<long long>(p.npp_units) * (p.ntiles - p.npp_t1);
      const long long h = pair - p.npp_units, nh = p.npairs - p.npp_units;
      g0 = h * tot / nh; nt = (h + 1) * tot / nh - g0;
    }
  } else {
    tile_lo = 0; tile_hi = p.ntiles;
    g0 = pair * p.total_tiles / p.npairs; nt = (pair + 1) * p.total_tiles / p.npairs - g0;
  }
}
__device__ __forceinline__ NppCursor npp_begin(long long g0, long long nt, const BwdParams& p, int tile_lo,
                                               int tile_hi) {
  NppCursor c;
  const int lt = tile_hi - tile_lo;
  const int unit = static_cast<int>(g0 / lt);
  c.tile_lo = tile_lo; c.tile_hi = tile_hi;
  c.tile = tile_lo + static_cast<int>(g0 - static_cast<long long>(unit) * lt);
  c.strip = unit / p.nrb;
  c.rb = unit - c.strip * p.nrb;
  c.seg = 0;
  c.first = true;
  c.last = (nt == 1) || (c.tile == tile_hi - 1);
  return c;
}
// advance to local tile t + 1 (t1 = t + 1 is the new local index)
__device__ __forceinline__ void npp_next(NppCursor& c, long long t1, long long nt, const BwdParams& p) {
  if (++c.tile == c.tile_hi) {
    c.tile = c.tile_lo;
    if (++c.rb == p.nrb) { c.rb = 0; ++c.strip; }
  }
  c.first = c.tile == c.tile_lo;
  if (c.first) ++c.seg;
  c.last = (t1 == nt - 1) || (c.tile == c.tile_hi - 1);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
clip_bwd_npp_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmBm0,
                    const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmBm1,
                    const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;
  uint8_t* smG = smA + static_cast<size_t>(p.kchunks) * NP_ACH;
  uint8_t* smR = smG + static_cast<size_t>(NP_NG) * NP_GBUF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smR + static_cast<size_t>(p.nr) * NP_STAGE);
  {
    uint32_t dyn;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
    if (reinterpret_cast<uint8_t*>(bars) + NP_BAR_BYTES > smem_raw + dyn) __trap();  // carve-up does not fit
  }
  uint64_t* fullR = bars;                 // leader only
  uint64_t* emptyR = fullR + NP_MAXR;
  uint64_t* a_full = emptyR + NP_MAXR;    // leader only
  uint64_t* a_empty = a_full + 1;         // both CTAs (multicast commit)
  uint64_t* s_full = a_empty + 1;         // [NP_NS]
  uint64_t* g_ready = s_full + NP_NS;     // [NP_NS] leader only, 16 arrivals
  uint64_t* g_empty = g_ready + NP_NS;    // [NP_NG]
  uint64_t* da_full = g_empty + NP_NG;    // both CTAs (multicast commit)
  uint64_t* da_empty = da_full + 1;       // leader only, 16 arrivals
  uint64_t* b_full = da_empty + 1;        // [2]
  uint64_t* b_empty = b_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_empty + 2);
  float* cfbuf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + NP_CF_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  long long g0, nt;  // nt >= 1 (host: every pair gets at least one tile)
  int tile_lo, tile_hi;
  npp_range(blockIdx.x >> 1, p, g0, nt, tile_lo, tile_hi);
  const int n1 = (p.kchunks + 1) / 2;  // MMA1 stages per tile (2 chunks each)
  const int nfb = (p.kchunks + 3) / 4; // 256-feature blocks of dA; MMA2 stages per tile = 2 nfb

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA0);
      tma_prefetch_desc(&tmBm0);
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmBm1);
      for (int i = 0; i < NP_MAXR; ++i) { mbar_init(&fullR[i], 1); mbar_init(&emptyR[i], 1); }
      mbar_init(a_full, 1);
      mbar_init(a_empty, 1);
      for (int i = 0; i < NP_NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&g_ready[i], 2 * SM_WARPS); }
      for (int i = 0; i < NP_NG; ++i) mbar_init(&g_empty[i], 1);
      mbar_init(da_full, 1);
      mbar_init(da_empty, 2 * SM_WARPS);
      for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], SM_WARPS); }
      fence_barrier_init();
    }
  } else if (warp == 2) {
    tmem_alloc_pair(tmem_ptr, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  float lse_mu;
  bool factored;
  {
    int lo = __ldg(p.lse_minmax), hi = __ldg(p.lse_minmax + 1);
    lo = lo >= 0 ? lo : lo ^ 0x7fffffff;
    hi = hi >= 0 ? hi : hi ^ 0x7fffffff;
    const float fmin = __int_as_float(lo), fmax = __int_as_float(hi);
    factored = (fmax - fmin) < kFactorRange;
    lse_mu = 0.5f * (fmax + fmin);
  }

  // schedule shared by producer and issuer: step tau: [tau < nt] MMA1(tau); [tau >= 1] MMA2(tau - 1)
  if (warp == 0) {
    int sr = 0;
    uint32_t pr = 0;
    NppCursor cur = npp_begin(g0, nt, p, tile_lo, tile_hi), prv = cur;
    for (long long tau = 0; tau <= nt; ++tau) {
      if (tau < nt) {
        const CUtensorMap* tmBm = cur.strip == 0 ? &tmBm0 : &tmBm1;
        if (cur.first) {
          // the previous segment's MMA1s have finished reading A
          mbar_wait_parked(a_empty, (static_cast<uint32_t>(cur.seg) & 1u) ^ 1u);
          if (elect_one()) {
            const CUtensorMap* tmA = cur.strip == 0 ? &tmA0 : &tmA1;
            const int row0 = p.row_begin + cur.rb * 2 * NP_ROWS + static_cast<int>(rank) * NP_ROWS;
            if (leader) mbar_arrive_expect_tx(a_full, 2u * static_cast<uint32_t>(p.kchunks) * NP_ACH);
            for (int c = 0; c < p.kchunks; ++c)
              tma_load_2d_pair(smA + static_cast<size_t>(c) * NP_ACH, tmA, a_full, c * BK, row0);
          }
          __syncwarp();
        }
        const int col0 = cur.tile * NP_KT + static_cast<int>(rank) * (NP_KT / 2);
        for (int j = 0; j < n1; ++j) {
          const int nck = min(2, p.kchunks - 2 * j);
          mbar_wait_parked(&emptyR[sr], pr ^ 1u);
          if (elect_one()) {
            uint8_t* st = smR + static_cast<size_t>(sr) * NP_STAGE;
            if (leader) mbar_arrive_expect_tx(&fullR[sr], 2u * static_cast<uint32_t>(nck) * B_CHUNK);
            for (int ci = 0; ci < nck; ++ci)
              tma_load_2d_pair(st + ci * B_CHUNK, tmBm, &fullR[sr], (2 * j + ci) * BK, col0);
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1) {
        const CUtensorMap* tmBm = prv.strip == 0 ? &tmBm0 : &tmBm1;
        const int col0 = prv.tile * NP_KT;
        for (int kh = 0; kh < 2; ++kh) {
          for (int fb = 0; fb < nfb; ++fb) {
            mbar_wait_parked(&emptyR[sr], pr ^ 1u);
            if (elect_one()) {
              uint8_t* st = smR + static_cast<size_t>(sr) * NP_STAGE;
              if (leader) mbar_arrive_expect_tx(&fullR[sr], 2u * 2u * B_CHUNK);
              for (int ci = 0; ci < 2; ++ci)
                tma_load_2d_pair(st + ci * B_CHUNK, tmBm, &fullR[sr],
                                 (4 * fb + 2 * static_cast<int>(rank) + ci) * BK, col0 + kh * (NP_KT / 2));
            }
            __syncwarp();
            if (++sr == p.nr) { sr = 0; pr ^= 1u; }
          }
        }
      }
      prv = cur;
      if (tau + 1 < nt) npp_next(cur, tau + 1, nt, p);
    }
  } else if (warp == 1 && leader) {
    const uint32_t fmt = p.idesc1_fmt;
    const uint32_t idesc1 = make_idesc(fmt, fmt, 0, 0, 2 * NP_ROWS, NP_KT);
    const uint32_t idesc2 = make_idesc(p.g_fmt, fmt, 0, 1, 2 * NP_ROWS, SLICE);
    const uint32_t smA_addr = smem_u32(smA), smR_addr = smem_u32(smR), smG_addr = smem_u32(smG);
    int sr = 0;
    uint32_t pr = 0;
    NppCursor cur = npp_begin(g0, nt, p, tile_lo, tile_hi), prv = cur;
    for (long long tau = 0; tau <= nt; ++tau) {
      if (tau < nt) {
        if (cur.first) {
          mbar_wait(a_full, static_cast<uint32_t>(cur.seg) & 1u);
          tc_fence_after();
        }
        const uint32_t d_S = tmem_base + TMEM_S + static_cast<uint32_t>((tau % NP_NS) * (NP_KT / 2));
        for (int j = 0; j < n1; ++j) {
          const int nck = min(2, p.kchunks - 2 * j);
          mbar_wait(&fullR[sr], pr);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad0 = make_smem_desc(smA_addr + static_cast<uint32_t>(2 * j) * NP_ACH, 16, 1024);
            const uint64_t bd0 = make_smem_desc(smR_addr + static_cast<uint32_t>(sr) * NP_STAGE, 16, 1024);
            for (int ci = 0; ci < nck; ++ci) {
              const uint64_t ad = ad0 + static_cast<uint64_t>(ci * (NP_ACH >> 4));
              const uint64_t bd = bd0 + static_cast<uint64_t>(ci * (B_CHUNK >> 4));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                mma_ss_pair(d_S, ad + 2 * k, bd + 2 * k, idesc1, (j | ci | k) != 0 ? 1u : 0u);
            }
            tc_commit_pair(&emptyR[sr], 3);
            if (j == n1 - 1) {
              tc_commit_pair(&s_full[tau % NP_NS], 3);
              if (cur.last) tc_commit_pair(a_empty, 3);  // A may be replaced by the next segment's rows
            }
          }
          __syncwarp();
          if (++sr == p.nr) { sr = 0; pr ^= 1u; }
        }
      }
      if (tau >= 1) {
        const long long u = tau - 1;
        mbar_wait(&g_ready[u % NP_NS], static_cast<uint32_t>(u / NP_NS) & 1u);
        if (prv.first && prv.seg > 0)  // the softmax warps have drained the previous segment's dA
          mbar_wait(da_empty, static_cast<uint32_t>(prv.seg - 1) & 1u);
        tc_fence_after();
        const uint32_t g_addr = smG_addr + static_cast<uint32_t>(u % NP_NG) * NP_GBUF;
        for (int kh = 0; kh < 2; ++kh) {
          for (int fb = 0; fb < nfb; ++fb) {
            mbar_wait(&fullR[sr], pr);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t gd0 = make_smem_desc(g_addr + static_cast<uint32_t>(2 * kh) * NP_ACH, 16, 1024);
              const uint64_t bd = make_smem_desc(smR_addr + static_cast<uint32_t>(sr) * NP_STAGE, B_CHUNK, 1024);
              const uint32_t d_dA = tmem_base + static_cast<uint32_t>(fb * (SLICE / 2));
              const bool fresh = prv.first && kh == 0;  // first MMA into this block of a new segment
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t gd = gd0 + static_cast<uint64_t>((kk >> 2) * (NP_ACH >> 4) + (kk & 3) * 2);
                mma_ss_pair(d_dA, gd, bd + 128 * kk, idesc2, (fresh && kk == 0) ? 0u : 1u);
              }
              tc_commit_pair(&emptyR[sr], 3);
              if (kh == 1 && fb == nfb - 1) {
                tc_commit_pair(&g_empty[u % NP_NG], 3);
                if (prv.last) tc_commit_pair(da_full, 3);
              }
            }
            __syncwarp();
            if (++sr == p.nr) { sr = 0; pr ^= 1u; }
          }
        }
      }
      prv = cur;
      if (tau + 1 < nt) npp_next(cur, tau + 1, nt, p);
    }
  } else if (warp == 3) {
    NppCursor cur = npp_begin(g0, nt, p, tile_lo, tile_hi);
    for (long long t = 0; t < nt; ++t) {
      const float* lse_col = p.lse_col[cur.strip];
      const int bb = static_cast<int>(t & 1);
      mbar_wait_parked(&b_empty[bb], (static_cast<uint32_t>(t >> 1) & 1u) ^ 1u);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int cb = cur.tile * NP_KT + hh * 128 + lane * 4;
        float v[4];
        if (cb + 4 <= p.ncols) {
          const float4 f = __ldg(reinterpret_cast<const float4*>(lse_col + cb));
          v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldg(lse_col + max(min(cb + k, p.ncols - 1), 0));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = factored ? fast_exp2(lse_mu - v[k]) : v[k] - kGShiftLog2;
        *reinterpret_cast<float4*>(cfbuf + bb * NP_KT + hh * 128 + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[bb]);
      if (t + 1 < nt) npp_next(cur, t + 1, nt, p);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int h = (warp - 4) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const int rloc = (q & 1) * 32 + lane;
    const int ctile = (q >> 1) * 128 + h * 64;  // first tile column of this thread
    const float s = __ldg(p.s_dev);
    const float c = s * kLog2e;
    const float coef = __ldg(p.grad_out_dev) * s * p.coef_host;
    const bool g_bf16 = p.g_fmt != 0;
    const uint32_t cf_addr0 = smem_u32(cfbuf) + static_cast<uint32_t>(ctile) * 4;
    const uint32_t g_row_addr = smem_u32(smG) + static_cast<uint32_t>((q >> 1) * 2 + h) * NP_ACH +
                                static_cast<uint32_t>(rloc) * 128;
    // per-segment row state
    int row = 0, label = 0, warp_label_lo = 0;
    bool valid = false;
    float lr2 = INFINITY, a_i = 0.f;

    NppCursor cur = npp_begin(g0, nt, p, tile_lo, tile_hi);
    for (long long t = 0; t < nt; ++t) {
      if (cur.first) {
        row = p.row_begin + cur.rb * 2 * NP_ROWS + static_cast<int>(rank) * NP_ROWS + rloc;
        valid = row < p.row_end;
        const float lr = valid ? __ldg(p.lse_row[cur.strip] + row) : 0.f;
        lr2 = valid ? lr - kGShiftLog2 : INFINITY;
        a_i = valid ? fast_exp2(lr - lse_mu) : 0.f;
        label = row + p.label_shift;
        warp_label_lo = label - lane;
      }
      const int sb = static_cast<int>(t % NP_NS);
      const int bb = static_cast<int>(t & 1);
      const int cb = cur.tile * NP_KT + ctile;
      mbar_wait_parked(&b_full[bb], static_cast<uint32_t>(t >> 1) & 1u);
      mbar_wait_parked(&s_full[sb], static_cast<uint32_t>(t / NP_NS) & 1u);
      tc_fence_after();
      uint32_t go[32];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int cbh = cb + 32 * hf;
        const bool has_label = (warp_label_lo < cbh + 32) && (warp_label_lo + 31 >= cbh);
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + TMEM_S + sb * (NP_KT / 2) + h * 64 + hf * 32, r);
        tmem_wait_ld();
        const uint32_t cfa = cf_addr0 + static_cast<uint32_t>(bb * NP_KT + 32 * hf) * 4;
        const int label_rel = label - cbh;
        if (factored) {
          if (g_bf16) np_grad32_dispatch<true, true>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
          else np_grad32_dispatch<true, false>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
        } else {
          if (g_bf16) np_grad32_dispatch<false, true>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
          else np_grad32_dispatch<false, false>(has_label, r, cfa, c, lr2, a_i, label_rel, go + 16 * hf);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[bb]);
      const int gbi = static_cast<int>(t % NP_NG);
      mbar_wait_parked(&g_empty[gbi], (static_cast<uint32_t>(t / NP_NG) & 1u) ^ 1u);
      {
        const uint32_t grow = g_row_addr + static_cast<uint32_t>(gbi) * NP_GBUF;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts_v4(grow + static_cast<uint32_t>((j ^ (rloc & 7)) * 16), go[4 * j], go[4 * j + 1], go[4 * j + 2], go[4 * j + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&g_ready[sb], 0);

      if (cur.last) {
        // ---- end of a segment: dA (lanes 0-63 features [0,128) of each 256-block, 64-127 [128,256)) ----
        mbar_wait_parked(da_full, static_cast<uint32_t>(cur.seg) & 1u);
        tc_fence_after();
        float* out = p.out[cur.strip] + static_cast<long long>(row - p.row_begin) * p.D;
        for (int fb = 0; fb < nfb; ++fb) {
          for (int ch = h; ch < 4; ch += 2) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + fb * (SLICE / 2) + ch * 32, r);
            tmem_wait_ld();
            const int f0 = fb * SLICE + (q >> 1) * (SLICE / 2) + ch * 32;
            if (valid && f0 < p.D) {
              if (f0 + 32 <= p.D) {
#pragma unroll
                for (int k = 0; k < 32; k += 4)
                  red_add_v4(out + f0 + k, __uint_as_float(r[k]) * coef, __uint_as_float(r[k + 1]) * coef,
                             __uint_as_float(r[k + 2]) * coef, __uint_as_float(r[k + 3]) * coef);
              } else {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                  if (f0 + k < p.D) atomicAdd(out + f0 + k, __uint_as_float(r[k]) * coef);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(da_empty, 0);
      }
      if (t + 1 < nt) npp_next(cur, t + 1, nt, p);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// min / max over both lse arrays, written as order-preserving ints.  One block (2N floats is at
// most a few hundred KB): a single launch, no memset, no atomics.
__global__ void __launch_bounds__(1024) lse_minmax_kernel(const float* __restrict__ a,
                                                          const float* __restrict__ b, int n,
                                                          int* __restrict__ out) {
  __shared__ float slo[32], shi[32];
  float lo = INFINITY, hi = -INFINITY;
  // both arrays are 16-byte aligned (checked by the caller): independent float4 loads in flight
  const int n4 = n >> 2;
#pragma unroll
  for (int arr = 0; arr < 2; ++arr) {
    const float* __restrict__ src = arr == 0 ? a : b;
    const float4* __restrict__ v4 = reinterpret_cast<const float4*>(src);
#pragma unroll 8
    for (int i = threadIdx.x; i < n4; i += 1024) {
      const float4 v = __ldg(v4 + i);
      lo = fminf(fminf(lo, fminf(v.x, v.y)), fminf(v.z, v.w));
      hi = fmaxf(fmaxf(hi, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += 1024) {
      lo = fminf(lo, src[i]);
      hi = fmaxf(hi, src[i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    slo[threadIdx.x >> 5] = lo;
    shi[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    lo = slo[threadIdx.x];
    hi = shi[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (threadIdx.x == 0) {
      int il = __float_as_int(lo), ih = __float_as_int(hi);
      out[0] = il >= 0 ? il : il ^ 0x7fffffff;
      out[1] = ih >= 0 ? ih : ih ^ 0x7fffffff;
    }
  }
}

// fp32 -> 16-bit cast of the gradient when the caller wants fp16 / bf16 outputs
__global__ void cast_out_kernel(const float* __restrict__ in, void* __restrict__ out, int out_dtype,
                                long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 v = *reinterpret_cast<const float2*>(in + i);
    if (out_dtype == NANS_BF16)
      *reinterpret_cast<__nv_bfloat162*>(static_cast<__nv_bfloat16*>(out) + i) = __floats2bfloat162_rn(v.x, v.y);
    else
      *reinterpret_cast<__half2*>(static_cast<__half*>(out) + i) = __floats2half2_rn(v.x, v.y);
  } else if (i < n) {
    if (out_dtype == NANS_BF16) static_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16(in[i]);
    else static_cast<__half*>(out)[i] = __float2half(in[i]);
  }
}

// Narrow pairs: whole waves of units run unsplit; the partial last wave (r units) is cut ns ways by
// columns so that r * ns sub-units fill the pairs again.  Per-unit overhead (cluster start, A load,
// pipeline fill, dA write-out) is worth about 3 tiles of 256 columns (6 of 128).
struct NpTail {
  int n_full, ns_tail;
};
NpTail plan_np_tail(int64_t units, int64_t ntiles, int slots, double overhead_tiles) {
  NpTail t;
  t.n_full = static_cast<int>(units / slots * slots);
  t.ns_tail = 1;
  const int64_t r = units - t.n_full;
  if (r == 0) return t;
  const int64_t max_ns = std::max<int64_t>(1, std::min<int64_t>(32, ntiles / 4));
  double best = 1e300;
  for (int64_t ns = 1; ns <= max_ns; ++ns) {
    const double waves = static_cast<double>(ceil_div(r * ns, slots));
    const double cost = waves * (static_cast<double>(ceil_div(ntiles, ns)) + overhead_tiles);
    if (cost < best - 1e-9) {
      best = cost;
      t.ns_tail = static_cast<int>(ns);
    }
  }
  return t;
}

int choose_bwd_nsplit(int64_t rows, int64_t N, int npass, int rows_per_unit, int slots) {
  const int64_t base = 2 * ceil_div(rows, rows_per_unit) * npass;
  const int64_t ntiles = ceil_div(N, KT);
  const int sms = slots;
  int best = 1;
  double best_cost = 1e300;
  const int64_t max_ns = ntiles / 8 < 1 ? 1 : (ntiles / 8 > 32 ? 32 : ntiles / 8);
  for (int64_t ns = 1; ns <= max_ns; ++ns) {
    const double waves = static_cast<double>(ceil_div(base * ns, sms));
    // per-unit overhead: prologue + A block load + dA write-out, worth about 6 tiles
    const double cost = waves * (static_cast<double>(ceil_div(ntiles, ns)) + 6.0);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = static_cast<int>(ns);
    }
  }
  return best;
}

}  // namespace
}  // namespace nans

using namespace nans;

extern "C" size_t nans_clip_loss_bwd_workspace_bytes(int64_t grad_row_count, int64_t N, int64_t D) {
  (void)N;
  if (grad_row_count <= 0 || D <= 0) return 512;
  return 2 * align_up(static_cast<size_t>(grad_row_count) * D * 4, 256) + 512;
}

extern "C" int nans_clip_loss_bwd(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                  const void* T_all, const void* I_all, int64_t ld_all,
                                  int feat_dtype, int64_t n_loc, int64_t N, int64_t D,
                                  int64_t label_begin, const float* s_dev, const float* lse_img_all,
                                  const float* lse_txt_all, const float* grad_out_dev,
                                  float grad_mult, int64_t grad_row_begin, int64_t grad_row_count,
                                  void* dI_loc, void* dT_loc, int out_dtype, void* ws,
                                  size_t ws_bytes, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(feat_dtype == NANS_F16 || feat_dtype == NANS_BF16,
               "loss_bwd: feat_dtype must be NANS_F16 or NANS_BF16");
  NANS_REQUIRE(out_dtype == NANS_F32 || out_dtype == NANS_F16 || out_dtype == NANS_BF16,
               "loss_bwd: bad out_dtype");
  NANS_REQUIRE(n_loc > 0 && N > 0 && D > 0 && D % 8 == 0, "loss_bwd: bad sizes (D must be a multiple of 8)");
  NANS_REQUIRE(n_loc < (1ll << 30) && N < (1ll << 30) && D <= 8192, "loss_bwd: size too large");
  NANS_REQUIRE(grad_row_begin >= 0 && grad_row_count >= 0 && grad_row_begin + grad_row_count <= n_loc,
               "loss_bwd: gradient rows [%lld, +%lld) outside the local block of %lld rows",
               (long long)grad_row_begin, (long long)grad_row_count, (long long)n_loc);
  NANS_REQUIRE(label_begin >= 0 && label_begin + n_loc <= N, "loss_bwd: labels outside [0, N)");
  if (grad_row_count == 0) return NANS_OK;
  NANS_REQUIRE(I_loc && T_loc && T_all && I_all && s_dev && lse_img_all && lse_txt_all &&
                   grad_out_dev && dI_loc && dT_loc,
               "loss_bwd: null pointer");
  NANS_REQUIRE(ld_loc >= D && ld_all >= D, "loss_bwd: leading dimension smaller than D");
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(lse_img_all) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(lse_txt_all) & 15) == 0,
               "loss_bwd: lse arrays must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int kchunks = static_cast<int>(ceil_div(D, BK));
  const int npass = static_cast<int>(ceil_div(kchunks, SLICE / BK));
  // CTA pairs (cta_group::2) by default; NANS_BWD_1CTA=1 selects the single-CTA kernel
  bool use_pair = true;
  {
    const char* e = getenv("NANS_BWD_1CTA");
    if (e && e[0] == '1') use_pair = false;
  }
  // narrow pairs (64 rows per CTA, no S recompute per feature slice) whenever D <= 512;
  // NANS_BWD_NP=0 falls back to the 128-row pair kernel
  // D <= 768: dA of 64 rows fits TMEM beside two S buffers; D <= 1024: two passes of 512 features, A streamed
  bool use_np = use_pair && kchunks <= 16;
  const bool np_ares = kchunks <= 12;
  const int np_tk = (kchunks <= 8 || !np_ares) ? NP_KT : 128;  // 128-column tiles only for 512 < D <= 768
  const int np_npass = np_ares ? 1 : 2;
  {
    const char* e = getenv("NANS_BWD_NP");
    if (e && e[0] == '0') use_np = false;
  }
  // persistent load-balanced form of the narrow-pair kernel: NANS_BWD_PERSIST=1
  bool use_npp = false;
  int npp_t1 = 0, npp_units = 0;
  if (use_np) {
    // worth it when whole waves of one unit per CTA pair would leave SMs idle (64 units on 74 pairs at
    // n_loc = 4096); with many waves (N = 32768 on one GPU: 512 units, 98.8 % full) the plain-store
    // kernel wins: no output memset, no red.add, no segment hand-overs
    const int64_t units = 2 * ceil_div(grad_row_count, 2 * NP_ROWS);
    const int64_t slots = sm_count() / 2;
    const double fill = static_cast<double>(units) / static_cast<double>(ceil_div(units, slots) * slots);
    // Equal contiguous ranges (NANS_BWD_PERSIST=1): measured at 2 GPUs (n_loc = 16384) 2.61 ms/step
    // against 2.50 — the pairs no longer walk the column tiles in lockstep, so the column operands
    // stop hitting in L2.  Opt-in only.
    // Helper mode (default when there are FEWER units than CTA pairs, e.g. 64 units on 74 pairs at
    // n_loc = 4096): every unit keeps its own pair for the first T1 tiles (lockstep preserved) and the
    // idle pairs share the last ntiles - T1 tiles of all units.  NANS_BWD_PERSIST=0 disables it.
    const char* e = getenv("NANS_BWD_PERSIST");
    const int64_t nt256 = ceil_div(N, NP_KT);
    const bool force_helpers = e && e[0] == '2';  // tests: helper mode wherever it is feasible
    if (e && e[0] == '1') {
      use_npp = kchunks <= 8;
    } else if (!(e && e[0] == '0') && kchunks <= 8 && units < slots && nt256 >= 2 &&
               (force_helpers || (fill < 0.95 && slots - units >= 2 && nt256 >= 32))) {
      // main pair: T1 tiles + start/drain (~3 tiles); helper: units * (nt - T1) / helpers tiles + ~1.2
      // tiles per unit boundary + the same start/drain  =>  T1 = units * (nt + 1.2) / slots
      int64_t t1 = std::max<int64_t>(1, std::min<int64_t>(
          nt256 - 1, (units * (10 * nt256 + 12) + 10 * slots - 1) / (10 * slots)));
      if (const char* o = getenv("NANS_NPP_T1")) t1 = std::max<int64_t>(1, std::min<int64_t>(nt256 - 1, atoi(o)));  // tuning
      if (units * (nt256 - t1) >= slots - units) {  // every helper pair gets at least one tile
        use_npp = true;
        npp_t1 = static_cast<int>(t1);
        npp_units = static_cast<int>(units);
      }
    }
  }
  const BwdPlan plan = plan_bwd(kchunks);
  const PairPlan pplan = plan_pair(kchunks);
  const NpPlan nplan = plan_np(kchunks, np_tk, np_ares);
  const int64_t np_units = 2 * ceil_div(grad_row_count, 2 * NP_ROWS) * np_npass;
  const NpTail tail = plan_np_tail(np_units, ceil_div(N, np_tk), sm_count() / 2, np_tk == NP_KT ? 3.0 : 6.0);
  const int nsplit = use_npp    ? 2  /* outputs are always accumulated (zeroed below) */
                     : use_np   ? 1  /* per-unit: see plan_np_tail */
                     : use_pair ? choose_bwd_nsplit(grad_row_count, N, npass, 2 * BM, sm_count() / 2)
                                : choose_bwd_nsplit(grad_row_count, N, npass, BM, sm_count());
  const size_t out_bytes = static_cast<size_t>(grad_row_count) * D * 4;
  // workspace: [0,256) lse min/max slots, then (16-bit outputs only) the two fp32 gradient buffers
  if (ws == nullptr || ws_bytes < 512) {
    set_error("loss_bwd: workspace %zu < 512 bytes", ws_bytes);
    return NANS_ERR_WORKSPACE;
  }
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "loss_bwd: workspace must be 16-byte aligned");
  int* minmax = static_cast<int*>(ws);
  lse_minmax_kernel<<<1, 1024, 0, st>>>(lse_img_all, lse_txt_all, static_cast<int>(N), minmax);
  NANS_CUDA_OK(cudaGetLastError());
  float* out32[2];
  if (out_dtype == NANS_F32) {
    NANS_REQUIRE((reinterpret_cast<uintptr_t>(dI_loc) & 15) == 0 && (reinterpret_cast<uintptr_t>(dT_loc) & 15) == 0,
                 "loss_bwd: outputs must be 16-byte aligned");
    out32[0] = static_cast<float*>(dI_loc);
    out32[1] = static_cast<float*>(dT_loc);
  } else {
    const size_t need = nans_clip_loss_bwd_workspace_bytes(grad_row_count, N, D);
    if (ws_bytes < need) {
      set_error("loss_bwd: workspace %zu < %zu bytes", ws_bytes, need);
      return NANS_ERR_WORKSPACE;
    }
    out32[0] = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256);
    out32[1] = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 256 + align_up(out_bytes, 256));
  }
  if (nsplit > 1) {
    NANS_CUDA_OK(cudaMemsetAsync(out32[0], 0, out_bytes, st));
    NANS_CUDA_OK(cudaMemsetAsync(out32[1], 0, out_bytes, st));
  } else if (use_np && !use_npp && tail.ns_tail > 1 && np_npass > 1) {
    NANS_CUDA_OK(cudaMemsetAsync(out32[0], 0, out_bytes, st));
    NANS_CUDA_OK(cudaMemsetAsync(out32[1], 0, out_bytes, st));
  } else if (use_np && !use_npp && tail.ns_tail > 1) {
    // only the rows of the column-split tail units are accumulated
    const int64_t nrb_np = np_units / 2;
    for (int strip = 0; strip < 2; ++strip) {
      const int64_t u0 = std::max<int64_t>(tail.n_full, strip * nrb_np), u1 = (strip + 1) * nrb_np;
      if (u0 >= u1) continue;
      const int64_t r0 = (u0 - strip * nrb_np) * 2 * NP_ROWS;
      const int64_t r1 = std::min<int64_t>(grad_row_count, (u1 - strip * nrb_np) * 2 * NP_ROWS);
      NANS_CUDA_OK(cudaMemsetAsync(out32[strip] + r0 * D, 0, static_cast<size_t>(r1 - r0) * D * 4, st));
    }
  }

  CUtensorMap tmA0, tmB0, tmA1, tmB1, tmBk0, tmBk1;
  if ((rc = make_tmap_16b(&tmA0, I_loc, feat_dtype, n_loc, D, ld_loc, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB0, T_all, feat_dtype, N, D, ld_all, KT)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmA1, T_loc, feat_dtype, n_loc, D, ld_loc, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB1, I_all, feat_dtype, N, D, ld_all, KT)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmBk0, T_all, feat_dtype, N, D, ld_all, KT / 2)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmBk1, I_all, feat_dtype, N, D, ld_all, KT / 2)) != NANS_OK) return rc;
  CUtensorMap tmAn0, tmAn1;  // narrow pairs: 64-row A boxes
  if ((rc = make_tmap_16b(&tmAn0, I_loc, feat_dtype, n_loc, D, ld_loc, NP_ROWS)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmAn1, T_loc, feat_dtype, n_loc, D, ld_loc, NP_ROWS)) != NANS_OK) return rc;

  BwdParams p;
  p.row_begin = static_cast<int>(grad_row_begin);
  p.row_end = static_cast<int>(grad_row_begin + grad_row_count);
  p.ncols = static_cast<int>(N);
  p.D = static_cast<int>(D);
  p.kchunks = kchunks;
  p.nrb = static_cast<int>(ceil_div(grad_row_count, use_np ? 2 * NP_ROWS : (use_pair ? 2 * BM : BM)));
  p.npass = npass;
  p.nsplit = nsplit;
  p.ntiles = static_cast<int>(ceil_div(N, use_np ? np_tk : KT));
  p.nr = use_np ? nplan.nr : (use_pair ? pplan.nr : plan.nr);
  p.idesc1_fmt = static_cast<uint32_t>(idesc_fmt(feat_dtype));
  // tcgen05.mma kind::f16 wants A and B in the same 16-bit format (a mixed f16 x bf16 descriptor
  // faults as an illegal instruction on sm_100a), so G is written in the features' format.
  // NANS_BWD_MIXED_G=1 forces f16 G regardless (bring-up experiment).
  {
    const char* e = getenv("NANS_BWD_MIXED_G");
    p.g_fmt = (e && e[0] == '1') ? 0u : p.idesc1_fmt;
  }
  p.label_shift = static_cast<int>(label_begin);
  p.s_dev = s_dev;
  p.grad_out_dev = grad_out_dev;
  p.coef_host = grad_mult / (2.0f * static_cast<float>(N)) / 4096.0f;
  p.lse_row[0] = lse_img_all + label_begin;
  p.lse_row[1] = lse_txt_all + label_begin;
  p.lse_col[0] = lse_txt_all;
  p.lse_col[1] = lse_img_all;
  p.out[0] = out32[0];
  p.out[1] = out32[1];
  p.accumulate = nsplit > 1 ? 1 : 0;
  p.lse_minmax = minmax;
  {
    const char* e = getenv("NANS_BWD_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }

  if (use_npp) {
    p.total_tiles = 2ll * p.nrb * p.ntiles;
    p.npp_t1 = npp_t1;
    p.npp_units = npp_units;
    if (npp_t1 > 0) {
      p.npairs = sm_count() / 2;  // npp_units main pairs + the helpers
    } else {
      // at least ~4 tiles per pair, at most one pair per two SMs
      const long long want = std::max<long long>(1, p.total_tiles / 4);
      p.npairs = static_cast<int>(std::min<long long>(sm_count() / 2, want));
    }
    NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_npp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(nplan.bytes)));
    clip_bwd_npp_kernel<<<static_cast<unsigned>(2 * p.npairs), NUM_THREADS, nplan.bytes, st>>>(tmAn0, tmB0, tmAn1,
                                                                                            tmB1, p);
  } else if (use_np) {
    p.n_full = tail.n_full;
    p.ns_tail = tail.ns_tail;
    const unsigned grid = static_cast<unsigned>(2 * (tail.n_full + (np_units - tail.n_full) * tail.ns_tail));
    p.npass = np_npass;
    if (np_tk == NP_KT && !np_ares) {
      NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_np_kernel<NP_KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(nplan.bytes)));
      clip_bwd_np_kernel<NP_KT, false><<<grid, NUM_THREADS, nplan.bytes, st>>>(tmAn0, tmB0, tmB0, tmAn1, tmB1, tmB1, p);
    } else if (np_tk == NP_KT) {
      NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_np_kernel<NP_KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(nplan.bytes)));
      clip_bwd_np_kernel<NP_KT, true><<<grid, NUM_THREADS, nplan.bytes, st>>>(tmAn0, tmB0, tmB0, tmAn1, tmB1, tmB1, p);
    } else {
      NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_np_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(nplan.bytes)));
      clip_bwd_np_kernel<128, true><<<grid, NUM_THREADS, nplan.bytes, st>>>(tmAn0, tmBk0, tmB0, tmAn1, tmBk1, tmB1, p);
    }
  } else if (use_pair) {
    auto kern = pplan.a_resident ? clip_bwd_pair_kernel<true> : clip_bwd_pair_kernel<false>;
    NANS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(pplan.bytes)));
    const unsigned grid = static_cast<unsigned>(2 * 2 * p.nrb * p.npass * p.nsplit);  // 2 CTAs per unit
    kern<<<grid, NUM_THREADS, pplan.bytes, st>>>(tmA0, tmBk0, tmB0, tmA1, tmBk1, tmB1, p);
  } else {
    auto kern = plan.a_resident ? clip_bwd_kernel<true> : clip_bwd_kernel<false>;
    NANS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(plan.bytes)));
    const unsigned grid = static_cast<unsigned>(2 * p.nrb * p.npass * p.nsplit);
    kern<<<grid, NUM_THREADS, plan.bytes, st>>>(tmA0, tmB0, tmA1, tmB1, p);
  }
  NANS_CUDA_OK(cudaGetLastError());

  if (out_dtype != NANS_F32) {
    const long long n = static_cast<long long>(grad_row_count) * D;
    const unsigned g = static_cast<unsigned>(ceil_div(ceil_div(n, 2), 256));
    cast_out_kernel<<<g, 256, 0, st>>>(out32[0], dI_loc, out_dtype, n);
    cast_out_kernel<<<g, 256, 0, st>>>(out32[1], dT_loc, out_dtype, n);
    NANS_CUDA_OK(cudaGetLastError());
  }
  return NANS_OK;
}
