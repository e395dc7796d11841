// strip_fwd.cu — kernel (2): fused contrastive forward.
//
// Replaces cn_clip/training/train.py:87-88 / 103-104 (the (s*I)@T^T logits), :109-115 (two
// CrossEntropyLoss against arange labels) and :117-121 (argmax accuracy).  The logits matrix is
// never written: every 128x256 tile lives in TMEM and is consumed by an online log-sum-exp.
//
// Per rank two strips are swept (strip 0: I_loc x T_cols^T, strip 1: T_loc x I_cols^T); each row
// keeps, in the log2 domain with t = cos * s * log2(e):
//     m = running max t,  l = sum 2^(t - m),  w = sum 2^(t - m) * cos        (for d loss / d s)
// plus cos at the label column and (optionally) the running arg-max.  A unit (= one CTA) is
// (strip, 128-row block, column split); units write partial statistics into workspace slots, and
// clip_fwd_finalize_kernel merges the slots into lse, loss / d(scale) sums and hit counts.
//
// Roofline: tensor cores.  Algorithmic flops per unit row block = 2 * 128 * ncols * D per strip.
#include <stdlib.h>

#include "strip_sweep.cuh"

namespace nans {
namespace {

using namespace sweep;

constexpr float kMasked = -1e30f;

struct FwdParams {
  int n_loc, ncols, kchunks, stages;
  uint32_t idesc;
  int nrb, nsplit, ntiles;
  int strip0;       // first strip of this launch (1 for a text-rows-only launch)
  int label_shift;  // label column of row r in phase coordinates = r + label_shift ...
  int lab_row0, lab_row1;  // ... for rows in [lab_row0, lab_row1); the other rows have no label in this phase
  int col_global_begin;  // global column of phase column 0
  int skip_begin, skip_count;  // tiles of the column operand this phase does not sweep
  const float* s_dev;
  float* part_m;
  float* part_l;
  float* part_w;
  float* part_bv;
  int* part_bi;
  float* diag;
  int slot_begin;
  long long slot_stride;  // floats between consecutive slots
  // exchange mode (nans_clip_loss_fwd_xchg; see SweepArgs): xw = 0 otherwise
  int xw, xrank, xsrc_tiles;
  int xslot_rows;                // rows of one slot of the gathered tensor (= N)
  const uint32_t* xflags[2];     // per strip: arrival flags of its column modality
  const uint32_t* xepoch;        // device word: completed steps; this launch belongs to step *xepoch + 1
  long long* xwait_ns;           // diagnostics (NANS_XCHG_PROBE): per-CTA flag-wait nanoseconds, or null
};

template <bool WITH_ACC>
struct LseEpi {
  float c;  // s * log2(e)
  int ncols;
  int label;          // this thread's label column (phase coordinates)
  int warp_label_lo;  // label column of lane 0 of this warp
  float m, l[4], w[4], diag, bv;
  int bi;

  __device__ __forceinline__ void init(float c_, int ncols_, int label_, int lane) {
    c = c_;
    ncols = ncols_;
    label = label_;
    warp_label_lo = label_ - lane;
    m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = w[i] = 0.f;
    diag = 0.f;
    bv = -INFINITY;
    bi = -1;
  }

  __device__ __forceinline__ void tile(uint32_t taddr, int tile_idx) {
    const int col0 = tile_idx * BN;
    const bool tail = col0 + BN > ncols;
    const bool has_label = (warp_label_lo < col0 + BN) && (warp_label_lo + 31 >= col0);
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      const int cb = col0 + ch * 32;
      if (cb >= ncols) break;  // whole chunk past the last column: the running max must only
                               // ever come from real columns (an all-masked max would make
                               // exp2(v*c - m) a rounding-error lottery)
      uint32_t r[32];
      tmem_ld32(taddr + ch * 32, r);
      tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
      if (tail) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (cb + k >= ncols) v[k] = kMasked;
      }
      if (has_label) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (cb + k == label) diag = v[k];
      }
      if (WITH_ACC) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (v[k] > bv) {  // strict: the first maximum wins, as torch.argmax does
            bv = v[k];
            bi = cb + k;
          }
      }
      float mx[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx[i] = v[i];
#pragma unroll
      for (int k = 4; k < 32; ++k) mx[k & 3] = fmaxf(mx[k & 3], v[k]);
      const float cmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const float m_new = fmaxf(m, cmax * c);
      const float alpha = fast_exp2(m - m_new);
      m = m_new;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        l[i] *= alpha;
        w[i] *= alpha;
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float e = fast_exp2(fmaf(v[k], c, -m));
        l[k & 3] += e;
        w[k & 3] = fmaf(e, v[k], w[k & 3]);
      }
    }
  }
};

// CP: CTA-pair mode (cluster of 2, cta_group::2 MMAs); a unit is then a 256-row block.
template <bool A_RES, bool WITH_ACC, bool CP>
__global__ void __cluster_dims__(CP ? 2 : 1, 1, 1) __launch_bounds__(NUM_THREADS, 1)
clip_fwd_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                const FwdParams p) {
  const int unit = CP ? (blockIdx.x >> 1) : blockIdx.x;
  const int split = unit % p.nsplit;
  const int rb = (unit / p.nsplit) % p.nrb;
  const int strip = unit / (p.nsplit * p.nrb) + p.strip0;

  SweepArgs a;
  a.tmA = strip == 0 ? &tmA0 : &tmA1;
  a.tmB = strip == 0 ? &tmB0 : &tmB1;
  a.row0 = CP ? rb * 2 * BM + static_cast<int>(blockIdx.x & 1) * BM : rb * BM;
  if (p.xw > 0) {
    // column split = a sub-range of every source rank's tiles
    a.xw = p.xw;
    a.xrank = p.xrank;
    a.xsrc_tiles = p.xsrc_tiles;
    a.xsub_begin = static_cast<int>(static_cast<long long>(split) * p.xsrc_tiles / p.nsplit);
    a.xsub_count = static_cast<int>(static_cast<long long>(split + 1) * p.xsrc_tiles / p.nsplit) - a.xsub_begin;
    a.xstep = *p.xepoch + 1u;
    a.xrow_off = static_cast<int>(a.xstep & 1u) * p.xslot_rows;
    a.xflags = p.xflags[strip];
    a.xwait_ns = p.xwait_ns;
    a.tile_begin = 0;
    a.tile_end = p.xw * a.xsub_count;
  } else {
    a.tile_begin = static_cast<int>(static_cast<long long>(split) * p.ntiles / p.nsplit);
    a.tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / p.nsplit);
  }
  a.skip_begin = p.skip_begin;
  a.skip_count = p.skip_count;
  a.kchunks = p.kchunks;
  a.stages = p.stages;
  a.idesc = p.idesc;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = a.row0 + (warp & 3) * 32 + lane;

  LseEpi<WITH_ACC> epi;
  epi.init(__ldg(p.s_dev) * kLog2e, p.ncols,
           (row >= p.lab_row0 && row < p.lab_row1) ? row + p.label_shift : -(1 << 29), lane);

  uint8_t* scratch = run<A_RES, CP>(a, epi);

  // The even and odd tiles of a row were drained by two threads (warps 4-7 and 8-11): merge
  // through shared memory (the pipeline buffers are dead after run()).
  float* xm = reinterpret_cast<float*>(scratch);  // [6][128]
  const int r = (warp & 3) * 32 + lane;
  float L = (epi.l[0] + epi.l[1]) + (epi.l[2] + epi.l[3]);
  float Wt = (epi.w[0] + epi.w[1]) + (epi.w[2] + epi.w[3]);
  // does this unit sweep the tile that holds the row's label?  (swept index = tile index minus the
  // skipped tiles before it; tiles inside the skipped range belong to another phase)
  const int lab_tile = epi.label >= 0 && epi.label < p.ncols ? epi.label / BN : -1;
  const int lab_at = swept_index_of(a, lab_tile);  // swept index at which this unit visits the label's tile
  const bool has_diag = lab_at >= 0;
  if (warp >= 8) {
    xm[0 * 128 + r] = epi.m;
    xm[1 * 128 + r] = L;
    xm[2 * 128 + r] = Wt;
    xm[3 * 128 + r] = epi.diag;
    xm[4 * 128 + r] = epi.bv;
    reinterpret_cast<int*>(xm)[5 * 128 + r] = epi.bi;
  }
  __syncthreads();
  if (warp >= 4 && warp < 8 && row < p.n_loc) {
    const float m2 = xm[0 * 128 + r];
    const float M = fmaxf(epi.m, m2);
    const float s1 = fast_exp2(epi.m - M), s2 = fast_exp2(m2 - M);
    L = L * s1 + xm[1 * 128 + r] * s2;
    Wt = Wt * s1 + xm[2 * 128 + r] * s2;
    // the label column lies in exactly one tile: the set that owns that tile holds cos there
    const int lab_set = has_diag ? (lab_at & 1) : 0;
    const float diag = lab_set == 0 ? epi.diag : xm[3 * 128 + r];
    float bv = epi.bv;
    int bi = epi.bi;
    if (WITH_ACC) {
      const float bv2 = xm[4 * 128 + r];
      const int bi2 = reinterpret_cast<const int*>(xm)[5 * 128 + r];
      if (bv2 > bv || (bv2 == bv && bi2 >= 0 && (bi < 0 || bi2 < bi))) {
        bv = bv2;
        bi = bi2;
      }
    }
    const long long idx = static_cast<long long>(p.slot_begin + split) * p.slot_stride +
                          static_cast<long long>(strip) * p.n_loc + row;
    p.part_m[idx] = M;
    p.part_l[idx] = L;
    p.part_w[idx] = Wt;
    if (WITH_ACC) {
      p.part_bv[idx] = bv;
      p.part_bi[idx] = bi < 0 ? -1 : bi + p.col_global_begin;
    }
    if (has_diag) p.diag[static_cast<long long>(strip) * p.n_loc + row] = diag;
  }
}

// ---- finalize ------------------------------------------------------------------------------
struct FinParams {
  int n_loc, total_slots, with_acc;
  long long slot_stride;
  int label_begin;  // global label column of local row 0 (arg-max indices are global columns)
  const float* s_dev;
  const float* part_m;
  const float* part_l;
  const float* part_w;
  const float* part_bv;
  const int* part_bi;
  const float* diag;
  float* lse_img;
  float* lse_txt;
  double* block_part;  // [gridDim.x][6]
  unsigned* counter;
  float* scalars;  // [8]
  float* row_stats;  // optional [3][2 * n_loc]: per-row loss term, d/ds term, arg-max column (int bits)
  // push mode (nans_clip_loss_fwd_finalize_push): the packed result [lse_img (pad) | lse_txt (pad) | 8
  // scalars] also goes to row xrank of slot (step & 1) of every rank's lse table, then their flag xrank
  int xw, xrank, xpad;
  long long xlse_off, xlse_len, xlflag_off;
  const uint32_t* xepoch;
  uint8_t* xbase[NANS_MAX_PEERS];
};

__global__ void __launch_bounds__(256) clip_fwd_finalize_kernel(const FinParams p) {
  __shared__ double red[6][8];
  __shared__ bool is_last;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 2 * p.n_loc;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (gid < total) {
    const int strip = gid / p.n_loc;
    const int row = gid - strip * p.n_loc;
    float M = -INFINITY;
    for (int s = 0; s < p.total_slots; ++s) M = fmaxf(M, p.part_m[s * p.slot_stride + gid]);
    float L = 0.f, Wt = 0.f, bv = -INFINITY;
    int bi = -1;
    for (int s = 0; s < p.total_slots; ++s) {
      const long long i = s * p.slot_stride + gid;
      const float sc = exp2f(p.part_m[i] - M);
      L = fmaf(p.part_l[i], sc, L);
      Wt = fmaf(p.part_w[i], sc, Wt);
      if (p.with_acc) {
        const float v = p.part_bv[i];
        const int b = p.part_bi[i];
        if (v > bv || (v == bv && b >= 0 && (bi < 0 || b < bi))) {
          bv = v;
          bi = b;
        }
      }
    }
    const float sc = __ldg(p.s_dev);
    const float lse2 = M + log2f(L);  // base-2 domain: log2 sum_j 2^(cos_ij * s * log2 e)
    const float lse = lse2 * kLn2;
    const float cosd = p.diag[gid];
    (strip == 0 ? p.lse_img : p.lse_txt)[row] = lse2;
    if (p.xw > 0) {
      const uint32_t step = *p.xepoch + 1u;
      const long long at = p.xlse_off + ((static_cast<long long>(step & 1u) * p.xw + p.xrank) * p.xlse_len +
                                         static_cast<long long>(strip) * p.xpad + row) * 4;
      for (int r = 0; r < p.xw; ++r) *reinterpret_cast<float*>(p.xbase[r] + at) = lse2;
    }
    acc[strip] = static_cast<double>(lse) - static_cast<double>(sc) * cosd;
    acc[2 + strip] = static_cast<double>(Wt) / static_cast<double>(L) - cosd;
    if (p.with_acc) acc[4 + strip] = (bi == p.label_begin + row) ? 1.0 : 0.0;
    if (p.row_stats) {
      p.row_stats[gid] = static_cast<float>(acc[strip]);
      p.row_stats[total + gid] = static_cast<float>(acc[2 + strip]);
      p.row_stats[2 * total + gid] = __int_as_float(bi);
    }
  }
  // block reduction (fixed order -> deterministic)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[q][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0;
    for (int i = 0; i < 8; ++i) v += red[threadIdx.x][i];
    p.block_part[blockIdx.x * 6 + threadIdx.x] = v;
  }
  // every thread's stores (block partials; in push mode also the lse values written into the peers'
  // tables) are made visible — at system scope when they crossed NVLink — before this block is counted
  __syncthreads();
  if (threadIdx.x == 0) {
    // cumulative over the block's stores (ordered before this thread by the barrier): one fence per block
    if (p.xw > 0) __threadfence_system();
    else __threadfence();
    const unsigned done = atomicAdd(p.counter, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (threadIdx.x < 8) {
      double v = 0;
      if (threadIdx.x < 6)
        for (unsigned b = 0; b < gridDim.x; ++b)
          v += *(volatile double*)&p.block_part[b * 6 + threadIdx.x];
      p.scalars[threadIdx.x] = static_cast<float>(v);
      if (p.xw > 0) {
        const uint32_t step = *p.xepoch + 1u;
        const long long at = p.xlse_off + ((static_cast<long long>(step & 1u) * p.xw + p.xrank) * p.xlse_len +
                                           2ll * p.xpad + threadIdx.x) * 4;
        for (int r = 0; r < p.xw; ++r) *reinterpret_cast<float*>(p.xbase[r] + at) = static_cast<float>(v);
      }
    }
    if (threadIdx.x == 0) *p.counter = 0;
    if (p.xw > 0) {
      // all blocks' pushes happen-before their counter increment (fence.sys above), which this block
      // observed; its own scalar pushes are ordered by the barrier before the release stores (one per peer)
      __syncthreads();
      if (threadIdx.x < p.xw) {
        __threadfence_system();
        const uint32_t step = *p.xepoch + 1u;
        uint32_t* flag = reinterpret_cast<uint32_t*>(p.xbase[threadIdx.x] + p.xlflag_off) + p.xrank;
        st_release_sys(flag, step);
      }
    }
  }
}

// ---- after the cross-rank exchange ----------------------------------------------------------------
// One launch instead of a permute copy, a sum, three scalar ops and the backward's lse min/max:
// `gathered` holds, per rank, [lse_img (pad) | lse_txt (pad) | 8 partial scalars].  Writes the
// rank-major [2][N] lse table the backward reads, the reduced results
//   out[0] loss = (S0 + S1) / 2N   out[1] dloss/ds = (S2 + S3) / 2N   out[2], out[3] accuracies = S4/N, S5/N
// and the order-preserving int encodings of min / max over all lse values (kernel (3)'s one-exp form).
// A single block: 2N floats are at most a few hundred KB, the loads are independent.
//
// Push mode (lflags != nullptr, nans_clip_loss_exchange_finish_xchg): `gathered` is slot 0 of this rank's
// own lse table, filled by the peers' finalize kernels over NVLink; the block first waits for the W
// arrival flags of step *epoch + 1, reads the slot of that step, and at the end publishes the step
// (step_out for this call's backward, *epoch for the next forward).
// Several blocks (gridDim.x <= 32) share the copy; `scratch` (null for one block) is a zero-initialised
// counter word followed, at word 2, by one (min, max) pair per block: the last block to finish reduces them.
__global__ void __launch_bounds__(1024) clip_exchange_finish_kernel(const float* __restrict__ gathered, int W,
                                                                    int n_loc, int pad, float* __restrict__ lse_all,
                                                                    long long ld, float* __restrict__ out,
                                                                    int* __restrict__ minmax,
                                                                    const uint32_t* __restrict__ lflags,
                                                                    uint32_t* __restrict__ epoch,
                                                                    uint32_t* __restrict__ step_out,
                                                                    uint32_t* __restrict__ scratch) {
  __shared__ float slo[32], shi[32];
  __shared__ bool is_last;
  const long long L = 2ll * pad + 8;
  const int total = W * n_loc;
  const int nb = gridDim.x;
  uint32_t step = 0;
  if (lflags != nullptr) {
    step = *epoch + 1u;
    if (threadIdx.x < W) wait_flag_sys(lflags + threadIdx.x, step);
    __syncthreads();
    gathered += static_cast<long long>(step & 1u) * W * L;
  }
  float lo = INFINITY, hi = -INFINITY;
#pragma unroll
  for (int strip = 0; strip < 2; ++strip) {
#pragma unroll 4
    for (int idx = blockIdx.x * 1024 + threadIdx.x; idx < total; idx += nb * 1024) {
      const int rank = idx / n_loc;
      const int row = idx - rank * n_loc;
      const float v = __ldcg(gathered + rank * L + static_cast<long long>(strip) * pad + row);  // L2: peer-written
      lse_all[strip * ld + idx] = v;
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
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
  }
  if (nb > 1) {
    float* part = reinterpret_cast<float*>(scratch + 2);
    if (threadIdx.x == 0) {
      part[2 * blockIdx.x] = lo;
      part[2 * blockIdx.x + 1] = hi;
      __threadfence();
      is_last = atomicAdd(scratch, 1u) == static_cast<unsigned>(nb - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 32) {
      lo = threadIdx.x < nb ? *(volatile float*)&part[2 * threadIdx.x] : INFINITY;
      hi = threadIdx.x < nb ? *(volatile float*)&part[2 * threadIdx.x + 1] : -INFINITY;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (threadIdx.x == 0) *scratch = 0;
    }
  }
  if (threadIdx.x < 32) {
    // lanes 0..5: sum of partial scalar k over the ranks, in rank order (deterministic)
    float sum = 0.f;
    if (threadIdx.x < 6)
      for (int r = 0; r < W; ++r) sum += __ldcg(gathered + r * L + 2ll * pad + threadIdx.x);
    const float s01 = sum + __shfl_down_sync(0xffffffffu, sum, 1);  // lanes 0, 2: S0+S1, S2+S3
    const float n = static_cast<float>(total);
    if (threadIdx.x == 0) {
      out[0] = s01 / (2.0f * n);
      const int il = __float_as_int(lo), ih = __float_as_int(hi);
      minmax[0] = il >= 0 ? il : il ^ 0x7fffffff;
      minmax[1] = ih >= 0 ? ih : ih ^ 0x7fffffff;
    }
    if (threadIdx.x == 2) out[1] = s01 / (2.0f * n);
    if (threadIdx.x == 4) out[2] = sum / n;
    if (threadIdx.x == 5) out[3] = sum / n;
    if (lflags != nullptr && threadIdx.x == 0) {
      *step_out = step;
      *epoch = step;
    }
  }
}

// ---- workspace layout ------------------------------------------------------------------------
// [diag 2*n_loc][block partials][counter][slot 0][slot 1]...   slot = {m, l, w, bv, bi} x 2*n_loc
// Slot addresses do not depend on how many slots follow, so phases need not know the total.
struct FwdWs {
  float *part_m, *part_l, *part_w, *part_bv, *diag;
  int* part_bi;
  double* block_part;
  unsigned* counter;
  long long slot_stride;  // floats between the same array of consecutive slots
  size_t bytes;
};

FwdWs carve_fwd_ws(void* ws, int64_t n_loc, int64_t total_slots) {
  FwdWs w;
  uint8_t* p = static_cast<uint8_t*>(ws);
  size_t off = 0;
  auto take = [&](size_t n) {
    uint8_t* q = p ? p + off : nullptr;
    off += n;
    return q;
  };
  const size_t part = align_up(static_cast<size_t>(2) * n_loc * 4, 256);
  w.diag = reinterpret_cast<float*>(take(part));
  const size_t nblocks = static_cast<size_t>(ceil_div(2 * n_loc, 256));
  w.block_part = reinterpret_cast<double*>(take(align_up(nblocks * 6 * 8, 256)));
  w.counter = reinterpret_cast<unsigned*>(take(256));
  w.part_m = reinterpret_cast<float*>(take(0));
  w.part_l = w.part_m ? w.part_m + part / 4 : nullptr;
  w.part_w = w.part_m ? w.part_m + 2 * (part / 4) : nullptr;
  w.part_bv = w.part_m ? w.part_m + 3 * (part / 4) : nullptr;
  w.part_bi = w.part_m ? reinterpret_cast<int*>(w.part_m + 4 * (part / 4)) : nullptr;
  w.slot_stride = static_cast<long long>(5 * (part / 4));
  off += static_cast<size_t>(total_slots) * 5 * part;
  w.bytes = off;
  return w;
}

// every launch runs as CTA pairs (cta_group::2); the single-CTA instantiation of the sweep was retired in round 2
constexpr bool fwd_pair_mode() { return true; }

int strips_of(int flags) {
  return (flags & (NANS_LOSS_STRIP_IMG | NANS_LOSS_STRIP_TXT)) ? 1 : 2;
}

int choose_nsplit(int64_t n_loc, int64_t ncols, int nstrips) {
  const bool pair = fwd_pair_mode();
  const int64_t base = nstrips * ceil_div(n_loc, pair ? 2 * BM : BM);
  const int64_t ntiles = ceil_div(ncols, BN);
  const int sms = pair ? sm_count() / 2 : sm_count();
  int best = 1;
  double best_cost = 1e300;
  const int64_t max_ns = ntiles < 32 ? ntiles : 32;
  for (int64_t ns = 1; ns <= max_ns; ++ns) {
    const double waves = static_cast<double>(ceil_div(base * ns, sms));
    const double cost = waves * (static_cast<double>(ceil_div(ntiles, ns)) + 0.75);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = static_cast<int>(ns);
    }
  }
  return best;
}

// exchange mode: a split is a sub-range of every source rank's tiles, so ns <= tiles per source.
// Every unit visits ALL sources, i.e. it cannot finish before the last source has arrived, and it holds
// its SMs while it waits: units beyond the first wave would only start after that.  (With the usual 0.75
// tiles of per-unit overhead the model chose 16 splits = 7 waves at n_loc = 4096, world 8: the first wave
// sat on the SMs until the last peer's rows were in, 400 us instead of 200.)  3 tiles per unit (pipeline
// fill, epilogue, a wave's launch) make one wave win wherever one wave is possible.
int choose_nsplit_xchg(int64_t n_loc, int64_t world) {
  const bool pair = fwd_pair_mode();
  const int64_t base = 2 * ceil_div(n_loc, pair ? 2 * BM : BM);
  const int64_t src_tiles = n_loc / BN;
  const int sms = pair ? sm_count() / 2 : sm_count();
  int best = 1;
  double best_cost = 1e300;
  const int64_t max_ns = src_tiles < 32 ? src_tiles : 32;
  for (int64_t ns = 1; ns <= max_ns; ++ns) {
    const double waves = static_cast<double>(ceil_div(base * ns, sms));
    const double cost = waves * (static_cast<double>(world * ceil_div(src_tiles, ns)) + 3.0);
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

extern "C" int64_t nans_clip_loss_fwd_xchg_slots(int64_t n_loc, int64_t world, int64_t D) {
  (void)D;
  if (n_loc <= 0 || world <= 0 || n_loc % BN != 0) return 0;
  return choose_nsplit_xchg(n_loc, world);
}

extern "C" int nans_clip_loss_fwd_xchg(const nans_xchg_t* x, const void* I16_loc, const void* T16_loc,
                                       int feat_dtype, const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                       void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(x != nullptr && x->world >= 1 && x->world <= NANS_MAX_PEERS && x->rank >= 0 && x->rank < x->world,
               "loss_fwd_xchg: bad exchange descriptor");
  NANS_REQUIRE(feat_dtype == NANS_F16 || feat_dtype == NANS_BF16, "loss_fwd_xchg: feat_dtype must be NANS_F16 or NANS_BF16");
  const int64_t n_loc = x->n_loc, D = x->D, W = x->world, N = W * n_loc;
  NANS_REQUIRE(n_loc > 0 && n_loc % BN == 0 && D > 0 && D % 8 == 0 && D <= 8192 && N < (1ll << 29),
               "loss_fwd_xchg: n_loc must be a positive multiple of 256 and D a multiple of 8");
  NANS_REQUIRE(I16_loc && T16_loc && s_dev && ws && x->base[x->rank] && x->epoch, "loss_fwd_xchg: null pointer");
  NANS_REQUIRE((flags & (NANS_LOSS_STRIP_IMG | NANS_LOSS_STRIP_TXT)) == 0, "loss_fwd_xchg: both strips run in one launch");
  const int nsplit = choose_nsplit_xchg(n_loc, W);
  const size_t need = carve_fwd_ws(nullptr, n_loc, nsplit).bytes;
  if (ws_bytes < need) {
    set_error("loss_fwd_xchg: workspace %zu < %zu bytes", ws_bytes, need);
    return NANS_ERR_WORKSPACE;
  }
  FwdWs w = carve_fwd_ws(ws, n_loc, nsplit);
  const bool pair = fwd_pair_mode();
  const int kchunks = static_cast<int>(ceil_div(D, BK));
  const SmemPlan plan = plan_smem(kchunks, pair);

  // gathered operands of this rank: per modality one [2 slots * N, D] tensor
  uint8_t* own = static_cast<uint8_t*>(x->base[x->rank]);
  const size_t mod_bytes = static_cast<size_t>(2) * N * D * 2;
  const void* I_gath = own + x->feat_off;
  const void* T_gath = own + x->feat_off + mod_bytes;
  CUtensorMap tmA0, tmB0, tmA1, tmB1;
  const uint32_t bbox = pair ? BN / 2 : BN;
  if ((rc = make_tmap_16b(&tmA0, I16_loc, feat_dtype, n_loc, D, D, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB0, T_gath, feat_dtype, 2 * N, D, D, bbox)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmA1, T16_loc, feat_dtype, n_loc, D, D, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB1, I_gath, feat_dtype, 2 * N, D, D, bbox)) != NANS_OK) return rc;

  FwdParams p;
  p.n_loc = static_cast<int>(n_loc);
  p.ncols = static_cast<int>(N);
  p.kchunks = kchunks;
  p.stages = plan.stages;
  p.idesc = make_idesc(idesc_fmt(feat_dtype), idesc_fmt(feat_dtype), 0, 0, pair ? 2 * BM : BM, BN);
  p.nrb = static_cast<int>(ceil_div(n_loc, pair ? 2 * BM : BM));
  p.nsplit = nsplit;
  p.strip0 = 0;
  p.ntiles = static_cast<int>(N / BN);
  p.skip_begin = 1 << 30;
  p.skip_count = 0;
  p.label_shift = static_cast<int>(x->rank * n_loc);
  p.lab_row0 = 0;
  p.lab_row1 = static_cast<int>(n_loc);
  p.col_global_begin = 0;
  p.s_dev = s_dev;
  p.part_m = w.part_m;
  p.part_l = w.part_l;
  p.part_w = w.part_w;
  p.part_bv = w.part_bv;
  p.part_bi = w.part_bi;
  p.diag = w.diag;
  p.slot_begin = 0;
  p.slot_stride = w.slot_stride;
  p.xw = static_cast<int>(W);
  p.xrank = x->rank;
  p.xsrc_tiles = static_cast<int>(n_loc / BN);
  p.xslot_rows = static_cast<int>(N);
  const uint32_t* fflags = reinterpret_cast<const uint32_t*>(own + x->fflag_off);
  const int64_t blocks = W * (n_loc / NANS_XCHG_FLAG_ROWS);
  p.xflags[0] = fflags + blocks;  // strip 0 reads the TEXT columns (modality 1)
  p.xflags[1] = fflags;           // strip 1 reads the IMAGE columns (modality 0)
  p.xepoch = x->epoch;
  // diagnostics: NANS_XCHG_PROBE=<device address of a zeroed int64[grid] array> (tools/xchg_probe.py)
  p.xwait_ns = nullptr;
  if (const char* e = getenv("NANS_XCHG_PROBE")) p.xwait_ns = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));

  const bool with_acc = (flags & NANS_LOSS_WITH_ACC) != 0;
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FwdParams);
  kern = plan.a_resident ? (with_acc ? clip_fwd_kernel<true, true, true> : clip_fwd_kernel<true, false, true>)
                         : (with_acc ? clip_fwd_kernel<false, true, true> : clip_fwd_kernel<false, false, true>);
  NANS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(plan.bytes)));
  const unsigned grid = static_cast<unsigned>((pair ? 2 : 1) * 2 * p.nrb * p.nsplit);
  kern<<<grid, NUM_THREADS, plan.bytes, static_cast<cudaStream_t>(stream)>>>(tmA0, tmB0, tmA1, tmB1, p);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int64_t nans_clip_loss_fwd_phase_slots(int64_t n_loc, int64_t ncols, int64_t D) {
  (void)D;
  if (n_loc <= 0 || ncols <= 0) return 0;
  return choose_nsplit(n_loc, ncols, 2);
}

extern "C" int64_t nans_clip_loss_fwd_phase_slots_flags(int64_t n_loc, int64_t ncols, int64_t D, int flags) {
  (void)D;
  if (n_loc <= 0 || ncols <= 0) return 0;
  return choose_nsplit(n_loc, ncols, strips_of(flags));
}

extern "C" size_t nans_clip_loss_fwd_workspace_bytes(int64_t n_loc, int64_t total_slots) {
  if (n_loc <= 0 || total_slots <= 0) return 256;
  return carve_fwd_ws(nullptr, n_loc, total_slots).bytes;
}

extern "C" int nans_clip_loss_fwd_phase(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                        const void* T_cols, const void* I_cols, int64_t ld_cols,
                                        int feat_dtype, int64_t n_loc, int64_t ncols, int64_t D,
                                        int64_t col_global_begin, int64_t label_begin,
                                        int64_t skip_col_begin, int64_t skip_col_count,
                                        const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                        int64_t slot_begin, void* stream) {
  return nans_clip_loss_fwd_phase_rows(I_loc, T_loc, ld_loc, T_cols, I_cols, ld_cols, feat_dtype, n_loc, ncols, D,
                                       col_global_begin, label_begin, 0, n_loc, skip_col_begin, skip_col_count,
                                       s_dev, flags, ws, ws_bytes, slot_begin, stream);
}

extern "C" int nans_clip_loss_fwd_phase_rows(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                             const void* T_cols, const void* I_cols, int64_t ld_cols,
                                             int feat_dtype, int64_t n_loc, int64_t ncols, int64_t D,
                                             int64_t col_global_begin, int64_t label_begin,
                                             int64_t label_row_begin, int64_t label_row_count,
                                             int64_t skip_col_begin, int64_t skip_col_count,
                                             const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                             int64_t slot_begin, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(feat_dtype == NANS_F16 || feat_dtype == NANS_BF16,
               "loss_fwd: feat_dtype must be NANS_F16 or NANS_BF16");
  NANS_REQUIRE(n_loc >= 0 && ncols >= 0 && D > 0 && D % 8 == 0,
               "loss_fwd: sizes must be non-negative and D a multiple of 8 (D=%lld)", (long long)D);
  NANS_REQUIRE(n_loc < (1ll << 30) && ncols < (1ll << 30) && D <= 8192, "loss_fwd: size too large");
  if (n_loc == 0 || ncols == 0) return NANS_OK;
  NANS_REQUIRE(I_loc && T_loc && T_cols && I_cols && s_dev && ws, "loss_fwd: null pointer");
  NANS_REQUIRE(ld_loc >= D && ld_cols >= D, "loss_fwd: leading dimension smaller than D");
  NANS_REQUIRE(slot_begin >= 0, "loss_fwd: negative slot");
  NANS_REQUIRE(label_row_begin >= 0 && label_row_count >= 0 && label_row_begin + label_row_count <= n_loc,
               "loss_fwd: the label row range must lie inside the local rows");
  NANS_REQUIRE(skip_col_begin >= 0 && skip_col_count >= 0 && skip_col_begin % BN == 0 &&
                   skip_col_count % BN == 0 && skip_col_begin + skip_col_count <= ncols,
               "loss_fwd: the skipped column range must be made of whole 256-column tiles inside the operand");
  if (skip_col_count == ncols) return NANS_OK;
  NANS_REQUIRE((flags & NANS_LOSS_STRIP_IMG) == 0 || (flags & NANS_LOSS_STRIP_TXT) == 0,
               "loss_fwd: NANS_LOSS_STRIP_IMG and NANS_LOSS_STRIP_TXT are exclusive (pass neither for both)");

  const int nstrips = strips_of(flags);
  const int nsplit = choose_nsplit(n_loc, ncols - skip_col_count, nstrips);
  // the caller sized the workspace for total_slots >= slot_begin + nsplit
  const size_t need = carve_fwd_ws(nullptr, n_loc, slot_begin + nsplit).bytes;
  if (ws_bytes < need) {
    set_error("loss_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return NANS_ERR_WORKSPACE;
  }
  FwdWs w = carve_fwd_ws(ws, n_loc, slot_begin + nsplit);

  const bool pair = fwd_pair_mode();
  const int kchunks = static_cast<int>(ceil_div(D, BK));
  const SmemPlan plan = plan_smem(kchunks, pair);

  CUtensorMap tmA0, tmB0, tmA1, tmB1;
  const uint32_t bbox = pair ? BN / 2 : BN;  // in pair mode a CTA loads 128 of a tile's 256 rows
  if ((rc = make_tmap_16b(&tmA0, I_loc, feat_dtype, n_loc, D, ld_loc, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB0, T_cols, feat_dtype, ncols, D, ld_cols, bbox)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmA1, T_loc, feat_dtype, n_loc, D, ld_loc, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB1, I_cols, feat_dtype, ncols, D, ld_cols, bbox)) != NANS_OK) return rc;

  FwdParams p;
  p.n_loc = static_cast<int>(n_loc);
  p.ncols = static_cast<int>(ncols);
  p.kchunks = kchunks;
  p.stages = plan.stages;
  p.idesc = make_idesc(idesc_fmt(feat_dtype), idesc_fmt(feat_dtype), 0, 0, pair ? 2 * BM : BM, BN);
  p.nrb = static_cast<int>(ceil_div(n_loc, pair ? 2 * BM : BM));
  p.nsplit = nsplit;
  p.strip0 = (flags & NANS_LOSS_STRIP_TXT) ? 1 : 0;
  p.ntiles = static_cast<int>(ceil_div(ncols, BN) - skip_col_count / BN);
  p.skip_begin = skip_col_count > 0 ? static_cast<int>(skip_col_begin / BN) : (1 << 30);
  p.skip_count = static_cast<int>(skip_col_count / BN);
  p.label_shift = static_cast<int>(label_begin - col_global_begin);
  p.lab_row0 = static_cast<int>(label_row_begin);
  p.lab_row1 = static_cast<int>(label_row_begin + label_row_count);
  p.col_global_begin = static_cast<int>(col_global_begin);
  p.s_dev = s_dev;
  p.part_m = w.part_m;
  p.part_l = w.part_l;
  p.part_w = w.part_w;
  p.part_bv = w.part_bv;
  p.part_bi = w.part_bi;
  p.diag = w.diag;
  p.slot_begin = static_cast<int>(slot_begin);
  p.slot_stride = w.slot_stride;
  p.xw = 0;
  p.xrank = p.xsrc_tiles = p.xslot_rows = 0;
  p.xflags[0] = p.xflags[1] = nullptr;
  p.xepoch = nullptr;
  p.xwait_ns = nullptr;

  const bool with_acc = (flags & NANS_LOSS_WITH_ACC) != 0;
  void (*kern)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const FwdParams);
  kern = plan.a_resident ? (with_acc ? clip_fwd_kernel<true, true, true> : clip_fwd_kernel<true, false, true>)
                         : (with_acc ? clip_fwd_kernel<false, true, true> : clip_fwd_kernel<false, false, true>);
  NANS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(plan.bytes)));
  const unsigned grid = static_cast<unsigned>((pair ? 2 : 1) * nstrips * p.nrb * p.nsplit);
  kern<<<grid, NUM_THREADS, plan.bytes, static_cast<cudaStream_t>(stream)>>>(tmA0, tmB0, tmA1, tmB1, p);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_clip_loss_fwd_finalize(int64_t n_loc, int64_t total_slots, int64_t label_begin,
                                           const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                           float* lse_img_loc, float* lse_txt_loc, float* scalars,
                                           void* stream) {
  return nans_clip_loss_fwd_finalize_rows(n_loc, total_slots, label_begin, s_dev, flags, ws, ws_bytes, lse_img_loc,
                                          lse_txt_loc, scalars, nullptr, stream);
}

extern "C" int nans_clip_loss_fwd_finalize_rows(int64_t n_loc, int64_t total_slots, int64_t label_begin,
                                                const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                                float* lse_img_loc, float* lse_txt_loc, float* scalars,
                                                float* row_stats, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(n_loc > 0 && total_slots > 0, "loss_fwd_finalize: empty problem");
  NANS_REQUIRE(s_dev && ws && lse_img_loc && lse_txt_loc && scalars, "loss_fwd_finalize: null pointer");
  FwdWs w = carve_fwd_ws(ws, n_loc, total_slots);
  if (w.bytes > ws_bytes) {
    set_error("loss_fwd_finalize: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    return NANS_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NANS_CUDA_OK(cudaMemsetAsync(w.counter, 0, sizeof(unsigned), st));
  FinParams p;
  p.n_loc = static_cast<int>(n_loc);
  p.total_slots = static_cast<int>(total_slots);
  p.with_acc = (flags & NANS_LOSS_WITH_ACC) ? 1 : 0;
  p.slot_stride = w.slot_stride;
  p.label_begin = static_cast<int>(label_begin);
  p.s_dev = s_dev;
  p.part_m = w.part_m;
  p.part_l = w.part_l;
  p.part_w = w.part_w;
  p.part_bv = w.part_bv;
  p.part_bi = w.part_bi;
  p.diag = w.diag;
  p.lse_img = lse_img_loc;
  p.lse_txt = lse_txt_loc;
  p.block_part = w.block_part;
  p.counter = w.counter;
  p.scalars = scalars;
  p.row_stats = row_stats;
  p.xw = 0;
  const unsigned grid = static_cast<unsigned>(ceil_div(2 * n_loc, 256));
  clip_fwd_finalize_kernel<<<grid, 256, 0, st>>>(p);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_clip_loss_fwd_finalize_push(const nans_xchg_t* x, int64_t total_slots, const float* s_dev,
                                                int flags, void* ws, size_t ws_bytes, float* lse_img_loc,
                                                float* lse_txt_loc, float* scalars, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(x != nullptr && x->world >= 1 && x->world <= NANS_MAX_PEERS && x->rank >= 0 && x->rank < x->world &&
                   x->n_loc > 0 && x->epoch,
               "loss_fwd_finalize_push: bad exchange descriptor");
  const int64_t n_loc = x->n_loc, pad = (n_loc + 3) / 4 * 4;
  NANS_REQUIRE(total_slots > 0 && x->lse_len == 2 * pad + 8, "loss_fwd_finalize_push: bad sizes");
  NANS_REQUIRE(s_dev && ws && lse_img_loc && lse_txt_loc && scalars, "loss_fwd_finalize_push: null pointer");
  for (int r = 0; r < x->world; ++r) NANS_REQUIRE(x->base[r] != nullptr, "loss_fwd_finalize_push: peer %d is not mapped", r);
  FwdWs w = carve_fwd_ws(ws, n_loc, total_slots);
  if (w.bytes > ws_bytes) {
    set_error("loss_fwd_finalize_push: workspace %zu < %zu bytes", ws_bytes, w.bytes);
    return NANS_ERR_WORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NANS_CUDA_OK(cudaMemsetAsync(w.counter, 0, sizeof(unsigned), st));
  FinParams p;
  p.n_loc = static_cast<int>(n_loc);
  p.total_slots = static_cast<int>(total_slots);
  p.with_acc = (flags & NANS_LOSS_WITH_ACC) ? 1 : 0;
  p.slot_stride = w.slot_stride;
  p.label_begin = static_cast<int>(x->rank * n_loc);
  p.s_dev = s_dev;
  p.part_m = w.part_m;
  p.part_l = w.part_l;
  p.part_w = w.part_w;
  p.part_bv = w.part_bv;
  p.part_bi = w.part_bi;
  p.diag = w.diag;
  p.lse_img = lse_img_loc;
  p.lse_txt = lse_txt_loc;
  p.block_part = w.block_part;
  p.counter = w.counter;
  p.scalars = scalars;
  p.row_stats = nullptr;
  p.xw = x->world;
  p.xrank = x->rank;
  p.xpad = static_cast<int>(pad);
  p.xlse_off = x->lse_off;
  p.xlse_len = x->lse_len;
  p.xlflag_off = x->lflag_off;
  p.xepoch = x->epoch;
  for (int r = 0; r < NANS_MAX_PEERS; ++r) p.xbase[r] = r < x->world ? static_cast<uint8_t*>(x->base[r]) : nullptr;
  const unsigned grid = static_cast<unsigned>(ceil_div(2 * n_loc, 256));
  clip_fwd_finalize_kernel<<<grid, 256, 0, st>>>(p);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_clip_loss_fwd(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                  const void* T_all, const void* I_all, int64_t ld_all,
                                  int feat_dtype, int64_t n_loc, int64_t N, int64_t D,
                                  int64_t label_begin, const float* s_dev, int flags,
                                  float* lse_img_loc, float* lse_txt_loc, float* scalars, void* ws,
                                  size_t ws_bytes, void* stream) {
  NANS_REQUIRE(n_loc > 0 && N > 0, "loss_fwd: empty batch");
  int rc = nans_clip_loss_fwd_phase(I_loc, T_loc, ld_loc, T_all, I_all, ld_all, feat_dtype, n_loc, N,
                                    D, 0, label_begin, 0, 0, s_dev, flags, ws, ws_bytes, 0, stream);
  if (rc != NANS_OK) return rc;
  const int64_t slots = nans_clip_loss_fwd_phase_slots(n_loc, N, D);
  return nans_clip_loss_fwd_finalize(n_loc, slots, label_begin, s_dev, flags, ws, ws_bytes,
                                     lse_img_loc, lse_txt_loc, scalars, stream);
}

extern "C" int nans_clip_loss_exchange_finish(const float* gathered, int64_t world, int64_t n_loc, int64_t pad,
                                              float* lse_all, int64_t ld, float* out, int* lse_minmax,
                                              void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(world > 0 && n_loc > 0 && pad >= n_loc && ld >= world * n_loc && world * n_loc < (1ll << 30),
               "loss_exchange_finish: bad sizes");
  NANS_REQUIRE(gathered && lse_all && out && lse_minmax, "loss_exchange_finish: null pointer");
  clip_exchange_finish_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      gathered, static_cast<int>(world), static_cast<int>(n_loc), static_cast<int>(pad), lse_all,
      static_cast<long long>(ld), out, lse_minmax, nullptr, nullptr, nullptr, nullptr);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_clip_loss_exchange_finish_xchg(const nans_xchg_t* x, float* lse_all, int64_t ld, float* out,
                                                   int* lse_minmax, uint32_t* step_out, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(x != nullptr && x->world >= 1 && x->world <= NANS_MAX_PEERS && x->rank >= 0 && x->rank < x->world &&
                   x->n_loc > 0 && x->base[x->rank] && x->epoch,
               "loss_exchange_finish_xchg: bad exchange descriptor");
  const int64_t pad = (x->n_loc + 3) / 4 * 4;
  NANS_REQUIRE(x->lse_len == 2 * pad + 8 && ld >= x->world * x->n_loc && x->world * x->n_loc < (1ll << 30),
               "loss_exchange_finish_xchg: bad sizes");
  NANS_REQUIRE(lse_all && out && lse_minmax && step_out, "loss_exchange_finish_xchg: null pointer");
  uint8_t* own = static_cast<uint8_t*>(x->base[x->rank]);
  // the flag words take the first 64 bytes of their 1 KB region; words 64.. are this kernel's scratch
  // (block counter + per-block min / max), zero from the allocation and reset by the kernel itself
  const int64_t total = x->world * x->n_loc;
  const unsigned nb = static_cast<unsigned>(total >= 16384 ? 16 : (total >= 4096 ? 4 : 1));
  uint32_t* scratch = reinterpret_cast<uint32_t*>(own + x->lflag_off) + 64;
  clip_exchange_finish_kernel<<<nb, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float*>(own + x->lse_off), x->world, static_cast<int>(x->n_loc), static_cast<int>(pad),
      lse_all, static_cast<long long>(ld), out, lse_minmax, reinterpret_cast<const uint32_t*>(own + x->lflag_off),
      x->epoch, step_out, nb > 1 ? scratch : nullptr);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}
