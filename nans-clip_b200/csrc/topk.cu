// topk.cu — kernel (4): top-k inner-product retrieval.
//
// Replaces cn_clip/eval/make_topk_predictions.py:71-85 (and make_topk_predictions_tr.py): per
// query, the k gallery rows with the largest inner product in (score desc, gallery index asc)
// order — what Python's stable sorted(score_tuples, key=score, reverse=True)[:k] yields.
//
// All sweeps use the strip_sweep.cuh skeleton (256 queries per CTA pair resident in shared memory,
// 256-row gallery tiles through tcgen05.mma into TMEM, two epilogue warp sets alternating tiles).
//
// Floor pass (topk_floor_kernel, galleries >= 48K rows): over the first 16K gallery rows each
//   epilogue thread keeps only the k_cand largest 32-column CHUNK MAXIMA of its query (values only,
//   branch-free insert).  The k_cand-th largest of them over both threads of a query is a lower
//   bound ("floor") of the query's k_cand-th best score in the shard.  MMA-bound.
// Main pass (topk_sweep_kernel): unit = (query block, gallery split).  The epilogue thread keeps a
//   sorted k_cand-entry (score, index) list in registers that only admits scores >= floor; per
//   32-column chunk one max tree + one compare + one warp vote; on the rare path each lane builds
//   a bit mask of its qualifying columns and the lanes insert in lockstep.  Lists go to the
//   workspace, one per (split, tile parity, query).  Without a floor a cold list spends its first
//   ~16K columns almost entirely on the insertion path (22 % tensor pipe measured).
// Finalize (topk_finalize_kernel): one warp per query k-way-merges the sorted lists, keeps the best
//   k_cand by 16-bit score, recomputes those scores exactly in fp32 from the fp32 copies, and emits
//   the top k in the reference order.  (16-bit scores alone flip near-ties)
//
// Roofline: tensor cores, 2*Q*G*D flops; the candidate lists are O(Q * k_cand) bytes.
#include <limits.h>
#include <stdlib.h>

#include "strip_sweep.cuh"

namespace nans {
namespace {

using namespace sweep;

template <int KC>
struct TopkEpi {
  float ls[KC];
  int li[KC];
  int ncols;
  int col_offset;  // index of column 0 of this pass inside the shard
  float floor_;    // nothing at or below this score can reach the final top k_cand (pre-pass)

  __device__ __forceinline__ void init(int ncols_, int col_offset_, float floor) {
    ncols = ncols_;
    col_offset = col_offset_;
    floor_ = floor;
#pragma unroll
    for (int i = 0; i < KC; ++i) {
      ls[i] = -INFINITY;
      li[i] = -1;
    }
  }

  // precondition: s > ls[KC-1].  Entries >= s keep their place (earlier index first on ties).
  __device__ __forceinline__ void insert(float s, int idx) {
#pragma unroll
    for (int k = KC - 1; k >= 1; --k) {
      const bool keep = ls[k] >= s;
      const bool here = !keep && (ls[k - 1] >= s);
      const float ns = keep ? ls[k] : (here ? s : ls[k - 1]);
      const int ni = keep ? li[k] : (here ? idx : li[k - 1]);
      ls[k] = ns;
      li[k] = ni;
    }
    if (!(ls[0] >= s)) {
      ls[0] = s;
      li[0] = idx;
    }
  }

  __device__ __forceinline__ void tile(uint32_t taddr, int tile_idx) {
    const int col0 = tile_idx * BN;
    const bool tail = col0 + BN > ncols;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      const int cb = col0 + ch * 32;
      if (cb >= ncols) break;
      uint32_t r[32];
      tmem_ld32(taddr + ch * 32, r);
      tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
      if (tail) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (cb + k >= ncols) v[k] = -INFINITY;
      }
      // common case: nothing in the chunk beats the list's minimum -> one max tree + one compare
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(fmaxf(v[i], v[i + 8]), fmaxf(v[i + 16], v[i + 24]));
      const float m = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                            fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      if (__any_sync(0xffffffffu, m > fmaxf(ls[KC - 1], floor_))) {
        // rare path, entered by the whole warp and kept small (a 32x unrolled insertion thrashed
        // the instruction cache): each lane collects a bit mask of its qualifying columns, then all
        // lanes insert their own next candidate in lockstep, in ascending column order so that
        // equal scores keep the lower index first.
        const float thr = fmaxf(ls[KC - 1], floor_);
        uint32_t mask = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) mask |= (v[k] > thr) ? (1u << k) : 0u;
        while (__any_sync(0xffffffffu, mask != 0)) {
          if (mask != 0) {
            const int k = __ffs(mask) - 1;
            mask &= mask - 1;
            // x = v[k] through a 5-level select tree (register arrays cannot be indexed)
            float t16[16], t8[8], t4[4], t2[2];
#pragma unroll
            for (int i = 0; i < 16; ++i) t16[i] = (k & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
            for (int i = 0; i < 8; ++i) t8[i] = (k & 2) ? t16[2 * i + 1] : t16[2 * i];
#pragma unroll
            for (int i = 0; i < 4; ++i) t4[i] = (k & 4) ? t8[2 * i + 1] : t8[2 * i];
#pragma unroll
            for (int i = 0; i < 2; ++i) t2[i] = (k & 8) ? t4[2 * i + 1] : t4[2 * i];
            const float x = (k & 16) ? t2[1] : t2[0];
            if (x > ls[KC - 1]) insert(x, col_offset + cb + k);
          }
        }
      }
    }
  }
};


// Floor pass epilogue: per row, the KC largest CHUNK MAXIMA (32-column chunks) of a sample of the
// gallery.  The KC-th of them is a lower bound of the row's KC-th best score (KC distinct chunks
// each hold an element at least that large).  One value per chunk, values only, no votes, no
// divergent loops: the pass stays MMA-bound, unlike a cold candidate list.
template <int KC>
struct FloorEpi {
  float lv[KC];
  int ncols;

  __device__ __forceinline__ void init(int ncols_) {
    ncols = ncols_;
#pragma unroll
    for (int i = 0; i < KC; ++i) lv[i] = -INFINITY;
  }

  __device__ __forceinline__ void tile(uint32_t taddr, int tile_idx) {
    const int col0 = tile_idx * BN;
    const bool tail = col0 + BN > ncols;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      const int cb = col0 + ch * 32;
      if (cb >= ncols) break;
      uint32_t r[32];
      tmem_ld32(taddr + ch * 32, r);
      tmem_wait_ld();
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
      if (tail) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (cb + k >= ncols) v[k] = -INFINITY;
      }
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(fmaxf(v[i], v[i + 8]), fmaxf(v[i + 16], v[i + 24]));
      const float m = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                            fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      // sorted insert of m, branch-free: new lv[k] = median-like pick of (lv[k-1], lv[k], m)
      //   lv[k] >= m -> lv[k];  lv[k-1] >= m > lv[k] -> m;  m > lv[k-1] -> lv[k-1]
      // = max(lv[k], min(lv[k-1], m)) for a descending list; all steps read the old values.
#pragma unroll
      for (int k = KC - 1; k >= 1; --k) lv[k] = fmaxf(lv[k], fminf(lv[k - 1], m));
      lv[0] = fmaxf(lv[0], m);
    }
  }
};

struct TopkParams {
  int Q, G, kchunks, stages;
  uint32_t idesc;
  int nqb, nsplit, ntiles;
  float* cand_s;  // [lists][Q][KC]  (one list per gallery split and tile parity)
  int* cand_i;
  int list_base;     // first list index written by this launch
  int col_offset;    // shard row of gallery column 0 of this launch
  int floor_lists;   // > 0: per-row floor = max over floor_vals[0 .. floor_lists)[row]
  float* floor_vals; // [floor lists <= 8][Q][KC]: sorted chunk maxima from the floor pass
};

template <bool A_RES, int KC, bool CP>
__global__ void __cluster_dims__(CP ? 2 : 1, 1, 1) __launch_bounds__(NUM_THREADS, 1)
topk_sweep_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG,
                  const TopkParams p) {
  const int unit = CP ? (blockIdx.x >> 1) : blockIdx.x;
  // split-major unit order: the CTA pairs that run together belong to the SAME gallery split and walk its
  // tiles in lockstep, so a tile is fetched from HBM once per wave instead of once per query block
  // (query-block-major order: ncu dram__bytes_read 17.7 GB for a 1 GB gallery at Q = 30000, L2 hit 79 %)
  const int split = unit / p.nqb;
  const int qb = unit - split * p.nqb;

  SweepArgs a;
  a.tmA = &tmQ;
  a.tmB = &tmG;
  a.row0 = CP ? qb * 2 * BM + static_cast<int>(blockIdx.x & 1) * BM : qb * BM;
  a.tile_begin = static_cast<int>(static_cast<long long>(split) * p.ntiles / p.nsplit);
  a.tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / p.nsplit);
  a.skip_begin = 1 << 30;
  a.skip_count = 0;
  a.kchunks = p.kchunks;
  a.stages = p.stages;
  a.idesc = p.idesc;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = a.row0 + (warp & 3) * 32 + lane;

  TopkEpi<KC> epi;
  float floor = -INFINITY;
  if (p.floor_lists > 0 && row < p.Q) {
    // the floor pass has completed (stream order): each of its values bounds the row's k_cand-th
    // best score over the whole shard from below
    // KC-th largest of the union of the floor lists (each sorted descending): KC rounds of
    // "take the largest head".  Distinct chunks back distinct values, so the bound stays valid.
    int head[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) head[l] = 0;
    for (int r = 0; r < KC; ++r) {
      float best = -INFINITY;
      int bl = 0;
#pragma unroll
      for (int l = 0; l < 8; ++l) {
        if (l < p.floor_lists && head[l] < KC) {
          const float v = __ldg(p.floor_vals + (static_cast<long long>(l) * p.Q + row) * KC + head[l]);
          if (v > best) {
            best = v;
            bl = l;
          }
        }
      }
#pragma unroll
      for (int l = 0; l < 8; ++l)
        if (l == bl) ++head[l];
      floor = best;
    }
    // candidates are the columns with score >= floor: lower it by a hair and keep the strict test
    floor = floor - fabsf(floor) * 2.4e-7f - 1e-37f;
  }
  epi.init(p.G, p.col_offset, floor);
  run<A_RES, CP>(a, epi);

  if (warp >= 4 && row < p.Q) {
    const int half = (warp - 4) >> 2;  // each row has one list per epilogue set (even / odd tiles)
    const long long base = (static_cast<long long>(p.list_base + split * 2 + half) * p.Q + row) * KC;
#pragma unroll
    for (int i = 0; i < KC; i += 4) {
      *reinterpret_cast<float4*>(p.cand_s + base + i) =
          make_float4(epi.ls[i], epi.ls[i + 1], epi.ls[i + 2], epi.ls[i + 3]);
      *reinterpret_cast<int4*>(p.cand_i + base + i) =
          make_int4(epi.li[i], epi.li[i + 1], epi.li[i + 2], epi.li[i + 3]);
    }
  }
}

template <bool A_RES, int KC, bool CP>
__global__ void __cluster_dims__(CP ? 2 : 1, 1, 1) __launch_bounds__(NUM_THREADS, 1)
topk_floor_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG,
                  const TopkParams p) {
  const int unit = CP ? (blockIdx.x >> 1) : blockIdx.x;
  const int split = unit % p.nsplit;
  const int qb = unit / p.nsplit;

  SweepArgs a;
  a.tmA = &tmQ;
  a.tmB = &tmG;
  a.row0 = CP ? qb * 2 * BM + static_cast<int>(blockIdx.x & 1) * BM : qb * BM;
  a.tile_begin = static_cast<int>(static_cast<long long>(split) * p.ntiles / p.nsplit);
  a.tile_end = static_cast<int>(static_cast<long long>(split + 1) * p.ntiles / p.nsplit);
  a.skip_begin = 1 << 30;
  a.skip_count = 0;
  a.kchunks = p.kchunks;
  a.stages = p.stages;
  a.idesc = p.idesc;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = a.row0 + (warp & 3) * 32 + lane;

  FloorEpi<KC> epi;
  epi.init(p.G);
  run<A_RES, CP>(a, epi);

  if (warp >= 4 && row < p.Q) {
    const int set = (warp - 4) >> 2;
    float* dst = p.floor_vals + (static_cast<long long>(split * 2 + set) * p.Q + row) * KC;
#pragma unroll
    for (int i = 0; i < KC; i += 4)
      *reinterpret_cast<float4*>(dst + i) = make_float4(epi.lv[i], epi.lv[i + 1], epi.lv[i + 2], epi.lv[i + 3]);
  }
}

// (score desc, index asc); invalid entries (index < 0) sort last
__device__ __forceinline__ bool beats(float sa, long long ia, float sb, long long ib) {
  const bool va = ia >= 0, vb = ib >= 0;
  if (va != vb) return va;
  if (!va) return false;
  return sa > sb || (sa == sb && ia < ib);
}

constexpr int FIN_WARPS = 4;
constexpr int FIN_MAXC = 1024;  // nsplit (<= 16) * 2 halves * k_cand (<= 32)
constexpr int MERGE_MAXC = 512;  // n_shards * k

struct FinParams {
  int Q, G, D, k, kc, nsplit;
  const float* cand_s;
  const int* cand_i;
  const float* Q32;
  const float* G32;
  long long index_offset;
  float* out_scores;
  long long* out_index;
};

__global__ void __launch_bounds__(FIN_WARPS * 32) topk_finalize_kernel(const FinParams p) {
  __shared__ float sel_s[FIN_WARPS][32];
  __shared__ int sel_i[FIN_WARPS][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * FIN_WARPS + w;
  if (q >= p.Q) return;
  // k-way merge of the (sorted) candidate lists: lane l owns list l and offers its head; kc rounds
  // of a warp arg-max in (score desc, index asc) order pick the best kc candidates by 16-bit score.
  const int nl = p.nsplit;  // candidate lists per query (<= 32)
  if (lane < 32) {
    sel_s[w][lane] = -INFINITY;
    sel_i[w][lane] = -1;
  }
  int pos = 0;
  const long long lbase = (static_cast<long long>(lane) * p.Q + q) * p.kc;
  float hs = -INFINITY;
  int hi = -1;
  if (lane < nl) {
    hs = p.cand_s[lbase];
    hi = p.cand_i[lbase];
  }
  for (int r = 0; r < p.kc; ++r) {
    float bs = hs;
    int bi = hi, bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (beats(os, oi, bs, bi) || (oi == bi && os == bs && ol < bl)) {
        bs = os;
        bi = oi;
        bl = ol;
      }
    }
    if (bi < 0) break;  // nothing left anywhere (uniform)
    if (lane == 0) {
      sel_s[w][r] = bs;
      sel_i[w][r] = bi;
    }
    if (lane == bl) {
      ++pos;
      if (pos < p.kc) {
        hs = p.cand_s[lbase + pos];
        hi = p.cand_i[lbase + pos];
      } else {
        hs = -INFINITY;
        hi = -1;
      }
    }
  }
  __syncwarp();
  // exact fp32 rescoring of the selected candidates
  if (p.Q32 != nullptr) {
    const float4* qv = reinterpret_cast<const float4*>(p.Q32 + static_cast<long long>(q) * p.D);
    for (int r = 0; r < p.kc; ++r) {
      const int gi = sel_i[w][r];
      if (gi < 0) continue;  // warp-uniform
      const float4* gv = reinterpret_cast<const float4*>(p.G32 + static_cast<long long>(gi) * p.D);
      float acc = 0.f;
      for (int v = lane; v < p.D / 4; v += 32) {
        const float4 a = __ldg(qv + v), b = __ldg(gv + v);
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) sel_s[w][r] = acc;
    }
    __syncwarp();
  }
  // final order among the kc rescored candidates
  if (lane < p.kc) {
    const float s = sel_s[w][lane];
    const int i = sel_i[w][lane];
    if (i >= 0) {
      int rank = 0;
      for (int d = 0; d < p.kc; ++d) rank += beats(sel_s[w][d], sel_i[w][d], s, i) ? 1 : 0;
      if (rank < p.k) {
        p.out_scores[static_cast<long long>(q) * p.k + rank] = s;
        p.out_index[static_cast<long long>(q) * p.k + rank] = p.index_offset + i;
      }
    }
  }
  // A query can end up with fewer candidates than min(G, k): NaN scores (a zero-norm row normalised to
  // NaN by extract_features.py) never pass `v > threshold`.  The reference always emits k valid ids
  // (make_topk_predictions.py:84-85; for an all-NaN row its stable sort leaves gallery order), so the
  // open ranks are filled with the lowest gallery rows not selected yet, score -inf — never left
  // unwritten (the outputs are torch.empty) and never -1 while the shard still has rows.
  __syncwarp();
  int nsel = 0;
  for (int d = 0; d < p.kc; ++d) nsel += sel_i[w][d] >= 0 ? 1 : 0;
  const int nvalid = p.G < p.k ? p.G : p.k;
  if (nsel < nvalid && lane == 0) {
    int cand = 0;
    for (int r = nsel; r < nvalid; ++r) {
      for (;; ++cand) {
        bool used = false;
        for (int d = 0; d < p.kc; ++d) used |= sel_i[w][d] == cand;
        if (!used) break;
      }
      p.out_scores[static_cast<long long>(q) * p.k + r] = -INFINITY;
      p.out_index[static_cast<long long>(q) * p.k + r] = p.index_offset + cand;
      ++cand;
    }
  }
  // pad when the shard has fewer than k rows
  if (lane >= nvalid && lane < p.k) {
    p.out_scores[static_cast<long long>(q) * p.k + lane] = -INFINITY;
    p.out_index[static_cast<long long>(q) * p.k + lane] = -1;
  }
}

struct MergeParams {
  int n_shards, Q, k;
  const float* in_s;
  const long long* in_i;
  float* out_s;
  long long* out_i;
};

// one warp per query; n_shards * k <= MERGE_MAXC candidates
__global__ void __launch_bounds__(FIN_WARPS * 32) topk_merge_kernel(const MergeParams p) {
  __shared__ float sh_s[FIN_WARPS][MERGE_MAXC];
  __shared__ long long sh_i[FIN_WARPS][MERGE_MAXC];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * FIN_WARPS + w;
  if (q >= p.Q) return;
  const int C = p.n_shards * p.k;
  for (int c = lane; c < C; c += 32) {
    const int sh = c / p.k, j = c - sh * p.k;
    const long long src = (static_cast<long long>(sh) * p.Q + q) * p.k + j;
    sh_s[w][c] = p.in_s[src];
    sh_i[w][c] = p.in_i[src];
  }
  __syncwarp();
  int nvalid = 0;
  for (int c = lane; c < C; c += 32) {
    const float s = sh_s[w][c];
    const long long i = sh_i[w][c];
    if (i < 0) continue;
    ++nvalid;
    int rank = 0;
    for (int d = 0; d < C; ++d) rank += beats(sh_s[w][d], sh_i[w][d], s, i) ? 1 : 0;
    if (rank < p.k) {
      p.out_s[static_cast<long long>(q) * p.k + rank] = s;
      p.out_i[static_cast<long long>(q) * p.k + rank] = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
  for (int r = nvalid + lane; r < p.k; r += 32) {
    p.out_s[static_cast<long long>(q) * p.k + r] = -INFINITY;
    p.out_i[static_cast<long long>(q) * p.k + r] = -1;
  }
}

// every launch runs as CTA pairs (cta_group::2); the single-CTA instantiation was retired in round 2
constexpr bool topk_pair_mode() { return true; }

// A cold candidate list spends its first ~16K columns almost entirely on the insertion path (an
// exact cold pre-pass over 16K columns ran at 22 % tensor pipe).  So a long gallery is swept twice:
//   * FLOOR pass over the first kFloorCols columns: per row only the k_cand-th largest 32-column
//     chunk maximum (cheap, MMA-bound) — a lower bound of the row's k_cand-th best score;
//   * MAIN pass over all columns, whose lists only admit scores >= floor (about
//     G * k_cand / kFloorCols columns per query).
// Without a floor every split is charged a warm-up worth about 12 tiles of MMA time.
constexpr int64_t kFloorColsDefault = 16384;
constexpr int kFloorSplitsMax = 4;

int64_t floor_cols() {
  const char* e = getenv("NANS_TOPK_FLOOR_COLS");
  const int64_t v = e ? atoll(e) : kFloorColsDefault;
  return v < 1024 ? 1024 : (v / 256) * 256;
}
int floor_splits() {
  const char* e = getenv("NANS_TOPK_FLOOR_SPLITS");
  const int v = e ? atoi(e) : 1;
  return v < 1 ? 1 : (v > kFloorSplitsMax ? kFloorSplitsMax : v);
}

int choose_topk_nsplit(int64_t Q, int64_t G, bool warm) {
  const bool pair = topk_pair_mode();
  const int64_t base = ceil_div(Q, pair ? 2 * BM : BM);
  const int64_t ntiles = ceil_div(G, BN);
  const int sms = pair ? sm_count() / 2 : sm_count();
  int best = 1;
  double best_cost = 1e300;
  const int64_t max_ns = ntiles / 16 < 1 ? 1 : (ntiles / 16 > 14 ? 14 : ntiles / 16);
  for (int64_t ns = 1; ns <= max_ns; ++ns) {
    const double waves = static_cast<double>(ceil_div(base * ns, sms));
    const double cost = waves * (static_cast<double>(ceil_div(ntiles, ns)) + (warm ? 1.5 : 12.0));
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = static_cast<int>(ns);
    }
  }
  return best;
}

struct TopkPlan {
  bool floor_pass;
  int ns_main;  // splits of the main pass
  int lists;    // candidate lists per query
};

TopkPlan plan_topk(int64_t Q, int64_t G) {
  TopkPlan t;
  t.floor_pass = G >= 3 * floor_cols();
  t.ns_main = choose_topk_nsplit(Q, G, t.floor_pass);
  t.lists = 2 * t.ns_main;
  return t;
}

}  // namespace
}  // namespace nans

using namespace nans;

extern "C" size_t nans_topk_ip_workspace_bytes(int64_t Q, int64_t G, int64_t D, int k_cand) {
  (void)D;
  if (Q <= 0 || G <= 0 || k_cand <= 0) return 256;
  const TopkPlan t = plan_topk(Q, G);
  return 2 * align_up(static_cast<size_t>(t.lists) * Q * k_cand * 4, 256) +
         align_up(static_cast<size_t>(2 * kFloorSplitsMax) * Q * k_cand * 4, 256) + 256;
}

extern "C" int nans_topk_ip(const void* Q16, const void* G16, int feat_dtype, const float* Q32,
                            const float* G32, int64_t Q, int64_t G, int64_t D, int k, int k_cand,
                            int64_t gallery_index_offset, float* out_scores, int64_t* out_index,
                            void* ws, size_t ws_bytes, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(feat_dtype == NANS_F16 || feat_dtype == NANS_BF16,
               "topk_ip: feat_dtype must be NANS_F16 or NANS_BF16");
  NANS_REQUIRE(k_cand == 16 || k_cand == 32, "topk_ip: k_cand must be 16 or 32 (got %d)", k_cand);
  NANS_REQUIRE(k >= 1 && k <= k_cand, "topk_ip: need 1 <= k <= k_cand (k=%d)", k);
  NANS_REQUIRE(Q >= 0 && G >= 0 && D > 0 && D % 8 == 0, "topk_ip: bad sizes (D must be a multiple of 8)");
  NANS_REQUIRE(Q < (1ll << 30) && G < (1ll << 31) - 512 && D <= 8192, "topk_ip: size too large");
  NANS_REQUIRE((Q32 == nullptr) == (G32 == nullptr), "topk_ip: pass both fp32 copies or neither");
  if (Q == 0) return NANS_OK;
  NANS_REQUIRE(out_scores && out_index, "topk_ip: null output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (G == 0) {
    // every slot is padding: -inf / -1 (0xFF bytes are -1 for int64; scores need a kernel -> reuse merge)
    MergeParams mp{0, static_cast<int>(Q), k, nullptr, nullptr, out_scores,
                   reinterpret_cast<long long*>(out_index)};
    topk_merge_kernel<<<static_cast<unsigned>(ceil_div(Q, FIN_WARPS)), FIN_WARPS * 32, 0, st>>>(mp);
    NANS_CUDA_OK(cudaGetLastError());
    return NANS_OK;
  }
  NANS_REQUIRE(Q16 && G16 && ws, "topk_ip: null pointer");
  NANS_REQUIRE(Q32 == nullptr || ((reinterpret_cast<uintptr_t>(Q32) & 15) == 0 &&
                                  (reinterpret_cast<uintptr_t>(G32) & 15) == 0 && D % 4 == 0),
               "topk_ip: fp32 copies must be 16-byte aligned");
  const size_t need = nans_topk_ip_workspace_bytes(Q, G, D, k_cand);
  if (ws_bytes < need) {
    set_error("topk_ip: workspace %zu < %zu bytes", ws_bytes, need);
    return NANS_ERR_WORKSPACE;
  }
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "topk_ip: workspace must be 16-byte aligned");

  const TopkPlan plan_t = plan_topk(Q, G);
  const int kchunks = static_cast<int>(ceil_div(D, BK));
  const bool pair = topk_pair_mode();
  const SmemPlan plan = plan_smem(kchunks, pair);

  void (*kern)(const CUtensorMap, const CUtensorMap, const TopkParams);
  if (k_cand == 16) kern = plan.a_resident ? topk_sweep_kernel<true, 16, true> : topk_sweep_kernel<false, 16, true>;
  else kern = plan.a_resident ? topk_sweep_kernel<true, 32, true> : topk_sweep_kernel<false, 32, true>;
  NANS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(plan.bytes)));
  void (*fkern)(const CUtensorMap, const CUtensorMap, const TopkParams);
  if (k_cand == 16) fkern = plan.a_resident ? topk_floor_kernel<true, 16, true> : topk_floor_kernel<false, 16, true>;
  else fkern = plan.a_resident ? topk_floor_kernel<true, 32, true> : topk_floor_kernel<false, 32, true>;
  NANS_CUDA_OK(cudaFuncSetAttribute(fkern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(plan.bytes)));

  CUtensorMap tmQ;
  if ((rc = make_tmap_16b(&tmQ, Q16, feat_dtype, Q, D, D, BM)) != NANS_OK) return rc;
  float* cand_s = static_cast<float*>(ws);
  const size_t list_bytes = align_up(static_cast<size_t>(plan_t.lists) * Q * k_cand * 4, 256);
  int* cand_i = reinterpret_cast<int*>(static_cast<uint8_t*>(ws) + list_bytes);
  float* floor_vals = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 2 * list_bytes);

  // launch one sweep over gallery rows [g0, g0 + gn)
  auto sweep = [&](bool floor_only, int64_t g0, int64_t gn, int nsplit, int list_base, int floor_lists) -> int {
    CUtensorMap tmG;
    const uint8_t* gbase = static_cast<const uint8_t*>(G16) + static_cast<size_t>(g0) * D * 2;
    int rc2 = make_tmap_16b(&tmG, gbase, feat_dtype, gn, D, D, pair ? BN / 2 : BN);
    if (rc2 != NANS_OK) return rc2;
    TopkParams p;
    p.Q = static_cast<int>(Q);
    p.G = static_cast<int>(gn);
    p.kchunks = kchunks;
    p.stages = plan.stages;
    p.idesc = make_idesc(idesc_fmt(feat_dtype), idesc_fmt(feat_dtype), 0, 0, pair ? 2 * BM : BM, BN);
    p.nqb = static_cast<int>(ceil_div(Q, pair ? 2 * BM : BM));
    p.nsplit = nsplit;
    p.ntiles = static_cast<int>(ceil_div(gn, BN));
    p.cand_s = cand_s;
    p.cand_i = cand_i;
    p.list_base = list_base;
    p.col_offset = static_cast<int>(g0);
    p.floor_lists = floor_lists;
    p.floor_vals = floor_vals;
    const unsigned grid = static_cast<unsigned>((pair ? 2 : 1) * p.nqb * p.nsplit);
    if (floor_only) fkern<<<grid, NUM_THREADS, plan.bytes, st>>>(tmQ, tmG, p);
    else kern<<<grid, NUM_THREADS, plan.bytes, st>>>(tmQ, tmG, p);
    if (cudaGetLastError() != cudaSuccess) {
      set_error("topk_ip: sweep launch failed");
      return NANS_ERR_CUDA;
    }
    return NANS_OK;
  };
  int dbg = 0;  // NANS_TOPK_DEBUG (timing experiments): 1 = no floor pass launch, 2 = no main pass, 4 = no finalize
  {
    const char* e = getenv("NANS_TOPK_DEBUG");
    if (e) dbg = atoi(e);
  }
  if (plan_t.floor_pass) {
    if (!(dbg & 1) && (rc = sweep(true, 0, floor_cols(), floor_splits(), 0, 0)) != NANS_OK) return rc;
    if (!(dbg & 2) && (rc = sweep(false, 0, G, plan_t.ns_main, 0, 2 * floor_splits())) != NANS_OK) return rc;
  } else {
    if ((rc = sweep(false, 0, G, plan_t.ns_main, 0, 0)) != NANS_OK) return rc;
  }

  FinParams f;
  f.Q = static_cast<int>(Q);
  f.G = static_cast<int>(G);
  f.D = static_cast<int>(D);
  f.k = k;
  f.kc = k_cand;
  f.nsplit = plan_t.lists;
  f.cand_s = cand_s;
  f.cand_i = cand_i;
  f.Q32 = Q32;
  f.G32 = G32;
  f.index_offset = gallery_index_offset;
  f.out_scores = out_scores;
  f.out_index = reinterpret_cast<long long*>(out_index);
  if (!(dbg & 4)) topk_finalize_kernel<<<static_cast<unsigned>(ceil_div(Q, FIN_WARPS)), FIN_WARPS * 32, 0, st>>>(f);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_topk_merge(const float* in_scores, const int64_t* in_index, int n_shards,
                               int64_t Q, int k, float* out_scores, int64_t* out_index,
                               void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(n_shards >= 1 && k >= 1 && static_cast<int64_t>(n_shards) * k <= MERGE_MAXC,
               "topk_merge: need n_shards * k <= %d", MERGE_MAXC);
  NANS_REQUIRE(Q >= 0 && Q < (1ll << 30), "topk_merge: bad Q");
  if (Q == 0) return NANS_OK;
  NANS_REQUIRE(in_scores && in_index && out_scores && out_index, "topk_merge: null pointer");
  MergeParams p{n_shards, static_cast<int>(Q), k, in_scores,
                reinterpret_cast<const long long*>(in_index), out_scores,
                reinterpret_cast<long long*>(out_index)};
  topk_merge_kernel<<<static_cast<unsigned>(ceil_div(Q, FIN_WARPS)), FIN_WARPS * 32, 0,
                      static_cast<cudaStream_t>(stream)>>>(p);
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}
