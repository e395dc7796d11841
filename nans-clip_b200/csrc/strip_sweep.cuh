// strip_sweep.cuh — the shared contraction skeleton of kernels (2) and (4).
//
// One CTA owns a block of BM = 128 rows of the row operand A (image or text features of the local
// batch, or a block of queries) and sweeps a range of BN = 256-column tiles of the column operand B
// (gathered features / a gallery shard).  For every tile S = A * B^T (fp32, K = D) is formed in
// TMEM by tcgen05.mma and handed to an epilogue functor; S never goes to shared or global memory.
//
//   warp 0    : TMA producer   (A block once if it fits in smem, else per K-chunk; B per K-chunk)
//   warp 1    : tcgen05.mma issuer (one lane)
//   warp 2    : TMEM allocator (512 columns = two 128x256 fp32 accumulators, double-buffered)
//   warps 4-11: epilogue in two sets of 4 warps that alternate tiles; thread = row (TMEM lane),
//               tcgen05.ld 32 columns at a time
//
// Shared memory (128-byte swizzle, K-major, 64 elements per line):
//   A resident: kchunks x 16 KB  +  stages x 32 KB (B)      when that fits (D <= 640)
//   A streamed: stages x (16 KB + 32 KB)
#pragma once
#include "common.cuh"

namespace nans {
namespace sweep {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int A_CHUNK = BM * BK * 2;  // 16 KB
constexpr int B_STAGE = BN * BK * 2;  // 32 KB
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int MAX_STAGES = 6;
constexpr int TMEM_COLS = 512;
constexpr int BAR_BYTES = 256;
constexpr size_t SMEM_CAP = 227 * 1024;

struct SmemPlan {
  bool a_resident;
  int stages;
  size_t bytes;  // dynamic shared memory to request (includes 1 KB alignment slack)
};

// pair = true: CTA-pair mode, a CTA stages half of every B stage (16 KB)
inline SmemPlan plan_smem(int kchunks, bool pair) {
  SmemPlan p;
  const size_t bst = pair ? B_STAGE / 2 : B_STAGE;
  const size_t cap = SMEM_CAP - 1024 - BAR_BYTES;
  const size_t a_res = static_cast<size_t>(kchunks) * A_CHUNK;
  // A stays resident only if at least 4 ring stages are left (D = 768 in pair mode would fit with 2:
  // measured 1253 TFLOP/s against 1620 when A is streamed through a 7-stage ring instead)
  if (a_res + 4 * bst <= cap) {
    p.a_resident = true;
    p.stages = static_cast<int>((cap - a_res) / bst);
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    p.bytes = a_res + static_cast<size_t>(p.stages) * bst + BAR_BYTES + 1024;
  } else {
    p.a_resident = false;
    p.stages = static_cast<int>(cap / (A_CHUNK + bst));
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    p.bytes = static_cast<size_t>(p.stages) * (A_CHUNK + bst) + BAR_BYTES + 1024;
  }
  return p;
}

#ifdef __CUDACC__

struct SweepArgs {
  const CUtensorMap* tmA;
  const CUtensorMap* tmB;
  int row0;        // first row of this CTA's block in A
  int tile_begin;  // column tiles [tile_begin, tile_end) of B, counted WITHOUT the skipped tiles
  int tile_end;
  int skip_begin;  // tiles [skip_begin, skip_begin + skip_count) of B are not swept (another phase
  int skip_count;  // covered them): swept index t maps to tile t + (t >= skip_begin ? skip_count : 0)
  int kchunks;     // ceil(D / 64)
  int stages;
  uint32_t idesc;  // M=128 (256 in pair mode), N=256, K-major A and B, fp32 accumulate
  // Exchange mode (xw > 0; multi-GPU push data plane, exchange.cu): B is the gathered buffer of an
  // xw-rank job, xsrc_tiles tiles per source rank.  The unit sweeps the tiles
  // [xsub_begin, xsub_begin + xsub_count) of EVERY source, sources in the order xrank, xrank + 1, ...
  // (its own rows first: they are there already; the others in the order the peers push them), so
  // tile_begin = 0 and tile_end = xw * xsub_count.  Before a tile's loads the producer waits for the
  // arrival flags (one per 64 rows, value >= xstep) of the rows it is about to read.
  int xw = 0, xrank = 0, xsrc_tiles = 1, xsub_begin = 0, xsub_count = 1;
  int xrow_off = 0;                  // row of column 0 in the gathered tensor map (slot of this step)
  const uint32_t* xflags = nullptr;  // [xw][xsrc_tiles * 4] arrival flags of the column operand's modality
  uint32_t xstep = 0;
  long long* xwait_ns = nullptr;     // diagnostics: per CTA, nanoseconds the producer spent waiting for flags
};

// swept index t of a unit -> tile of the column operand (in logical, global column tiles)
__device__ __forceinline__ int tile_of(const SweepArgs& a, int t) {
  if (a.xw > 0) {
    const int k = t / a.xsub_count;
    int src = a.xrank + k;
    if (src >= a.xw) src -= a.xw;
    return src * a.xsrc_tiles + a.xsub_begin + (t - k * a.xsub_count);
  }
  const int tix = a.tile_begin + t;
  return tix + (tix >= a.skip_begin ? a.skip_count : 0);
}
// inverse: the swept index at which this unit visits `tile`, or -1 if it does not
__device__ __forceinline__ int swept_index_of(const SweepArgs& a, int tile) {
  if (tile < 0) return -1;
  if (a.xw > 0) {
    const int src = tile / a.xsrc_tiles;
    const int j = tile - src * a.xsrc_tiles - a.xsub_begin;
    if (src >= a.xw || j < 0 || j >= a.xsub_count) return -1;
    int k = src - a.xrank;
    if (k < 0) k += a.xw;
    return k * a.xsub_count + j;
  }
  if (tile >= a.skip_begin && tile < a.skip_begin + a.skip_count) return -1;
  const int sw = tile < a.skip_begin ? tile : tile - a.skip_count;
  return (sw >= a.tile_begin && sw < a.tile_end) ? sw - a.tile_begin : -1;
}

// Epi must provide:  __device__ void tile(uint32_t taddr, int tile_idx)
//   taddr = TMEM address of this thread's warp lane group at column 0 of the accumulator buffer.
// A row is served by two threads (one per epilogue set): set p sees the tiles tile_begin + p,
// tile_begin + p + 2, ...; the caller merges the two partial results.
//
// CP = true: CTA-pair mode (cta_group::2).  The two CTAs of a cluster own 256 consecutive rows
// (SweepArgs::row0 is the CTA's own first row) and share every tcgen05.mma (M = 256, N = 256): each
// CTA stages only its half of every B stage (128 of the 256 tile rows), which halves the shared-
// memory traffic per MMA — the single-CTA kernel moves 4 KB (A) + 8 KB (B) of reads and 8 KB of TMA
// writes per 128-clk MMA against 128 B/clk of shared memory (84 % tensor pipe measured).  Only the
// leader (cluster rank 0) issues MMAs; loads of both CTAs credit the leader's barriers, the leader's
// commits are multicast, and the partner's epilogue warps release accumulators remotely.
//
// Returns the 1 KB-aligned base of the dynamic shared memory, free for reuse after the call
// (all TMA loads and MMAs of this CTA have completed and every thread has passed a barrier).
template <bool A_RES, bool CP, class Epi>
__device__ __forceinline__ uint8_t* run(const SweepArgs& a, Epi& epi) {
  constexpr int BST = CP ? B_STAGE / 2 : B_STAGE;  // this CTA's bytes of a B stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smem + static_cast<size_t>(A_RES ? a.kchunks : a.stages) * A_CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + static_cast<size_t>(a.stages) * BST);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* a_full = bars + 2 * MAX_STAGES;
  uint64_t* tfull = a_full + 1;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CP ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr uint32_t NCTA = CP ? 2u : 1u;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(a.tmA);
      tma_prefetch_desc(a.tmB);
      for (int i = 0; i < a.stages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      mbar_init(a_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull[i], 1);
        mbar_init(&tempty[i], NCTA * (EPI_WARPS / 2));
      }
      fence_barrier_init();
    }
  } else if (warp == 2) {
    if (CP) {
      tmem_alloc_pair(tmem_ptr, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CP) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int ntiles = a.tile_end - a.tile_begin;

  if (warp == 0) {
    // ---------------- TMA producer (whole warp in the loop, one elected lane issues) ----------
    if (A_RES) {
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(a_full, NCTA * static_cast<uint32_t>(a.kchunks) * A_CHUNK);
        for (int c = 0; c < a.kchunks; ++c) {
          if (CP) tma_load_2d_pair(smA + static_cast<size_t>(c) * A_CHUNK, a.tmA, a_full, c * BK, a.row0);
          else tma_load_2d(smA + static_cast<size_t>(c) * A_CHUNK, a.tmA, a_full, c * BK, a.row0);
        }
      }
      __syncwarp();
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int tile = tile_of(a, t);
      const int col0 = tile * BN + (CP ? static_cast<int>(rank) * (BN / 2) : 0) + a.xrow_off;
      if (a.xw > 0 && tile / a.xsrc_tiles != a.xrank) {
        // remote rows [col0, col0 + rows of this CTA's part of the tile): one flag per 64 rows
        constexpr int NFLAG = (CP ? BN / 2 : BN) / 64;
        long long t0 = 0;
        if (a.xwait_ns != nullptr) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        if (lane < NFLAG) wait_flag_sys(a.xflags + (col0 - a.xrow_off) / 64 + lane, a.xstep);
        __syncwarp();
        fence_proxy_async_global();  // the peer's generic-proxy writes -> this CTA's TMA (async proxy) reads
        if (a.xwait_ns != nullptr && lane == 0) {
          long long t1;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
          a.xwait_ns[blockIdx.x] += t1 - t0;
        }
      }
      for (int c = 0; c < a.kchunks; ++c) {
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&full[stage], NCTA * (A_RES ? BST : (A_CHUNK + BST)));
          if (CP) {
            if (!A_RES)
              tma_load_2d_pair(smA + static_cast<size_t>(stage) * A_CHUNK, a.tmA, &full[stage], c * BK, a.row0);
            tma_load_2d_pair(smB + static_cast<size_t>(stage) * BST, a.tmB, &full[stage], c * BK, col0);
          } else {
            if (!A_RES)
              tma_load_2d(smA + static_cast<size_t>(stage) * A_CHUNK, a.tmA, &full[stage], c * BK, a.row0);
            tma_load_2d(smB + static_cast<size_t>(stage) * BST, a.tmB, &full[stage], c * BK, col0);
          }
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && leader) {
    // ---------------- MMA issuer (leader CTA only in pair mode) ----------------
    if (A_RES) {
      mbar_wait(a_full, 0);
      tc_fence_after();
    }
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t smA_addr = smem_u32(smA);
    const uint32_t smB_addr = smem_u32(smB);
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int c = 0; c < a.kchunks; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smA_addr + static_cast<uint32_t>(A_RES ? c : stage) * A_CHUNK;
          const uint32_t b_addr = smB_addr + static_cast<uint32_t>(stage) * BST;
          const uint64_t ad = make_smem_desc(a_addr, 16, 1024);
          const uint64_t bd = make_smem_desc(b_addr, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {  // +32 bytes (>> 4 = 2) per K step inside the swizzle atom
            if (CP) mma_ss_pair(d_tmem, ad + 2 * k, bd + 2 * k, a.idesc, (c | k) != 0 ? 1u : 0u);
            else mma_ss(d_tmem, ad + 2 * k, bd + 2 * k, a.idesc, (c | k) != 0 ? 1u : 0u);
          }
          if (CP) {
            tc_commit_pair(&empty[stage], 3);
            if (c == a.kchunks - 1) tc_commit_pair(&tfull[acc], 3);
          } else {
            tc_commit(&empty[stage]);  // frees the smem stage once these MMAs have read it
            if (c == a.kchunks - 1) tc_commit(&tfull[acc]);  // accumulator complete -> epilogue
          }
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: 2 sets of 4 warps; set p owns the tiles t with (t & 1) == p ------
    // (= TMEM accumulator buffer p), so each set has two MMA tile times per tile and the other
    // set's tile is drained concurrently.  lane group = warp % 4.
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int set = (warp - 4) >> 2;
    uint32_t acc_phase = 0;
    for (int t = set; t < ntiles; t += 2) {
      mbar_wait(&tfull[set], acc_phase);
      tc_fence_after();
      epi.tile(tmem_base + lane_base + static_cast<uint32_t>(set * BN), tile_of(a, t));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CP) mbar_arrive_cluster(&tempty[set], 0);
        else mbar_arrive(&tempty[set]);
      }
      acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CP) cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if (CP) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  return smem;
}

#endif  // __CUDACC__

}  // namespace sweep
}  // namespace nans
