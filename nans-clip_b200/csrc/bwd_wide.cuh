// bwd_wide.cuh — the 128-rows-per-CTA backward kernel `clip_bwd_pair_kernel` (CTA pairs, M = 256): the
// kernel for D > 1024 (NANS_BWD_NP=0 forces it for narrower features).  Its single-CTA ancestor was
// retired in round 2 (the pair form is faster everywhere, see the traffic argument below).
// G is written into TMEM over S and MMA2 is a TS MMA; dA is produced in 256-feature passes.
// Included by strip_bwd.cu only.
#pragma once
//
// Design notes of the 128-row kernels.  For a block of 128 local
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
