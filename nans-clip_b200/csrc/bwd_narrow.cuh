// bwd_narrow.cuh — the 64-rows-per-CTA backward kernels (default for D <= 1024):
// `clip_bwd_np_kernel<TK, ARES>` (one unit per CTA pair) and `clip_bwd_npp_kernel` (persistent
// schedules).  Included by strip_bwd.cu only.
#pragma once
// ================================================================================================
// Narrow CTA pairs: the pair owns 128 rows, 64 per CTA (cta_group::2, M = 128).  With 64 rows per CTA
// the fp32 dA of D features takes D / 2 TMEM columns (the M = 128 pair layout puts the two N halves
// of an accumulator on lanes 0-63 / 64-127), which leaves room for two S buffers:
//     D <=  512 : dA 256 columns, S 2 x 128 columns -> 256-column tiles  (TK = 256, A resident)
//     D <=  768 : dA 384 columns, S 2 x  64 columns -> 128-column tiles  (TK = 128, A resident)
//     D <= 1024 : two passes of 512 features (dA 256 columns each), TK = 256, A streamed (ARES = false)
// so the logits are recomputed ONCE per tile (and pass) instead of once per 256-feature slice.
// The kernel is bound by the shared-memory data pipe (tensor-core operand reads + the softmax warps'
// shared traffic; ncu: tc 60 % + LSU 37 %), hence the widest MMA1 N that fits (A is re-read once per
// tile), ld/st.shared, one lane-distributed load + shuffles for the column factors, parked waits.
//   TMEM : dA blocks of 256 features (128 columns each) from column 0 | S buffers at the top.
//          S / dA rows 0-63 of the CTA sit on lanes 0-63 (first N half) and 64-127 (second N half).
//   SMEM : [A (64 rows, resident)] | 2 G buffers (64 rows x TK K, 16-bit, K-major swizzled: G goes
//          through shared memory because the TMEM-A form of a pair MMA wants a duplicated layout)
//          | ring of 32 KB (48 KB when A is streamed) stages: MMA1 = chunks of [TK/2 tile rows x 64
//          features] of this CTA's half of the tile [+ the matching A chunk]; MMA2 = [128 tile rows
//          x 64 features] x 2 of this CTA's 128 features of a 256-feature block.
//          MMA2 of tile t is issued after MMA1 of tile t + 1.
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
    const int coff = col_slot_offset(p);  // exchange mode: this step's slot of the gathered tensors
    for (int tau = 0; tau <= ntiles; ++tau) {
      if (tau < ntiles) {
        const int col0 = (tile_begin + tau) * TK + static_cast<int>(rank) * (TK / 2) + coff;
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
        const int col0 = (tile_begin + tau - 1) * TK + coff;
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
// Persistent, load-balanced form of the narrow-pair kernel (helper schedule; see the dispatch for when it
// is used).  With one unit per CTA pair, 64 units on 74 pairs leave 14 % of the machine idle at
// n_loc = 4096: every unit keeps its own pair for its first npp_t1 tiles, and the idle pairs share the
// remaining tiles of ALL units, linearised (unit major, tiles inside) and cut into equal contiguous ranges.
// A helper's range crosses unit boundaries: each piece is a SEGMENT with its own A block, its own dA
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
  if (pair < p.npp_units) {  // main pair: the first npp_t1 tiles of its own unit
    tile_lo = 0; tile_hi = p.npp_t1;
    g0 = pair * p.npp_t1; nt = p.npp_t1;
  } else {                   // helper pair: an equal share of the last tiles of ALL units
    tile_lo = p.npp_t1; tile_hi = p.ntiles;
    const long long tot = static_cast<long long>(p.npp_units) * (p.ntiles - p.npp_t1);
    const long long h = pair - p.npp_units, nh = p.npairs - p.npp_units;
    g0 = h * tot / nh; nt = (h + 1) * tot / nh - g0;
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
    const int coff = col_slot_offset(p);  // exchange mode: this step's slot of the gathered tensors
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
        const int col0 = cur.tile * NP_KT + static_cast<int>(rank) * (NP_KT / 2) + coff;
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
        const int col0 = prv.tile * NP_KT + coff;
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
