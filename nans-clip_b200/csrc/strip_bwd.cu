// strip_bwd.cu — kernel (3): fused contrastive backward (C entry point nans_clip_loss_bwd).
//
// Replaces the autograd backward of cn_clip/training/train.py:87-115.  For a block of local rows A
// (image or text features) and every tile B_j of columns of the gathered other modality:
//   MMA1 : S_j = A * B_j^T                      (recomputed logits / s, fp32 in TMEM)
//   warps: G_j = 2^12 * ( exp(s S - lse_row) + exp(s S - lse_col[j]) - 2 delta_label )  -> 16 bit.
//          One MUFU per element instead of two: exp(sS-lr) + exp(sS-lc) = exp(sS-lr) * (1 + a_i b_j),
//          a_i = 2^(lr_i - mu), b_j = 2^(mu - lc_j); b_j is produced once per tile by warp 3.
//          (Used when all lse lie within 2^+-50 of mu; otherwise the two-exp form runs.)
//   MMA2 : dA += G_j * B_j                      (B_j streamed a second time from L2 and read as an
//          MN-major operand: rows of B_j are the K dimension)
// Two strips per rank (dI: image rows x text columns, dT: text rows x image columns).
//
// Files: bwd_common.cuh (parameters, softmax-gradient loop), bwd_narrow.cuh (64 rows per CTA, CTA
// pairs with M = 128: the default for D <= 1024 — dA for all features of a pass stays in TMEM, so
// the logits are recomputed once per tile), bwd_wide.cuh (128 rows per CTA: the first two
// generations, kept for D > 1024 and as fallbacks), this file (lse min/max, output cast,
// scheduling plans, dispatch).
//
// Roofline: tensor cores.  Algorithmic flops = 4 * rows * N * D per rank (S recompute excluded).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace nans {
namespace {

#include "bwd_common.cuh"
#include "bwd_wide.cuh"
#include "bwd_narrow.cuh"

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

// ---- which kernel, which grid: everything that depends only on (gradient rows, N, D) --------------
enum BwdKind { BWD_WIDE_PAIR = 1, BWD_NARROW = 2, BWD_NARROW_PERSISTENT = 3 };  // 0 was the retired single-CTA kernel

struct BwdSchedule {
  int kind;
  int kchunks, npass;          // 64-feature chunks; passes over the logits (feature slices of dA)
  int tile_cols;               // column tile width
  bool a_resident;
  int nrb, ntiles, nr;         // row blocks per strip, column tiles, ring stages
  int nsplit;                  // uniform column splits (wide kernels); 2 = "outputs are accumulated" for persistent
  NpTail tail;                 // narrow: whole-wave units + column-split partial wave
  int npp_t1, npp_units, npairs;  // persistent narrow kernel
  size_t smem_bytes;
  unsigned grid;
  bool zero_all, zero_tail;    // which output rows must be zeroed before the launch
};

BwdSchedule plan_bwd_schedule(int64_t grad_row_count, int64_t N, int64_t D) {
  BwdSchedule sc{};
  const int kchunks = static_cast<int>(ceil_div(D, BK));
  const int npass = static_cast<int>(ceil_div(kchunks, SLICE / BK));
  const bool use_pair = true;  // every kernel is a CTA-pair kernel (cta_group::2)
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
  // persistent helper schedule of the narrow-pair kernel
  bool use_npp = false;
  int npp_t1 = 0, npp_units = 0;
  if (use_np) {
    // worth it when whole waves of one unit per CTA pair would leave SMs idle (64 units on 74 pairs at
    // n_loc = 4096); with many waves (N = 32768 on one GPU: 512 units, 98.8 % full) the plain-store
    // kernel wins: no output memset, no red.add, no segment hand-overs
    const int64_t units = 2 * ceil_div(grad_row_count, 2 * NP_ROWS);
    const int64_t slots = sm_count() / 2;
    const double fill = static_cast<double>(units) / static_cast<double>(ceil_div(units, slots) * slots);
    // (An "equal contiguous ranges" schedule was tried and retired: at 2 GPUs, n_loc = 16384, 2.61 ms/step
    // against 2.50 — pairs that stop walking the column tiles in lockstep lose their L2 hits.)
    // Helper mode (default when there are FEWER units than CTA pairs, e.g. 64 units on 74 pairs at
    // n_loc = 4096): every unit keeps its own pair for the first T1 tiles (lockstep preserved) and the
    // idle pairs share the last ntiles - T1 tiles of all units.  NANS_BWD_PERSIST=0 disables it.
    const char* e = getenv("NANS_BWD_PERSIST");
    const int64_t nt256 = ceil_div(N, NP_KT);
    const bool force_helpers = e && e[0] == '2';  // tests: helper mode wherever it is feasible
    if (!(e && e[0] == '0') && kchunks <= 8 && units < slots && nt256 >= 2 &&
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
  const PairPlan pplan = plan_pair(kchunks);
  const NpPlan nplan = plan_np(kchunks, np_tk, np_ares);
  const int64_t np_units = 2 * ceil_div(grad_row_count, 2 * NP_ROWS) * np_npass;
  const NpTail tail = plan_np_tail(np_units, ceil_div(N, np_tk), sm_count() / 2, np_tk == NP_KT ? 3.0 : 6.0);
  const int nsplit = use_npp    ? 2  /* outputs are always accumulated (zeroed below) */
                     : use_np   ? 1  /* per-unit: see plan_np_tail */
                                : choose_bwd_nsplit(grad_row_count, N, npass, 2 * BM, sm_count() / 2);
  sc.kchunks = kchunks;
  sc.npp_t1 = npp_t1;
  sc.npp_units = npp_units;
  sc.tail = tail;
  sc.nsplit = nsplit;
  sc.kind = use_npp ? BWD_NARROW_PERSISTENT : use_np ? BWD_NARROW : BWD_WIDE_PAIR;
  sc.npass = (use_np || use_npp) ? np_npass : npass;
  sc.tile_cols = use_np ? np_tk : KT;
  sc.a_resident = use_np ? np_ares : pplan.a_resident;
  sc.nrb = static_cast<int>(ceil_div(grad_row_count, use_np ? 2 * NP_ROWS : (use_pair ? 2 * BM : BM)));
  sc.ntiles = static_cast<int>(ceil_div(N, sc.tile_cols));
  sc.nr = use_np ? nplan.nr : pplan.nr;
  sc.smem_bytes = use_np ? nplan.bytes : pplan.bytes;
  sc.npairs = 0;
  if (use_npp) {
    sc.npairs = sm_count() / 2;  // npp_units main pairs + the helpers
    sc.grid = static_cast<unsigned>(2 * sc.npairs);
  } else if (use_np) {
    sc.grid = static_cast<unsigned>(2 * (tail.n_full + (np_units - tail.n_full) * tail.ns_tail));
  } else {
    sc.grid = static_cast<unsigned>((use_pair ? 2 : 1) * 2 * sc.nrb * sc.npass * sc.nsplit);
  }
  sc.zero_all = nsplit > 1 || (sc.kind == BWD_NARROW && tail.ns_tail > 1 && np_npass > 1);
  sc.zero_tail = !sc.zero_all && sc.kind == BWD_NARROW && tail.ns_tail > 1;
  return sc;
}

}  // namespace
}  // namespace nans

using namespace nans;

static int bwd_dispatch(const void* I_loc, const void* T_loc, int64_t ld_loc, const void* T_all, const void* I_all,
                        int64_t ld_all, const uint32_t* col_step, int feat_dtype, int64_t n_loc, int64_t N, int64_t D,
                        int64_t label_begin, const float* s_dev, const float* lse_img_all, const float* lse_txt_all,
                        const int* lse_minmax, const float* grad_out_dev, float grad_mult, int64_t grad_row_begin,
                        int64_t grad_row_count, void* dI_loc, void* dT_loc, int out_dtype, void* ws, size_t ws_bytes,
                        void* stream);

extern "C" size_t nans_clip_loss_bwd_workspace_bytes(int64_t grad_row_count, int64_t N, int64_t D) {
  (void)N;
  if (grad_row_count <= 0 || D <= 0) return 512;
  return 2 * align_up(static_cast<size_t>(grad_row_count) * D * 4, 256) + 512;
}

extern "C" int nans_clip_loss_bwd_plan(int64_t grad_row_count, int64_t N, int64_t D, int64_t* out, int n_out) {
  if (grad_row_count <= 0 || N <= 0 || D <= 0 || D % 8 != 0 || out == nullptr || n_out < NANS_BWD_PLAN_FIELDS) {
    set_error("loss_bwd_plan: bad arguments");
    return NANS_ERR_ARG;
  }
  const BwdSchedule sc = plan_bwd_schedule(grad_row_count, N, D);
  const int64_t v[NANS_BWD_PLAN_FIELDS] = {
      sc.kind, sc.tile_cols, sc.npass, sc.a_resident ? 1 : 0, sc.nrb, sc.ntiles, sc.nr,
      static_cast<int64_t>(sc.smem_bytes), static_cast<int64_t>(sc.grid), sc.nsplit, sc.tail.n_full,
      sc.tail.ns_tail, sc.npp_t1, sc.npp_units, sc.npairs, sc.zero_all ? 2 : (sc.zero_tail ? 1 : 0)};
  for (int i = 0; i < NANS_BWD_PLAN_FIELDS; ++i) out[i] = v[i];
  return NANS_OK;
}

extern "C" int nans_clip_loss_bwd(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                  const void* T_all, const void* I_all, int64_t ld_all,
                                  int feat_dtype, int64_t n_loc, int64_t N, int64_t D,
                                  int64_t label_begin, const float* s_dev, const float* lse_img_all,
                                  const float* lse_txt_all, const float* grad_out_dev,
                                  float grad_mult, int64_t grad_row_begin, int64_t grad_row_count,
                                  void* dI_loc, void* dT_loc, int out_dtype, void* ws,
                                  size_t ws_bytes, void* stream) {
  return nans_clip_loss_bwd_minmax(I_loc, T_loc, ld_loc, T_all, I_all, ld_all, feat_dtype, n_loc, N, D, label_begin,
                                   s_dev, lse_img_all, lse_txt_all, nullptr, grad_out_dev, grad_mult, grad_row_begin,
                                   grad_row_count, dI_loc, dT_loc, out_dtype, ws, ws_bytes, stream);
}

extern "C" int nans_clip_loss_bwd_minmax(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                         const void* T_all, const void* I_all, int64_t ld_all,
                                         int feat_dtype, int64_t n_loc, int64_t N, int64_t D,
                                         int64_t label_begin, const float* s_dev, const float* lse_img_all,
                                         const float* lse_txt_all, const int* lse_minmax,
                                         const float* grad_out_dev, float grad_mult, int64_t grad_row_begin,
                                         int64_t grad_row_count, void* dI_loc, void* dT_loc, int out_dtype,
                                         void* ws, size_t ws_bytes, void* stream) {
  return bwd_dispatch(I_loc, T_loc, ld_loc, T_all, I_all, ld_all, nullptr, feat_dtype, n_loc, N, D, label_begin, s_dev,
                      lse_img_all, lse_txt_all, lse_minmax, grad_out_dev, grad_mult, grad_row_begin, grad_row_count,
                      dI_loc, dT_loc, out_dtype, ws, ws_bytes, stream);
}

// Exchange mode: the column operands are this rank's gathered buffers ([2 slots * N, D] per modality);
// the slot is selected on the device from the step recorded by nans_clip_loss_exchange_finish_xchg.
extern "C" int nans_clip_loss_bwd_xchg(const nans_xchg_t* x, const uint32_t* step_dev, const void* I16_loc,
                                       const void* T16_loc, int feat_dtype, const float* s_dev,
                                       const float* lse_img_all, const float* lse_txt_all, const int* lse_minmax,
                                       const float* grad_out_dev, float grad_mult, int64_t grad_row_begin,
                                       int64_t grad_row_count, void* dI_loc, void* dT_loc, int out_dtype, void* ws,
                                       size_t ws_bytes, void* stream) {
  NANS_REQUIRE(x != nullptr && x->world >= 1 && x->world <= NANS_MAX_PEERS && x->rank >= 0 && x->rank < x->world &&
                   x->n_loc > 0 && x->D > 0 && x->base[x->rank] != nullptr,
               "loss_bwd_xchg: bad exchange descriptor");
  NANS_REQUIRE(step_dev != nullptr && lse_minmax != nullptr, "loss_bwd_xchg: step_dev and lse_minmax are required");
  const int64_t N = static_cast<int64_t>(x->world) * x->n_loc;
  const uint8_t* own = static_cast<const uint8_t*>(x->base[x->rank]);
  const size_t mod_bytes = static_cast<size_t>(2) * N * x->D * 2;
  return bwd_dispatch(I16_loc, T16_loc, x->D, own + x->feat_off + mod_bytes, own + x->feat_off, x->D, step_dev,
                      feat_dtype, x->n_loc, N, x->D, static_cast<int64_t>(x->rank) * x->n_loc, s_dev, lse_img_all,
                      lse_txt_all, lse_minmax, grad_out_dev, grad_mult, grad_row_begin, grad_row_count, dI_loc, dT_loc,
                      out_dtype, ws, ws_bytes, stream);
}

// col_step != nullptr: T_all / I_all hold 2 * N rows (two slots), see BwdParams::col_step
static int bwd_dispatch(const void* I_loc, const void* T_loc, int64_t ld_loc, const void* T_all, const void* I_all,
                        int64_t ld_all, const uint32_t* col_step, int feat_dtype, int64_t n_loc, int64_t N, int64_t D,
                        int64_t label_begin, const float* s_dev, const float* lse_img_all, const float* lse_txt_all,
                        const int* lse_minmax, const float* grad_out_dev, float grad_mult, int64_t grad_row_begin,
                        int64_t grad_row_count, void* dI_loc, void* dT_loc, int out_dtype, void* ws, size_t ws_bytes,
                        void* stream) {
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

  const BwdSchedule sc = plan_bwd_schedule(grad_row_count, N, D);
  NANS_REQUIRE(col_step == nullptr || sc.kind == BWD_NARROW || sc.kind == BWD_NARROW_PERSISTENT,
               "loss_bwd_xchg: the exchange mode needs the narrow-pair backward (D <= 1024, no NANS_BWD_* overrides)");
  const int64_t tm_rows = col_step != nullptr ? 2 * N : N;  // rows of the column operands' tensor maps
  const int kchunks = sc.kchunks;
  const size_t out_bytes = static_cast<size_t>(grad_row_count) * D * 4;
  // workspace: [0,256) lse min/max slots, then (16-bit outputs only) the two fp32 gradient buffers
  if (ws == nullptr || ws_bytes < 512) {
    set_error("loss_bwd: workspace %zu < 512 bytes", ws_bytes);
    return NANS_ERR_WORKSPACE;
  }
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "loss_bwd: workspace must be 16-byte aligned");
  const int* minmax = lse_minmax;
  if (minmax == nullptr) {  // not supplied by nans_clip_loss_exchange_finish: one small launch
    int* mm = static_cast<int*>(ws);
    lse_minmax_kernel<<<1, 1024, 0, st>>>(lse_img_all, lse_txt_all, static_cast<int>(N), mm);
    NANS_CUDA_OK(cudaGetLastError());
    minmax = mm;
  }
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
  if (sc.zero_all) {
    NANS_CUDA_OK(cudaMemsetAsync(out32[0], 0, out_bytes, st));
    NANS_CUDA_OK(cudaMemsetAsync(out32[1], 0, out_bytes, st));
  } else if (sc.zero_tail) {
    // only the rows of the column-split tail units are accumulated (single pass: unit = (strip, row block))
    const int64_t nrb_np = sc.nrb;
    for (int strip = 0; strip < 2; ++strip) {
      const int64_t u0 = std::max<int64_t>(sc.tail.n_full, strip * nrb_np), u1 = (strip + 1) * nrb_np;
      if (u0 >= u1) continue;
      const int64_t r0 = (u0 - strip * nrb_np) * 2 * NP_ROWS;
      const int64_t r1 = std::min<int64_t>(grad_row_count, (u1 - strip * nrb_np) * 2 * NP_ROWS);
      NANS_CUDA_OK(cudaMemsetAsync(out32[strip] + r0 * D, 0, static_cast<size_t>(r1 - r0) * D * 4, st));
    }
  }

  CUtensorMap tmA0, tmB0, tmA1, tmB1, tmBk0, tmBk1;
  if ((rc = make_tmap_16b(&tmA0, I_loc, feat_dtype, n_loc, D, ld_loc, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB0, T_all, feat_dtype, tm_rows, D, ld_all, KT)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmA1, T_loc, feat_dtype, n_loc, D, ld_loc, BM)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmB1, I_all, feat_dtype, tm_rows, D, ld_all, KT)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmBk0, T_all, feat_dtype, tm_rows, D, ld_all, KT / 2)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmBk1, I_all, feat_dtype, tm_rows, D, ld_all, KT / 2)) != NANS_OK) return rc;
  CUtensorMap tmAn0, tmAn1;  // narrow pairs: 64-row A boxes
  if ((rc = make_tmap_16b(&tmAn0, I_loc, feat_dtype, n_loc, D, ld_loc, NP_ROWS)) != NANS_OK) return rc;
  if ((rc = make_tmap_16b(&tmAn1, T_loc, feat_dtype, n_loc, D, ld_loc, NP_ROWS)) != NANS_OK) return rc;

  BwdParams p;
  p.row_begin = static_cast<int>(grad_row_begin);
  p.row_end = static_cast<int>(grad_row_begin + grad_row_count);
  p.ncols = static_cast<int>(N);
  p.D = static_cast<int>(D);
  p.kchunks = kchunks;
  p.nrb = sc.nrb;
  p.npass = sc.npass;
  p.nsplit = sc.nsplit;
  p.ntiles = sc.ntiles;
  p.nr = sc.nr;
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
  p.accumulate = sc.nsplit > 1 ? 1 : 0;
  p.lse_minmax = minmax;
  p.col_step = col_step;
  {
    const char* e = getenv("NANS_BWD_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }

  const size_t smem = sc.smem_bytes;
  p.n_full = sc.tail.n_full;
  p.ns_tail = sc.tail.ns_tail;
  p.total_tiles = 2ll * sc.nrb * sc.ntiles;
  p.npp_t1 = sc.npp_t1;
  p.npp_units = sc.npp_units;
  p.npairs = sc.npairs;
  if (sc.kind == BWD_NARROW_PERSISTENT) {
    NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_npp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    clip_bwd_npp_kernel<<<sc.grid, NUM_THREADS, smem, st>>>(tmAn0, tmB0, tmAn1, tmB1, p);
  } else if (sc.kind == BWD_NARROW) {
    if (sc.tile_cols == NP_KT && !sc.a_resident) {
      NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_np_kernel<NP_KT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
      clip_bwd_np_kernel<NP_KT, false><<<sc.grid, NUM_THREADS, smem, st>>>(tmAn0, tmB0, tmB0, tmAn1, tmB1, tmB1, p);
    } else if (sc.tile_cols == NP_KT) {
      NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_np_kernel<NP_KT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
      clip_bwd_np_kernel<NP_KT, true><<<sc.grid, NUM_THREADS, smem, st>>>(tmAn0, tmB0, tmB0, tmAn1, tmB1, tmB1, p);
    } else {
      NANS_CUDA_OK(cudaFuncSetAttribute(clip_bwd_np_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
      clip_bwd_np_kernel<128, true><<<sc.grid, NUM_THREADS, smem, st>>>(tmAn0, tmBk0, tmB0, tmAn1, tmBk1, tmB1, p);
    }
  } else {
    auto kern = sc.a_resident ? clip_bwd_pair_kernel<true> : clip_bwd_pair_kernel<false>;
    NANS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<sc.grid, NUM_THREADS, smem, st>>>(tmA0, tmBk0, tmB0, tmA1, tmBk1, tmB1, p);
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
