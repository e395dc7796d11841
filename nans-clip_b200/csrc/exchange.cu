// exchange.cu — the multi-GPU data plane of the loss path: peer-mapped exchange buffers (CUDA IPC) and
// kernel (1) fused with the feature exchange.
//
// Replaces the two feature all-gathers of cn_clip/training/train.py:53-84 (torch.distributed
// all_gather of image_features / text_features in every get_loss call).  Instead of a collective that
// sits between the cast and the forward, the cast kernel itself stores every 16-bit row into the
// gathered buffers of all ranks (its own through local memory, the peers' through NVLink: plain
// 16-byte st.global on peer-mapped addresses) and raises one arrival flag per 64 rows; the forward
// (strip_fwd.cu, exchange mode of strip_sweep.cuh) polls a tile's flags right before its TMA loads.
//
// Order of the pushes: source r serves destination r-1 first, then r-2, ... while destination d
// consumes its sources in the order d, d+1, d+2, ...: at any moment every destination is written by one
// source and every source writes one destination, and tiles arrive in the order they are consumed.
//
// Roofline: NVLink.  Bytes that must cross the link per rank and step: 2 modalities x (world - 1) x
// n_loc x D x 2 B (56 MiB at world 8, n_loc 4096, D 512: 76 us at the measured 770 GB/s per direction),
// hidden behind the forward, which needs 2 x that time for the same columns.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace nans {
namespace {

constexpr int PUSH_ROWS = NANS_XCHG_FLAG_ROWS;  // rows per CTA = rows per arrival flag
constexpr int PUSH_THREADS = 256;

struct PushParams {
  const void* src[2];   // image, text rows of this rank
  void* loc16[2];       // local 16-bit copies [n_loc, D]
  int x_dtype, feat_dtype, normalize;
  int n_loc, D;
  long long ld_x;
  int world, rank;
  long long feat_off, fflag_off;
  long long slot_rows;  // world * n_loc
  const uint32_t* epoch;
  int k_begin, k_end;   // REMOTE variant: peers rank - k for k in [k_begin, k_end)
  uint32_t* stepvals;   // LOCAL variant: [2][n_loc / 64] words = this step's number (source of the DMA'd flags)
  int expect_slot;      // LOCAL variant: the slot the host issued this step's DMA copies for, or -1
  uint8_t* base[NANS_MAX_PEERS];
};

__device__ __forceinline__ void load8(const void* row, int x_dtype, int e0, float (&f)[8]) {
  if (x_dtype == NANS_F32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(row) + e0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(row) + e0 + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(row) + e0));
    if (x_dtype == NANS_F16) {
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
      }
    } else {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
      }
    }
  }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8], int feat_dtype) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (feat_dtype == NANS_BF16) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// grid = (blocks, 2 modalities); a CTA owns the 64 rows of one arrival flag at a time: 8 warps x 8 rows,
// a warp moves a row in 16-byte pieces (lane = piece).  Both variants read the SOURCE rows and cast
// (optionally normalise) them, so neither depends on the other:
//
// PUSH_LOCAL : into the local 16-bit copy and this rank's own gathered slot, then the own flags.
//              HBM-bound, a few microseconds; the forward of this rank is launched behind it on the same
//              stream.
// PUSH_REMOTE: into the gathered slot of every PEER, destinations in the order rank - 1, rank - 2, ...;
//              the flag of a destination goes up as soon as this CTA's 64 rows are there.  NVLink-bound;
//              launched on a side stream that forks BEFORE the local cast, so that its whole (capped) grid
//              is resident while only the small local-cast kernel runs, and it then runs UNDER the forward,
//              which polls the peers' flags tile by tile.  It waits for nothing.  (Launched behind the
//              local cast instead, its CTAs competed with the forward's one-CTA-per-SM grid for residency,
//              the forward of every rank spun on flags its peers' starved push kernels never raised, and
//              the bounded waits trapped.)
template <bool REMOTE>
__global__ void __launch_bounds__(PUSH_THREADS, 4) cast_push_kernel(const PushParams p) {
  const int mod = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t step = *p.epoch + 1u;
  const long long slot = step & 1u;
  const int D = p.D;
  const int esz = p.x_dtype == NANS_F32 ? 4 : 2;
  // byte offset of this modality's slot; global row (rank * n_loc + r) inside it
  const long long mod_off = p.feat_off + (static_cast<long long>(mod) * 2 + slot) * p.slot_rows * D * 2;
  const int nblk = p.n_loc / PUSH_ROWS;

  constexpr int RPW = PUSH_ROWS / (PUSH_THREADS / 32);  // rows per warp: 8, rows warp, warp + 8, ...
  constexpr int RIF = 4;                                // rows in flight per warp (32 registers of payload)
  if (!REMOTE) {
    // The host addresses the DMA copies of this step (nans_xchg_push_dma) by ITS count of forwards; the
    // kernels address the slot by the device's.  They differ when a CUDA graph with an odd number of steps
    // is replayed: fail loudly instead of reading the wrong slot.
    if (p.expect_slot >= 0 && static_cast<int>(slot) != p.expect_slot) __trap();
    if (p.stepvals != nullptr)
      for (int b = blockIdx.x * PUSH_THREADS + threadIdx.x; b < nblk; b += gridDim.x * PUSH_THREADS)
        p.stepvals[mod * nblk + b] = step;
  }
  for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int r0 = blk * PUSH_ROWS + warp;
    const uint8_t* srow0 = static_cast<const uint8_t*>(p.src[mod]) + static_cast<long long>(r0) * p.ld_x * esz;
    const long long sstep = static_cast<long long>(PUSH_THREADS / 32) * p.ld_x * esz;  // bytes between a warp's rows
    float inv[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) inv[i] = 1.0f;
    if (p.normalize) {
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        float ss = 0.f;
        for (int e0 = lane * 8; e0 < D; e0 += 256) {
          float f[8];
          load8(srow0 + i * sstep, p.x_dtype, e0, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
        }
        inv[i] = 1.0f / sqrtf(warp_sum(ss));  // same arithmetic as l2norm_cast_kernel
      }
    }
    for (int k = REMOTE ? p.k_begin : 0; k < (REMOTE ? p.k_end : 1); ++k) {
      int dst = p.rank - k;
      if (dst < 0) dst += p.world;
      uint8_t* drow0 = p.base[dst] + mod_off + (static_cast<long long>(p.rank) * p.n_loc + r0) * D * 2;
      uint8_t* lrow0 = REMOTE ? nullptr : static_cast<uint8_t*>(p.loc16[mod]) + static_cast<long long>(r0) * D * 2;
      const long long dstep = static_cast<long long>(PUSH_THREADS / 32) * D * 2;
#pragma unroll
      for (int ib = 0; ib < RPW; ib += RIF) {
        for (int e0 = lane * 8; e0 < D; e0 += 256) {
          // RIF of the warp's rows in flight at once (the loop is latency-bound otherwise); the source
          // rows are re-read per destination: after the first pass they sit in L1 / L2
          float f[RIF][8];
#pragma unroll
          for (int i = 0; i < RIF; ++i) load8(srow0 + (ib + i) * sstep, p.x_dtype, e0, f[i]);
#pragma unroll
          for (int i = 0; i < RIF; ++i) {
            if (p.normalize) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[i][e] *= inv[ib + i];
            }
            const uint4 v = pack8(f[i], p.feat_dtype);
            *reinterpret_cast<uint4*>(drow0 + (ib + i) * dstep + e0 * 2) = v;
            if (!REMOTE) *reinterpret_cast<uint4*>(lrow0 + (ib + i) * dstep + e0 * 2) = v;
          }
        }
      }
      if (REMOTE) {
        // This CTA's 64 rows are on their way to `dst`.  The barrier orders every thread's stores before
        // thread 0, whose release store at system scope is cumulative over them: one fence per (CTA,
        // destination) instead of one per thread (256 membar.sys per block made the kernel 4x slower).
        __syncthreads();
        if (threadIdx.x == 0) {
          uint32_t* flag = reinterpret_cast<uint32_t*>(p.base[dst] + p.fflag_off) +
                           (static_cast<long long>(mod) * p.world + p.rank) * nblk + blk;
          st_release_sys(flag, step);
        }
      }
      // LOCAL: no flag — the forward of this rank is launched behind this kernel on the same stream and
      // never polls its own rank's tiles
    }
  }
}

// CTAs per modality.  Both kernels of a step are in flight together and must fit the machine at once
// (4 CTAs of 256 threads x <= 64 registers per SM) with room to spare, so that every CTA of the remote push
// is resident before the forward is launched: local 1 per SM per modality, remote 1 per 2 SMs per modality
// (148 CTAs x 256 threads x 4 rows x 16 bytes in flight are far more than NVLink needs).
unsigned push_grid_x(int64_t n_loc, bool remote) {
  const int64_t blocks = n_loc / PUSH_ROWS;
  int64_t cap = remote ? sm_count() / 2 : sm_count();
  if (remote) {
    if (const char* e = getenv("NANS_PUSH_CTAS")) cap = atoi(e) > 0 ? atoi(e) : cap;  // bring-up: CTAs per modality
  }
  return static_cast<unsigned>(blocks < cap ? blocks : cap);
}

int fill_push_params(PushParams& p, const nans_xchg_t* x) {
  p.n_loc = static_cast<int>(x->n_loc);
  p.D = static_cast<int>(x->D);
  p.world = x->world;
  p.rank = x->rank;
  p.feat_off = x->feat_off;
  p.fflag_off = x->fflag_off;
  p.slot_rows = static_cast<long long>(x->world) * x->n_loc;
  p.epoch = x->epoch;
  for (int r = 0; r < NANS_MAX_PEERS; ++r) p.base[r] = r < x->world ? static_cast<uint8_t*>(x->base[r]) : nullptr;
  return NANS_OK;
}

int check_xchg(const nans_xchg_t* x, const char* who) {
  NANS_REQUIRE(x != nullptr && x->world >= 1 && x->world <= NANS_MAX_PEERS && x->rank >= 0 && x->rank < x->world,
               "%s: bad exchange descriptor", who);
  NANS_REQUIRE(x->n_loc > 0 && x->n_loc % 256 == 0 && x->D > 0 && x->D % 8 == 0, "%s: bad sizes", who);
  NANS_REQUIRE(x->epoch != nullptr, "%s: null step counter", who);
  for (int r = 0; r < x->world; ++r) NANS_REQUIRE(x->base[r] != nullptr, "%s: peer %d is not mapped", who, r);
  return NANS_OK;
}

}  // namespace
}  // namespace nans

using namespace nans;

extern "C" int nans_peer_alloc(size_t bytes, void** ptr, void* handle64) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(bytes > 0 && ptr && handle64, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  void* p = nullptr;
  NANS_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("peer_alloc: %s", cudaGetErrorString(e));
    return NANS_ERR_CUDA;
  }
  memcpy(handle64, &h, sizeof(h));
  *ptr = p;
  return NANS_OK;
}

extern "C" int nans_peer_free(void* ptr) {
  if (ptr) NANS_CUDA_OK(cudaFree(ptr));
  return NANS_OK;
}

extern "C" int nans_peer_open(const void* handle64, void** ptr) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  NANS_REQUIRE(handle64 && ptr, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  NANS_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return NANS_OK;
}

extern "C" int nans_peer_close(void* ptr) {
  if (ptr) NANS_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return NANS_OK;
}

extern "C" int nans_peer_zero(void* ptr, size_t bytes, void* stream) {
  NANS_REQUIRE(ptr != nullptr, "peer_zero: null pointer");
  NANS_CUDA_OK(cudaMemsetAsync(ptr, 0, bytes, static_cast<cudaStream_t>(stream)));
  return NANS_OK;
}

extern "C" int nans_xchg_layout(nans_xchg_t* x, int64_t n_loc, int64_t D) {
  NANS_REQUIRE(x != nullptr && x->world >= 1 && x->world <= NANS_MAX_PEERS, "xchg_layout: world must be in [1, %d]",
               NANS_MAX_PEERS);
  NANS_REQUIRE(n_loc > 0 && n_loc % 256 == 0 && D > 0 && D % 8 == 0 && D <= 8192,
               "xchg_layout: n_loc must be a positive multiple of 256 and D a multiple of 8 (n_loc=%lld, D=%lld)",
               (long long)n_loc, (long long)D);
  const int64_t W = x->world, N = W * n_loc, pad = (n_loc + 3) / 4 * 4;
  NANS_REQUIRE(N < (1ll << 29), "xchg_layout: global batch too large");
  x->n_loc = n_loc;
  x->D = D;
  int64_t off = 0;
  x->feat_off = off;
  off += static_cast<int64_t>(align_up(static_cast<size_t>(2) * 2 * N * D * 2, 1024));
  x->lse_len = 2 * pad + 8;
  x->lse_off = off;
  off += static_cast<int64_t>(align_up(static_cast<size_t>(2) * W * x->lse_len * 4, 1024));
  x->fflag_off = off;
  off += static_cast<int64_t>(align_up(static_cast<size_t>(2) * W * (n_loc / NANS_XCHG_FLAG_ROWS) * 4, 1024));
  x->lflag_off = off;
  off += 1024;
  x->bytes = off;
  return NANS_OK;
}

static int launch_cast_push(bool remote, const nans_xchg_t* x, const void* img, const void* txt, int x_dtype,
                            int64_t ld_x, int feat_dtype, int normalize, void* I16_loc, void* T16_loc, void* stream,
                            uint32_t* stepvals = nullptr, int expect_slot = -1, int k_begin = 1, int k_end = -1) {
  const char* who = remote ? "xchg_push" : "xchg_cast_local";
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  if ((rc = check_xchg(x, who)) != NANS_OK) return rc;
  NANS_REQUIRE(x_dtype == NANS_F32 || x_dtype == NANS_F16 || x_dtype == NANS_BF16, "%s: bad x_dtype", who);
  NANS_REQUIRE(feat_dtype == NANS_F16 || feat_dtype == NANS_BF16, "%s: feat_dtype must be NANS_F16 or NANS_BF16", who);
  NANS_REQUIRE(ld_x >= x->D, "%s: leading dimension smaller than D", who);
  NANS_REQUIRE(img && txt && (remote || (I16_loc && T16_loc)), "%s: null pointer", who);
  NANS_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0 && (reinterpret_cast<uintptr_t>(txt) & 15) == 0 &&
                   (ld_x * (x_dtype == NANS_F32 ? 4 : 2)) % 16 == 0,
               "%s: feature rows must be 16-byte aligned", who);
  if (k_end < 0) k_end = x->world;
  NANS_REQUIRE(!remote || (k_begin >= 1 && k_end <= x->world), "%s: peer range [%d, %d) outside [1, world)", who,
               k_begin, k_end);
  if (remote && k_begin >= k_end) return NANS_OK;
  PushParams p;
  fill_push_params(p, x);
  p.src[0] = img;
  p.src[1] = txt;
  p.loc16[0] = I16_loc;
  p.loc16[1] = T16_loc;
  p.x_dtype = x_dtype;
  p.feat_dtype = feat_dtype;
  p.normalize = normalize ? 1 : 0;
  p.ld_x = ld_x;
  p.stepvals = stepvals;
  p.expect_slot = expect_slot;
  p.k_begin = k_begin;
  p.k_end = k_end;
  const dim3 grid(push_grid_x(x->n_loc, remote), 2);
  if (remote) {
    // same shared-memory carve-out as the forward (which takes nearly all of it): an SM does not have to
    // drain and reconfigure before a forward CTA can join the push CTAs already resident there
    static bool carveout_set = false;
    if (!carveout_set) {
      cudaFuncSetAttribute(cast_push_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
      carveout_set = true;
    }
    cast_push_kernel<true><<<grid, PUSH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    cast_push_kernel<false><<<grid, PUSH_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  }
  NANS_CUDA_OK(cudaGetLastError());
  return NANS_OK;
}

extern "C" int nans_xchg_cast_local(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype,
                                    int64_t ld_x, int feat_dtype, int normalize, void* I16_loc, void* T16_loc,
                                    void* stream) {
  return launch_cast_push(false, x, img, txt, x_dtype, ld_x, feat_dtype, normalize, I16_loc, T16_loc, stream);
}

extern "C" int nans_xchg_push(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype, int64_t ld_x,
                              int feat_dtype, int normalize, void* stream) {
  return launch_cast_push(true, x, img, txt, x_dtype, ld_x, feat_dtype, normalize, nullptr, nullptr, stream);
}

extern "C" int nans_xchg_push_peers(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype,
                                    int64_t ld_x, int feat_dtype, int normalize, int k_begin, int k_end,
                                    void* stream) {
  return launch_cast_push(true, x, img, txt, x_dtype, ld_x, feat_dtype, normalize, nullptr, nullptr, stream, nullptr,
                          -1, k_begin, k_end);
}

extern "C" int nans_xchg_cast_push(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype,
                                   int64_t ld_x, int feat_dtype, int normalize, void* I16_loc, void* T16_loc,
                                   void* stream) {
  int rc = nans_xchg_cast_local(x, img, txt, x_dtype, ld_x, feat_dtype, normalize, I16_loc, T16_loc, stream);
  if (rc != NANS_OK) return rc;
  return nans_xchg_push(x, img, txt, x_dtype, ld_x, feat_dtype, normalize, stream);
}

// ---- the feature push on the COPY ENGINES -----------------------------------------------------------
// nans_xchg_cast_local_dma: nans_xchg_cast_local into ONE local buffer loc16 = [2 modalities][n_loc][D]
// (image rows, then text rows), plus the flag source stepvals[2][n_loc / 64] (every word = this step's
// number).  `slot` is the host's parity of this forward (number of forwards issued before it + 1, & 1);
// the kernel traps if the device's step counter disagrees (see cast_push_kernel).
extern "C" int nans_xchg_cast_local_dma(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype,
                                        int64_t ld_x, int feat_dtype, int normalize, void* loc16,
                                        uint32_t* stepvals, int slot, void* stream) {
  NANS_REQUIRE(x != nullptr && loc16 != nullptr && stepvals != nullptr && (slot == 0 || slot == 1),
               "xchg_cast_local_dma: bad arguments");
  uint8_t* l = static_cast<uint8_t*>(loc16);
  return launch_cast_push(false, x, img, txt, x_dtype, ld_x, feat_dtype, normalize, l,
                          l + static_cast<size_t>(x->n_loc) * x->D * 2, stream, stepvals, slot);
}

// nans_xchg_push_dma: per peer, in the order rank - 1, rank - 2, ..., TWO strided copies on `stream`:
// the rank's image + text rows (loc16) into slot `slot` of the peer's gathered buffers, then the flag
// words (stepvals) into the peer's flag table.  cudaMemcpy2DAsync on peer-mapped addresses: the copy
// engines move the data over NVLink, no SM is involved, nothing can starve or be starved by the forward
// running meanwhile, and copies of one stream complete in order (flags after the rows they announce).
extern "C" int nans_xchg_push_dma(const nans_xchg_t* x, const void* loc16, const uint32_t* stepvals, int slot,
                                  void* stream) {
  return nans_xchg_push_dma_peers(x, loc16, stepvals, slot, 1, x != nullptr ? x->world : 1, stream);
}

extern "C" int nans_xchg_push_dma_peers(const nans_xchg_t* x, const void* loc16, const uint32_t* stepvals, int slot,
                                        int k_begin, int k_end, void* stream) {
  int rc = check_device();
  if (rc != NANS_OK) return rc;
  if ((rc = check_xchg(x, "xchg_push_dma")) != NANS_OK) return rc;
  NANS_REQUIRE(loc16 && stepvals && (slot == 0 || slot == 1), "xchg_push_dma: bad arguments");
  const size_t row_bytes = static_cast<size_t>(x->D) * 2;
  const size_t N = static_cast<size_t>(x->world) * x->n_loc;
  const size_t blk_bytes = static_cast<size_t>(x->n_loc) * row_bytes;         // one modality of this rank
  const size_t mod_pitch = 2 * N * row_bytes;                                 // image slot 0 -> text slot 0
  const size_t nflag = static_cast<size_t>(x->n_loc / NANS_XCHG_FLAG_ROWS);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  NANS_REQUIRE(k_begin >= 1 && k_end <= x->world, "xchg_push_dma: peer range outside [1, world)");
  for (int k = k_begin; k < k_end; ++k) {
    const int dst = (x->rank - k + x->world) % x->world;
    uint8_t* d = static_cast<uint8_t*>(x->base[dst]);
    uint8_t* drows = d + x->feat_off + (static_cast<size_t>(slot) * N + static_cast<size_t>(x->rank) * x->n_loc) * row_bytes;
    // one strided copy for both modalities (measured: 8 MiB in 20 us; two linear 4 MiB copies took 12 + 13 us,
    // and a second stream did not run concurrently: P2P copies of one direction share a copy engine)
    NANS_CUDA_OK(cudaMemcpy2DAsync(drows, mod_pitch, loc16, blk_bytes, blk_bytes, 2, cudaMemcpyDeviceToDevice, st));
    uint8_t* dflag = d + x->fflag_off + static_cast<size_t>(x->rank) * nflag * 4;
    NANS_CUDA_OK(cudaMemcpy2DAsync(dflag, static_cast<size_t>(x->world) * nflag * 4, stepvals, nflag * 4, nflag * 4, 2,
                                   cudaMemcpyDeviceToDevice, st));
  }
  return NANS_OK;
}
