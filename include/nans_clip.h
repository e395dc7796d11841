/*
 * nans_clip.h — C-ABI of the B200 (sm_100a) contrastive-loss / top-k retrieval hot path.
 *
 * This is the drop-in boundary.  The reference (n571e/NanS-CLIP, a Chinese-CLIP
 * fork) has no FFI of its own: the path is three Python call sites.  Each entry point below names
 * the reference lines it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions (all functions):
 *   - extern "C", return int: 0 = ok, <0 = error code (NANS_ERR_*); nans_last_error() returns a
 *     thread-local message for the most recent failure on the calling thread.
 *   - raw device pointers + explicit sizes; no torch types; a cudaStream_t passed as void*.
 *   - nothing is allocated: the caller passes outputs and a workspace whose size comes from the
 *     matching *_workspace_bytes query.  No device synchronisation, no global mutable state
 *     (one cached driver entry point), CUDA-graph capturable, safe on concurrent streams.
 *   - scalars that live on the device in the reference (logit_scale.exp(), the upstream grad of
 *     the loss) are passed as DEVICE pointers so that no call forces a host sync.
 *   - feature matrices are row-major [rows, D] with a leading dimension in elements.
 *   - there is NO CPU fallback: every compute call fails with NANS_ERR_DEVICE unless the
 *     current device is compute capability 10.x (sm_100a code only).
 */
#ifndef NANS_CLIP_H_
#define NANS_CLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NANS_VERSION 100 /* 0.1.0 */

/* element types */
#define NANS_F32 0
#define NANS_F16 1
#define NANS_BF16 2

/* error codes */
#define NANS_OK 0
#define NANS_ERR_ARG (-1)       /* bad argument (null pointer, size, alignment, dtype) */
#define NANS_ERR_DEVICE (-2)    /* current device is not sm_100 / no device */
#define NANS_ERR_CUDA (-3)      /* a CUDA runtime / driver call failed */
#define NANS_ERR_WORKSPACE (-4) /* workspace too small */

/* flags for nans_clip_loss_fwd */
#define NANS_LOSS_WITH_ACC 1 /* also count in-batch top-1 hits (train.py:117-121) */
/* nans_clip_loss_fwd_phase only: sweep one of the two strips (image rows x text columns / text rows
 * x image columns).  Lets the image strip start as soon as the gathered TEXT features have arrived
 * while the image features are still in flight.  The two single-strip launches of a column range
 * must use the SAME slot range (each fills its half of every slot). */
#define NANS_LOSS_STRIP_IMG 2
#define NANS_LOSS_STRIP_TXT 4

/* ---- plumbing -------------------------------------------------------------------------- */

int nans_version(void);
const char* nans_last_error(void);
/* 0 if the current CUDA device can run this library (cc 10.x), else NANS_ERR_DEVICE. */
int nans_device_check(void);

/* ---- (1) L2-normalise + 16-bit cast ---------------------------------------------------- */
/*
 * Replaces cn_clip/clip/model.py:412-413 (`x / x.norm(dim=-1, keepdim=True)`) fused with the cast
 * to the tensor-core operand type.  One pass over x.
 *   x        [rows, D] (ld_x elements between rows), dtype x_dtype (NANS_F32/F16/BF16)
 *   y16      [rows, D] contiguous, dtype y16_dtype (NANS_F16/BF16), or NULL
 *   y32      [rows, D] contiguous fp32 normalised copy (what the reference returns), or NULL
 *   inv_norm [rows] fp32 1/||x||, or NULL (kept for the normalise backward)
 *   normalize != 0: divide by the row norm; == 0: cast only (features already unit-norm)
 * D must be a multiple of 8.  Zero rows give inf/nan exactly as the reference's division does.
 */
int nans_l2norm_cast(const void* x, int x_dtype, int64_t rows, int64_t D, int64_t ld_x,
                     void* y16, int y16_dtype, float* y32, float* inv_norm, int normalize,
                     void* stream);

/*
 * Backward of the normalisation (autograd of model.py:412-413):
 *   dx = inv_norm * (dy - y * <dy, y>),  y = x * inv_norm.
 *   x [rows, D] x_dtype (ld_x), dy [rows, D] fp32 contiguous, dx [rows, D] fp32 contiguous.
 */
int nans_l2norm_bwd(const void* x, int x_dtype, int64_t ld_x, const float* inv_norm,
                    const float* dy, int64_t rows, int64_t D, float* dx, void* stream);

/* ---- (2) fused contrastive forward ------------------------------------------------------ */
/*
 * Replaces cn_clip/training/train.py:87-88 (+103-104), 109-121: the (s*I)@T^T logits, both
 * cross-entropies against arange labels and the optional argmax accuracy — without ever writing
 * the logits.  Each rank computes two strips: I_loc x T_cols^T (image->text softmax over columns)
 * and T_loc x I_cols^T (text->image softmax).
 *
 * The column sweep may be issued in PHASES (e.g. the local block while the all-gather of the
 * other ranks' features is still in flight, then the rest).  A phase covers `ncols` columns given
 * by two column operands; column j of the phase is GLOBAL column col_global_begin + j.
 * Row r of the local block has its label (the positive pair) at global column label_begin + r.
 * Every phase writes partial (max, sum) statistics into workspace slots
 * [slot_begin, slot_begin + nans_clip_loss_fwd_phase_slots(...)); nans_clip_loss_fwd_finalize
 * merges `total_slots` slots.
 *
 * Columns [skip_col_begin, skip_col_begin + skip_col_count) of the operands are NOT swept by this
 * phase (e.g. the local block inside the gathered buffer, already covered by the local phase);
 * both values must be multiples of 256.  Size the slots with
 * nans_clip_loss_fwd_phase_slots(n_loc, ncols - skip_col_count, D).
 *
 *   I_loc, T_loc   [n_loc, D] 16-bit (feat_dtype), leading dims ld_loc
 *   T_cols, I_cols [ncols, D] 16-bit, leading dim ld_cols
 *   s_dev          device pointer to the fp32 logit scale s = exp(logit_scale)
 */
int64_t nans_clip_loss_fwd_phase_slots(int64_t n_loc, int64_t ncols, int64_t D);
/* Same for a launch with `flags` (a single-strip launch splits its columns further). */
int64_t nans_clip_loss_fwd_phase_slots_flags(int64_t n_loc, int64_t ncols, int64_t D, int flags);
size_t nans_clip_loss_fwd_workspace_bytes(int64_t n_loc, int64_t total_slots);

int nans_clip_loss_fwd_phase(const void* I_loc, const void* T_loc, int64_t ld_loc,
                             const void* T_cols, const void* I_cols, int64_t ld_cols,
                             int feat_dtype, int64_t n_loc, int64_t ncols, int64_t D,
                             int64_t col_global_begin, int64_t label_begin,
                             int64_t skip_col_begin, int64_t skip_col_count, const float* s_dev,
                             int flags, void* ws, size_t ws_bytes, int64_t slot_begin,
                             void* stream);

/*
 * Merge the slots.  Outputs (all device, fp32):
 *   lse_img_loc [n_loc]  log-sum-exp of row i of logits_per_image, in the kernels' BASE-2 domain:
 *                        lse2_i = log2 sum_j 2^(s * log2(e) * cos_ij) = logsumexp_j(logits_ij) / ln 2
 *   lse_txt_loc [n_loc]  the same for row j of logits_per_text
 *                        (base 2 so that the backward reuses them without a rounding round trip)
 *   scalars[8]: [0] sum_i (lse_img_i - logit_ii)     (x 1/(2N) summed over ranks = loss, part 1)
 *               [1] sum_j (lse_txt_j - logit_jj)
 *               [2] sum_i (E_softmax_img_i[cos] - cos_ii)   (d loss / d s numerator, part 1)
 *               [3] sum_j (E_softmax_txt_j[cos] - cos_jj)
 *               [4] #rows whose image->text argmax is the label   (if NANS_LOSS_WITH_ACC)
 *               [5] #rows whose text->image argmax is the label
 *               [6],[7] reserved (0)
 */
int nans_clip_loss_fwd_finalize(int64_t n_loc, int64_t total_slots, int64_t label_begin,
                                const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                float* lse_img_loc, float* lse_txt_loc, float* scalars,
                                void* stream);

/* The two entry points above with the extras of the incremental gradient-accumulation path
 * (python: nans_clip_b200/accum.py):
 *   _phase_rows   : only rows [label_row_begin, label_row_begin + label_row_count) have their label
 *                   column inside this phase's columns (label_begin then refers to those rows:
 *                   label column of row r = label_begin + r in global column coordinates); the other
 *                   rows contribute to the log-sum-exp only.  nans_clip_loss_fwd_phase = (0, n_loc).
 *   _finalize_rows: additionally writes per-row terms, row_stats[3][2 * n_loc] (strip-major like the
 *                   lse arrays): [0] lse_r - logit_rr, [1] E_softmax_r[cos] - cos_rr, [2] the arg-max
 *                   column as int32 bits (-1 without NANS_LOSS_WITH_ACC).  NULL = not wanted. */
int nans_clip_loss_fwd_phase_rows(const void* I_loc, const void* T_loc, int64_t ld_loc,
                                  const void* T_cols, const void* I_cols, int64_t ld_cols,
                                  int feat_dtype, int64_t n_loc, int64_t ncols, int64_t D,
                                  int64_t col_global_begin, int64_t label_begin,
                                  int64_t label_row_begin, int64_t label_row_count,
                                  int64_t skip_col_begin, int64_t skip_col_count, const float* s_dev,
                                  int flags, void* ws, size_t ws_bytes, int64_t slot_begin,
                                  void* stream);
int nans_clip_loss_fwd_finalize_rows(int64_t n_loc, int64_t total_slots, int64_t label_begin,
                                     const float* s_dev, int flags, void* ws, size_t ws_bytes,
                                     float* lse_img_loc, float* lse_txt_loc, float* scalars,
                                     float* row_stats, void* stream);

/* Convenience: one phase over all columns + finalize (single-GPU / after a blocking gather). */
int nans_clip_loss_fwd(const void* I_loc, const void* T_loc, int64_t ld_loc, const void* T_all,
                       const void* I_all, int64_t ld_all, int feat_dtype, int64_t n_loc,
                       int64_t N, int64_t D, int64_t label_begin, const float* s_dev, int flags,
                       float* lse_img_loc, float* lse_txt_loc, float* scalars, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- (3) fused contrastive backward ------------------------------------------------------ */
/*
 * Replaces the autograd backward of train.py:87-115: recomputes the logit tiles, forms
 * G = softmax_rows + softmax_cols - 2*delta on chip and contracts it with the column features.
 *   dI_loc[r] = coef * sum_j G[r, j] * T_all[j],   dT_loc[r] = coef * sum_i G[i, r] * I_all[i]
 *   coef = grad_out * s * grad_mult / (2 N)      (grad_mult = W under --gather-with-grad)
 * for the rows [grad_row_begin, grad_row_begin + grad_row_count) of the local block (the chunk
 * rows of the accumulate path, train.py:48-51; pass 0, n_loc for the plain path).
 * d loss / d s is produced by the forward (scalars[2], [3]) — nothing to do here.
 *
 *   lse_img_all, lse_txt_all [N] fp32: the finalize outputs (base-2 lse) of all ranks, in global
 *                                      row order (each array 16-byte aligned)
 *   grad_out_dev: device pointer to the fp32 upstream gradient of the loss
 *   dI_loc, dT_loc [grad_row_count, D] contiguous, dtype out_dtype (NANS_F32/F16/BF16)
 */
size_t nans_clip_loss_bwd_workspace_bytes(int64_t grad_row_count, int64_t N, int64_t D);

/* The launch schedule nans_clip_loss_bwd would use for (gradient rows, N, D) on the current device
 * (148 SMs are assumed when there is none): diagnostics and the CPU test-suite.  Host only, no launch.
 * out[0] kernel (0 single CTA, 1 wide pairs, 2 narrow pairs, 3 narrow pairs persistent), [1] column
 * tile width, [2] passes over the logits, [3] A resident, [4] row blocks per strip, [5] column tiles,
 * [6] ring stages, [7] dynamic shared memory bytes, [8] grid (CTAs), [9] uniform column splits,
 * [10] units run unsplit, [11] column splits of the remaining units, [12] persistent: tiles a unit's
 * own pair sweeps (0 = equal ranges), [13] persistent: units, [14] persistent: CTA pairs,
 * [15] outputs zeroed first (0 none, 1 rows of the split units, 2 all). */
#define NANS_BWD_PLAN_FIELDS 16
int nans_clip_loss_bwd_plan(int64_t grad_row_count, int64_t N, int64_t D, int64_t* out, int n_out);

int nans_clip_loss_bwd(const void* I_loc, const void* T_loc, int64_t ld_loc, const void* T_all,
                       const void* I_all, int64_t ld_all, int feat_dtype, int64_t n_loc,
                       int64_t N, int64_t D, int64_t label_begin, const float* s_dev,
                       const float* lse_img_all, const float* lse_txt_all,
                       const float* grad_out_dev, float grad_mult, int64_t grad_row_begin,
                       int64_t grad_row_count, void* dI_loc, void* dT_loc, int out_dtype,
                       void* ws, size_t ws_bytes, void* stream);

/* nans_clip_loss_bwd with the lse min / max supplied by nans_clip_loss_exchange_finish (saves the
 * backward's own min/max launch); lse_minmax = NULL behaves exactly like nans_clip_loss_bwd. */
int nans_clip_loss_bwd_minmax(const void* I_loc, const void* T_loc, int64_t ld_loc, const void* T_all,
                              const void* I_all, int64_t ld_all, int feat_dtype, int64_t n_loc,
                              int64_t N, int64_t D, int64_t label_begin, const float* s_dev,
                              const float* lse_img_all, const float* lse_txt_all, const int* lse_minmax,
                              const float* grad_out_dev, float grad_mult, int64_t grad_row_begin,
                              int64_t grad_row_count, void* dI_loc, void* dT_loc, int out_dtype,
                              void* ws, size_t ws_bytes, void* stream);

/* After the cross-rank exchange of the forward (one all-gather of, per rank, [lse_img (pad floats) |
 * lse_txt (pad) | scalars[8]] = the `packed` layout nans_clip_loss_fwd_finalize's outputs have when
 * lse_txt_loc = lse_img_loc + pad and scalars = lse_img_loc + 2 pad): ONE launch that
 *   - writes the rank-major table lse_all[2][ld] (row 0 image->text, row 1 text->image; ld >= world * n_loc)
 *     the backward reads,
 *   - reduces the partial scalars over the ranks: out[0] = loss (train.py:112-115), out[1] = d loss / d s,
 *     out[2], out[3] = the two accuracies (train.py:117-121),
 *   - writes lse_minmax[2] for nans_clip_loss_bwd_minmax.
 * Replaces the host-side permute / sum / scale ops and the backward's min/max launch. */
int nans_clip_loss_exchange_finish(const float* gathered, int64_t world, int64_t n_loc, int64_t pad,
                                   float* lse_all, int64_t ld, float* out, int* lse_minmax, void* stream);

/* ---- (2x/3x) multi-GPU exchange over NVLink peer memory --------------------------------------- */
/*
 * Replaces the collectives of cn_clip/training/train.py:53-84 (two feature all-gathers per get_loss call)
 * and the lse / scalar exchange of this library's own multi-rank path with a PUSH data plane: the cast
 * kernel of rank r writes every 16-bit row block straight into the gathered buffers of all peers (plain
 * st.global on peer-mapped memory, NVLink / NVSwitch), then raises a per-64-row flag there; the forward's
 * TMA producer polls the flag of a tile just before loading it, so remote tiles are consumed as they
 * land and no collective kernel sits between the cast and the backward.  The per-rank log-sum-exps and
 * partial scalars travel the same way.  Everything is plain kernels on the caller's stream: capturable
 * in a CUDA graph (the step counter lives in device memory), no NCCL, no host synchronisation.
 *
 * Memory: every rank allocates ONE exchange buffer with nans_peer_alloc (cudaMalloc + a 64-byte CUDA IPC
 * handle), the host exchanges the handles (any transport: torch.distributed, MPI, a file), and every
 * rank maps its peers' buffers with nans_peer_open.  nans_xchg_t then describes the job: the buffers as
 * mapped in THIS process and the layout inside them (identical on all ranks; sizes from
 * nans_xchg_layout).  Features are double-buffered by step parity: a rank can be one forward ahead of a
 * peer that is still in the previous step's backward without overwriting what that peer reads.
 *
 * Requirements: n_loc a multiple of 256 (whole column tiles per rank), world <= NANS_MAX_PEERS, all
 * ranks on one NVLink domain with peer access; one process per GPU (kernels of different ranks wait on
 * each other's flags: never run two ranks of one job on the same GPU concurrently).
 */
#define NANS_MAX_PEERS 16
#define NANS_XCHG_FLAG_ROWS 64 /* rows per arrival flag */

typedef struct {
  int32_t world, rank;
  void* base[NANS_MAX_PEERS]; /* base[r]: rank r's exchange buffer as mapped here (base[rank] = own)  */
  int64_t n_loc, D;           /* rows per rank, feature width (elements) of the CURRENT layout         */
  int64_t feat_off;           /* 16-bit [2 modalities: 0 image, 1 text][2 slots][world * n_loc][D]      */
  int64_t lse_off;            /* fp32   [2 slots][world][lse_len]    lse_len = 2 * pad4(n_loc) + 8      */
  int64_t lse_len;
  int64_t fflag_off;          /* uint32 [2 modalities][world][n_loc / 64]   value = step of arrival     */
  int64_t lflag_off;          /* uint32 [world]                                                        */
  int64_t bytes;              /* total bytes the layout needs                                          */
  uint32_t* epoch;            /* LOCAL device word (not peer-mapped): completed forward exchanges      */
} nans_xchg_t;

/* cudaMalloc `bytes` (zero-filled) on the current device and export it: handle64 receives the 64-byte
 * cudaIpcMemHandle_t.  nans_peer_free releases it. */
int nans_peer_alloc(size_t bytes, void** ptr, void* handle64);
int nans_peer_free(void* ptr);
/* Map a peer's buffer from its handle (peer access is enabled lazily); nans_peer_close unmaps. */
int nans_peer_open(const void* handle64, void** ptr);
int nans_peer_close(void* ptr);
/* Zero `bytes` of a buffer on `stream` (a new layout moves the flag words: stale bytes must not read as
 * "arrived").  The caller brackets it with its own cross-rank barriers. */
int nans_peer_zero(void* ptr, size_t bytes, void* stream);
/* Fills the layout fields (n_loc, D, *_off, lse_len, bytes) of `x` for a job of x->world ranks. */
int nans_xchg_layout(nans_xchg_t* x, int64_t n_loc, int64_t D);

/* Kernel (1) fused with the feature exchange, in two independent launches so that the NVLink traffic runs
 * UNDER the forward.  Both read the same source rows img / txt ([n_loc, D] of x_dtype NANS_F32/F16/BF16,
 * row pitch ld_x) and cast (optionally L2-normalise) them to the 16-bit operand type:
 *   nans_xchg_cast_local: into a local copy (I16_loc / T16_loc, [n_loc, D] contiguous: the row operand of
 *       the strips) and into slot (step & 1) of this rank's OWN gathered buffers; raises the own flags.
 *       HBM-bound, a few microseconds.  Main stream; the forward goes behind it.
 *   nans_xchg_push: into the same slot of every PEER's gathered buffers (plain st.global on peer-mapped
 *       memory), peers in the order rank-1, rank-2, ... so that each destination is served by one source
 *       at a time; raises the peers' flags block by block.  NVLink-bound.  Meant for a SIDE stream that
 *       forks BEFORE nans_xchg_cast_local is launched: its (capped) grid is then resident before the
 *       forward's one-CTA-per-SM grid arrives, and the forward consumes remote tiles as they land.  It
 *       waits for nothing.  The caller joins the side stream before the sources may change.
 *   nans_xchg_cast_push: both on one stream (no overlap; tests, simple callers). */
int nans_xchg_cast_local(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype, int64_t ld_x,
                         int feat_dtype, int normalize, void* I16_loc, void* T16_loc, void* stream);
int nans_xchg_push(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype, int64_t ld_x,
                   int feat_dtype, int normalize, void* stream);
int nans_xchg_cast_push(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype, int64_t ld_x,
                        int feat_dtype, int normalize, void* I16_loc, void* T16_loc, void* stream);
/* nans_xchg_push / nans_xchg_push_dma restricted to the peers rank - k, k in [k_begin, k_end) (1 <= k < world):
 * the two engines can share the peers of one step (hybrid push: the copy engines serve the peers whose rows
 * are needed first, a small push kernel — forked before the local cast — the ones needed last). */
int nans_xchg_push_peers(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype, int64_t ld_x,
                         int feat_dtype, int normalize, int k_begin, int k_end, void* stream);
int nans_xchg_push_dma_peers(const nans_xchg_t* x, const void* loc16, const uint32_t* stepvals, int slot,
                             int k_begin, int k_end, void* stream);

/* The same exchange with the remote half on the COPY ENGINES (the default of the python layer): no SM
 * takes part in the NVLink traffic, so it neither slows the forward it runs under nor depends on being
 * co-resident with it.
 *   nans_xchg_cast_local_dma: nans_xchg_cast_local into ONE buffer loc16 = [2][n_loc][D] (image rows, then
 *       text rows) plus the flag source stepvals[2][n_loc / 64] (each word = this step's number).
 *   nans_xchg_push_dma: per peer (rank-1, rank-2, ...) two cudaMemcpy2DAsync on `stream` (a side stream
 *       behind the local cast): the rows into the peer's slot, then the flag words into its flag table.
 * `slot` = the HOST's parity of this forward ((forwards issued before + 1) & 1): copy destinations are
 * host addresses.  The kernels use the device's step counter; nans_xchg_cast_local_dma traps when the two
 * disagree (a CUDA graph holding an ODD number of steps was replayed) instead of reading a stale slot. */
int nans_xchg_cast_local_dma(const nans_xchg_t* x, const void* img, const void* txt, int x_dtype, int64_t ld_x,
                             int feat_dtype, int normalize, void* loc16, uint32_t* stepvals, int slot,
                             void* stream);
int nans_xchg_push_dma(const nans_xchg_t* x, const void* loc16, const uint32_t* stepvals, int slot, void* stream);

/* Kernel (2) over the gathered buffers: both strips of this rank against all world * n_loc columns in
 * ONE launch.  Every unit walks the column tiles source by source starting with its own rank's (local,
 * already there) and waits for a tile's flags right before its TMA loads.  Fills workspace slots
 * [0, nans_clip_loss_fwd_xchg_slots(...)); finish with nans_clip_loss_fwd_finalize_push. */
int64_t nans_clip_loss_fwd_xchg_slots(int64_t n_loc, int64_t world, int64_t D);
int nans_clip_loss_fwd_xchg(const nans_xchg_t* x, const void* I16_loc, const void* T16_loc, int feat_dtype,
                            const float* s_dev, int flags, void* ws, size_t ws_bytes, void* stream);

/* nans_clip_loss_fwd_finalize for this rank's rows (label_begin = rank * n_loc) that also PUSHES the packed
 * result [lse_img (pad) | lse_txt (pad) | scalars[8]] into slot (step & 1), row `rank`, of every rank's
 * lse table and flags it.  lse_*_loc / scalars: the local copies, as in nans_clip_loss_fwd_finalize. */
int nans_clip_loss_fwd_finalize_push(const nans_xchg_t* x, int64_t total_slots, const float* s_dev, int flags,
                                     void* ws, size_t ws_bytes, float* lse_img_loc, float* lse_txt_loc,
                                     float* scalars, void* stream);

/* nans_clip_loss_exchange_finish on the pushed table: waits (on the device) until every rank's packed
 * row of this step has arrived, then writes lse_all / out / lse_minmax exactly like
 * nans_clip_loss_exchange_finish, stores the step number in step_out[0] (a device word owned by this
 * call: the backward finds its slot through it) and publishes the step in x->epoch. */
int nans_clip_loss_exchange_finish_xchg(const nans_xchg_t* x, float* lse_all, int64_t ld, float* out,
                                        int* lse_minmax, uint32_t* step_out, void* stream);

/* nans_clip_loss_bwd_minmax with the column operands taken from the gathered buffers of the step
 * recorded in step_dev (the word nans_clip_loss_exchange_finish_xchg wrote).  D <= 1024. */
int nans_clip_loss_bwd_xchg(const nans_xchg_t* x, const uint32_t* step_dev, const void* I16_loc,
                            const void* T16_loc, int feat_dtype, const float* s_dev, const float* lse_img_all,
                            const float* lse_txt_all, const int* lse_minmax, const float* grad_out_dev,
                            float grad_mult, int64_t grad_row_begin, int64_t grad_row_count, void* dI_loc,
                            void* dT_loc, int out_dtype, void* ws, size_t ws_bytes, void* stream);

/* ---- (3b) label-smoothed variant ---------------------------------------------------------- */
/*
 * Replaces the fork's train_lora.py:95-110 (`contrastive_loss`: F.cross_entropy(logits, arange,
 * label_smoothing=eps) in both directions).  The smoothed loss is the plain fused loss plus O(N D)
 * terms (derivation in csrc/smooth.cu), so it needs two small HBM-bound kernels around (2) and (3):
 *   nans_label_smooth_stats: stats[0..D) = sum_i I_i, stats[D..2D) = sum_i T_i,
 *                            stats[2D] = sum_i I_i.T_i over this rank's rows (the function zeroes
 *                            `stats` first; sum over ranks with an all-reduce).  I, T fp32, row pitch ld.
 *       loss_eps = loss + eps*s/N * stats[2D] - eps*s/N^2 * (stats[0..D) . stats[D..2D))
 *       dloss/ds likewise without the factor s.
 *   nans_label_smooth_bwd: dI[r] += a*T[r] - (a*inv_n)*Tsum, dT[r] += a*I[r] - (a*inv_n)*Isum with
 *                            a = grad_out * s * coef, coef = grad_mult * eps / N, inv_n = 1 / N;
 *                            dI, dT fp32 [rows, D] contiguous (the outputs of nans_clip_loss_bwd),
 *                            I_rows / T_rows the same rows of the fp32 features (pitch ld),
 *                            stats the GLOBAL sums.
 */
int nans_label_smooth_stats(const float* I, const float* T, int64_t ld, int64_t rows, int64_t D,
                            float* stats, void* stream);
int nans_label_smooth_bwd(float* dI, float* dT, const float* I_rows, const float* T_rows, int64_t ld,
                          int64_t rows, int64_t D, const float* stats, const float* s_dev,
                          const float* grad_out_dev, float coef, float inv_n, void* stream);

/* ---- feature files (host only; no GPU) ----------------------------------------------------- */
/*
 * Parse the reference's JSONL feature files ({"image_id": int, "feature": [floats]} per line,
 * cn_clip/eval/extract_features.py:179-181, read back by make_topk_predictions.py:57-65 with json.loads)
 * into ids[rows] and feats[rows, D] float32, bit-identical to json.loads + np.float32 (correctly rounded
 * std::from_chars to double, then the float cast), lines parsed in parallel.
 *   nans_jsonl_scan : rows = non-blank lines, D = length of the first line's feature list
 *   nans_jsonl_parse: fills the arrays; fails on a malformed line or a different feature length
 *                     (callers fall back to a general JSON parser).  n_threads <= 0: all cores.
 */
int nans_jsonl_scan(const char* buf, int64_t len, const char* id_key, int64_t* rows, int64_t* D);
int nans_jsonl_parse(const char* buf, int64_t len, const char* id_key, int64_t rows, int64_t D,
                     int64_t* ids, float* feats, int n_threads);
/*   nans_jsonl_parse_rows: ids[rows] for every line, feature lists only for the lines
 *                     [row_begin, row_begin + row_count) into feats[row_count, D]: one rank's gallery
 *                     shard (the reference parses the whole file on its single GPU's host,
 *                     make_topk_predictions.py:57-65; a sharded run must not do that W times). */
int nans_jsonl_parse_rows(const char* buf, int64_t len, const char* id_key, int64_t rows, int64_t D,
                          int64_t row_begin, int64_t row_count, int64_t* ids, float* feats, int n_threads);

/* ---- (4) top-k inner-product retrieval ---------------------------------------------------- */
/*
 * Replaces cn_clip/eval/make_topk_predictions.py:71-85 (and _tr.py): for each query the k gallery
 * rows with the largest inner product, ordered by (score descending, gallery index ascending) —
 * the order Python's stable sorted(..., reverse=True) produces.
 * Pass 1 streams the 16-bit gallery shard through tensor-core tiles keeping k_cand candidates per
 * query; pass 2 rescores the candidates exactly in fp32 (Q32, G32; pass NULL for both to keep the
 * 16-bit scores) and emits the top k.
 *   Q16 [Q, D], G16 [G, D] 16-bit (feat_dtype); Q32 [Q, D], G32 [G, D] fp32 (contiguous)
 *   k <= k_cand, k_cand in {16, 32}
 *   out_scores [Q, k] fp32, out_index [Q, k] int64 = gallery_index_offset + row in this shard
 *   (index -1 / score -inf pad rows when G < k)
 */
size_t nans_topk_ip_workspace_bytes(int64_t Q, int64_t G, int64_t D, int k_cand);

int nans_topk_ip(const void* Q16, const void* G16, int feat_dtype, const float* Q32,
                 const float* G32, int64_t Q, int64_t G, int64_t D, int k, int k_cand,
                 int64_t gallery_index_offset, float* out_scores, int64_t* out_index, void* ws,
                 size_t ws_bytes, void* stream);

/*
 * Merge per-shard results: in_scores/in_index [n_shards, Q, k] -> out [Q, k], same ordering rule.
 */
int nans_topk_merge(const float* in_scores, const int64_t* in_index, int n_shards, int64_t Q,
                    int k, float* out_scores, int64_t* out_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NANS_CLIP_H_ */
