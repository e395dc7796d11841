"""Thin torch-tensor wrappers over the C-ABI (one function per exported kernel entry point).

Everything here takes CUDA tensors, hands raw device pointers and the current stream to
libnans_clip.so and returns tensors.  No arithmetic happens in Python and nothing falls back to
PyTorch ops: a missing library or a non-sm_100 device raises.

The distributed orchestration (loss.py, retrieval.py) only talks to the functions of this module,
so the CPU test-suite can substitute them to exercise the multi-rank host logic under gloo.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import (NANS_BF16, NANS_F16, NANS_F32, NANS_LOSS_STRIP_IMG, NANS_LOSS_STRIP_TXT,
                   NANS_LOSS_WITH_ACC, check)

_DT = {torch.float32: NANS_F32, torch.float16: NANS_F16, torch.bfloat16: NANS_BF16}

# Number of kernels of libnans_clip.so launched through this module (bench.py reports it).
LAUNCHES = 0


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt}: expected float32, float16 or bfloat16") from None


class _on_device:
    """`with torch.cuda.device(dev)` plus the raw handle of that device's current stream, without
    the Python-level device bookkeeping when `dev` already is the current device (the usual case:
    one process per GPU) — that bookkeeping was ~15 % of the host time of a loss step."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev: torch.device):
        self.idx = dev.index if dev.index is not None else torch._C._cuda_getDevice()

    def __enter__(self) -> int:
        self.prev = torch._C._cuda_getDevice()
        if self.prev != self.idx:
            torch.cuda.set_device(self.idx)
        return torch._C._cuda_getCurrentRawStream(self.idx)

    def __exit__(self, *exc) -> None:
        if self.prev != self.idx:
            torch.cuda.set_device(self.prev)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nans_clip_b200 kernels need CUDA tensors (there is no CPU path)")


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor with unit inner stride (a row pitch is allowed)."""
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D feature matrix, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


# --------------------------------------------------------------------------------------------
# (1) L2 normalise + cast
# --------------------------------------------------------------------------------------------
def l2norm_cast(x: torch.Tensor, out_dtype: torch.dtype | None = torch.bfloat16, *,
                normalize: bool = True, want_fp32: bool = False, want_inv_norm: bool = False):
    """Returns (y16 | None, y32 | None, inv_norm | None).  Reference: model.py:412-413."""
    _require_cuda(x)
    x = _rowmajor(x)
    rows, D = x.shape
    y16 = torch.empty((rows, D), dtype=out_dtype, device=x.device) if out_dtype is not None else None
    y32 = torch.empty((rows, D), dtype=torch.float32, device=x.device) if want_fp32 else None
    inv = torch.empty((rows,), dtype=torch.float32, device=x.device) if want_inv_norm else None
    with _on_device(x.device) as stream:
        check(_lib.load().nans_l2norm_cast(
            x.data_ptr(), dtype_code(x.dtype), rows, D, x.stride(0) if rows > 1 else D,
            _ptr(y16), dtype_code(out_dtype) if out_dtype is not None else NANS_BF16,
            _ptr(y32), _ptr(inv), 1 if normalize else 0, stream))
    _count(1 if rows else 0)
    return y16, y32, inv


def l2norm_bwd(x: torch.Tensor, inv_norm: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    _require_cuda(x, inv_norm, dy)
    x = _rowmajor(x)
    dy = dy.to(torch.float32).contiguous()
    rows, D = x.shape
    dx = torch.empty((rows, D), dtype=torch.float32, device=x.device)
    with _on_device(x.device) as stream:
        check(_lib.load().nans_l2norm_bwd(x.data_ptr(), dtype_code(x.dtype),
                                          x.stride(0) if rows > 1 else D, inv_norm.data_ptr(),
                                          dy.data_ptr(), rows, D, dx.data_ptr(), stream))
    _count(1 if rows else 0)
    return dx


# --------------------------------------------------------------------------------------------
# (2) fused forward
# --------------------------------------------------------------------------------------------
_STRIP_FLAG = {None: 0, "img": NANS_LOSS_STRIP_IMG, "txt": NANS_LOSS_STRIP_TXT}


def fwd_phase_slots(n_loc: int, ncols: int, D: int, strip: str | None = None) -> int:
    """Workspace slots a phase over `ncols` columns occupies (`strip`: see fwd_phase)."""
    if strip is None:
        return int(_lib.load().nans_clip_loss_fwd_phase_slots(n_loc, ncols, D))
    return int(_lib.load().nans_clip_loss_fwd_phase_slots_flags(n_loc, ncols, D, _STRIP_FLAG[strip]))


def fwd_workspace(n_loc: int, total_slots: int, device) -> torch.Tensor:
    nbytes = int(_lib.load().nans_clip_loss_fwd_workspace_bytes(n_loc, total_slots))
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def clone_workspace(ws: torch.Tensor) -> torch.Tensor:
    """A copy of a forward workspace with its slots (the incremental accumulate path keeps the slots of
    the cached features and overwrites one column chunk's slots per call)."""
    return ws.clone()


def fwd_phase(I_loc, T_loc, T_cols, I_cols, *, col_global_begin: int, label_begin: int,
              s_dev: torch.Tensor, with_acc: bool, ws: torch.Tensor, slot_begin: int,
              skip_begin: int = 0, skip_count: int = 0, strip: str | None = None,
              label_rows: tuple[int, int] | None = None) -> None:
    """One phase of the forward column sweep; columns [skip_begin, skip_begin + skip_count) of the
    operands (multiples of 256) are left to another phase.  `strip` = "img" sweeps only the image
    rows against `T_cols` (`I_cols` is not read), "txt" only the text rows against `I_cols`; the
    two launches of one column range share their slot range.  `label_rows` = (first row, count): only
    these rows have their label column among this phase's columns (`label_begin` refers to them:
    label column of row r = label_begin + r); default all rows."""
    _require_cuda(I_loc, T_loc, T_cols, I_cols, s_dev, ws)
    n_loc, D = I_loc.shape
    ncols = T_cols.shape[0]
    assert I_loc.dtype == T_loc.dtype == T_cols.dtype == I_cols.dtype
    assert I_loc.stride(0) == T_loc.stride(0) and T_cols.stride(0) == I_cols.stride(0)
    lr0, lrn = label_rows if label_rows is not None else (0, n_loc)
    with _on_device(I_loc.device) as stream:
        check(_lib.load().nans_clip_loss_fwd_phase_rows(
            I_loc.data_ptr(), T_loc.data_ptr(), I_loc.stride(0), T_cols.data_ptr(),
            I_cols.data_ptr(), T_cols.stride(0), dtype_code(I_loc.dtype), n_loc, ncols, D,
            col_global_begin, label_begin, lr0, lrn, skip_begin, skip_count, s_dev.data_ptr(),
            (NANS_LOSS_WITH_ACC if with_acc else 0) | _STRIP_FLAG[strip], ws.data_ptr(), ws.numel(),
            slot_begin, stream))
    _count(1)


def fwd_finalize(n_loc: int, total_slots: int, label_begin: int, s_dev: torch.Tensor,
                 with_acc: bool, ws: torch.Tensor, want_row_stats: bool = False):
    """Returns (lse2 [2, n_loc] (row 0 image->text, row 1 text->image; base-2 log-sum-exp),
    scalars [8], packed) — `packed` is the single contiguous buffer [2 * pad + 8] both views live
    in (pad = n_loc rounded up to 4, keeping row 1 16-byte aligned), so that a multi-rank caller
    can exchange lse and the partial scalars with ONE all-gather.  With `want_row_stats` a fourth
    value follows: [3, 2, n_loc] fp32 per-row terms (loss term, d/ds term, arg-max column as int32
    bits), strip-major like lse — what the incremental accumulate path sums over row subsets."""
    pad = (n_loc + 3) // 4 * 4
    packed = torch.empty((2 * pad + 8,), dtype=torch.float32, device=ws.device)
    lse = packed[:2 * pad].view(2, pad)[:, :n_loc]
    scalars = packed[2 * pad:]
    rows = torch.empty((3, 2, n_loc), dtype=torch.float32, device=ws.device) if want_row_stats else None
    with _on_device(ws.device) as stream:
        check(_lib.load().nans_clip_loss_fwd_finalize_rows(
            n_loc, total_slots, label_begin, s_dev.data_ptr(),
            NANS_LOSS_WITH_ACC if with_acc else 0, ws.data_ptr(), ws.numel(),
            lse[0].data_ptr(), lse[1].data_ptr(), scalars.data_ptr(), _ptr(rows), stream))
    _count(1)
    if want_row_stats:
        return lse, scalars, packed, rows
    return lse, scalars, packed


def exchange_finish(gathered: torch.Tensor, n_loc: int):
    """`gathered` [W, 2 * pad + 8]: every rank's `packed` buffer of fwd_finalize, all-gathered.
    Returns (lse_all [2, N] rank-major, out [4] = loss, d loss / d s, acc i2t, acc t2i,
    lse_minmax int32 [2] for bwd) from one launch."""
    _require_cuda(gathered)
    W, L = gathered.shape
    pad = (L - 8) // 2
    N = W * n_loc
    ld = (N + 3) // 4 * 4
    dev = gathered.device
    assert gathered.dtype == torch.float32 and gathered.is_contiguous() and pad >= n_loc
    lse_all = torch.empty((2, ld), dtype=torch.float32, device=dev)[:, :N]
    out = torch.empty((4,), dtype=torch.float32, device=dev)
    mm = torch.empty((2,), dtype=torch.int32, device=dev)
    with _on_device(dev) as stream:
        check(_lib.load().nans_clip_loss_exchange_finish(gathered.data_ptr(), W, n_loc, pad, lse_all.data_ptr(), ld,
                                                         out.data_ptr(), mm.data_ptr(), stream))
    _count(1)
    return lse_all, out, mm


# --------------------------------------------------------------------------------------------
# (3) fused backward
# --------------------------------------------------------------------------------------------
def bwd(I_loc, T_loc, T_all, I_all, *, label_begin: int, s_dev: torch.Tensor,
        lse_all: torch.Tensor, grad_out: torch.Tensor, grad_mult: float, row_begin: int,
        row_count: int, out_dtype: torch.dtype, lse_minmax: torch.Tensor | None = None):
    """lse_all: [2, N] fp32 (image->text, text->image) in global row order.  Returns (dI, dT).
    `lse_minmax`: the int32 [2] of exchange_finish (else the backward computes it itself)."""
    _require_cuda(I_loc, T_loc, T_all, I_all, s_dev, lse_all, grad_out)
    n_loc, D = I_loc.shape
    N = T_all.shape[0]
    dev = I_loc.device
    dI = torch.empty((row_count, D), dtype=out_dtype, device=dev)
    dT = torch.empty((row_count, D), dtype=out_dtype, device=dev)
    lib = _lib.load()
    nbytes = int(lib.nans_clip_loss_bwd_workspace_bytes(row_count, N, D))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    assert lse_all.shape == (2, N) and lse_all.dtype == torch.float32
    if lse_all.stride(1) != 1 or lse_all[0].data_ptr() % 16 or lse_all[1].data_ptr() % 16:
        pad = (N + 3) // 4 * 4
        buf = torch.empty((2, pad), dtype=torch.float32, device=dev)[:, :N]
        buf.copy_(lse_all)
        lse_all = buf
    assert grad_out.dtype == torch.float32 and grad_out.numel() == 1
    with _on_device(dev) as stream:
        check(lib.nans_clip_loss_bwd_minmax(
            I_loc.data_ptr(), T_loc.data_ptr(), I_loc.stride(0), T_all.data_ptr(),
            I_all.data_ptr(), T_all.stride(0), dtype_code(I_loc.dtype), n_loc, N, D, label_begin,
            s_dev.data_ptr(), lse_all[0].data_ptr(), lse_all[1].data_ptr(), _ptr(lse_minmax), grad_out.data_ptr(),
            float(grad_mult), row_begin, row_count, dI.data_ptr(), dT.data_ptr(),
            dtype_code(out_dtype), ws.data_ptr(), ws.numel(), stream))
    _count((1 if lse_minmax is not None else 2) + (0 if out_dtype == torch.float32 else 2))  # [lse min/max +] backward (+ 2 casts)
    return dI, dT


# --------------------------------------------------------------------------------------------
# (2x/3x) multi-GPU exchange over NVLink peer memory (csrc/exchange.cu; descriptors: exchange.py)
# --------------------------------------------------------------------------------------------
def _desc_ref(desc):
    import ctypes
    return ctypes.byref(desc)


def _xchg_sources(desc, img: torch.Tensor, txt: torch.Tensor):
    _require_cuda(img, txt)
    img, txt = _rowmajor(img), _rowmajor(txt)
    n_loc, D = img.shape
    assert txt.shape == (n_loc, D) and img.dtype == txt.dtype and (n_loc, D) == (desc.n_loc, desc.D)
    if img.stride(0) != txt.stride(0):
        img, txt = img.contiguous(), txt.contiguous()
    return img, txt


def xchg_push(desc, img: torch.Tensor, txt: torch.Tensor, feat_dtype: torch.dtype, normalize: bool = False,
              stream: "torch.cuda.Stream | None" = None, peers: "tuple[int, int] | None" = None):
    """Kernel (1) into every PEER's gathered slot over NVLink (`stream`: a side stream forked before
    xchg_cast_local, so that the traffic runs under the forward; default the current stream).  `peers` =
    (k_begin, k_end): only the peers rank - k for k in that range.  Returns the (possibly re-laid-out)
    sources it reads: keep them alive until the side stream has been joined."""
    img, txt = _xchg_sources(desc, img, txt)
    k0, k1 = peers if peers is not None else (1, int(desc.world))
    with _on_device(img.device) as cur:
        check(_lib.load().nans_xchg_push_peers(_desc_ref(desc), img.data_ptr(), txt.data_ptr(), dtype_code(img.dtype),
                                               img.stride(0), dtype_code(feat_dtype), 1 if normalize else 0, k0, k1,
                                               stream.cuda_stream if stream is not None else cur))
    _count(1 if k1 > k0 else 0)
    return img, txt


def xchg_cast_local(desc, img: torch.Tensor, txt: torch.Tensor, feat_dtype: torch.dtype, normalize: bool = False):
    """Kernel (1) into the local 16-bit copies (returned: I16, T16) and this rank's own gathered slot."""
    img, txt = _xchg_sources(desc, img, txt)
    n_loc, D = img.shape
    I16 = torch.empty((n_loc, D), dtype=feat_dtype, device=img.device)
    T16 = torch.empty((n_loc, D), dtype=feat_dtype, device=img.device)
    with _on_device(img.device) as stream:
        check(_lib.load().nans_xchg_cast_local(_desc_ref(desc), img.data_ptr(), txt.data_ptr(), dtype_code(img.dtype),
                                               img.stride(0), dtype_code(feat_dtype), 1 if normalize else 0,
                                               I16.data_ptr(), T16.data_ptr(), stream))
    _count(1)
    return I16, T16


def xchg_cast_local_dma(desc, img: torch.Tensor, txt: torch.Tensor, feat_dtype: torch.dtype, slot: int,
                        normalize: bool = False):
    """xchg_cast_local into one [2, n_loc, D] buffer + the flag source words.  Returns (loc16, stepvals);
    loc16[0] / loc16[1] are the image / text row operands."""
    img, txt = _xchg_sources(desc, img, txt)
    n_loc, D = img.shape
    loc16 = torch.empty((2, n_loc, D), dtype=feat_dtype, device=img.device)
    stepvals = torch.empty((2, n_loc // 64), dtype=torch.int32, device=img.device)
    with _on_device(img.device) as stream:
        check(_lib.load().nans_xchg_cast_local_dma(_desc_ref(desc), img.data_ptr(), txt.data_ptr(), dtype_code(img.dtype),
                                                   img.stride(0), dtype_code(feat_dtype), 1 if normalize else 0,
                                                   loc16.data_ptr(), stepvals.data_ptr(), int(slot), stream))
    _count(1)
    return loc16, stepvals


def xchg_push_dma(desc, loc16: torch.Tensor, stepvals: torch.Tensor, slot: int, stream: "torch.cuda.Stream",
                  peers: "tuple[int, int] | None" = None) -> None:
    """The rows and their flags into the peers' buffers with copy-engine copies on `stream`; `peers` =
    (k_begin, k_end) as in xchg_push."""
    _require_cuda(loc16, stepvals)
    k0, k1 = peers if peers is not None else (1, int(desc.world))
    check(_lib.load().nans_xchg_push_dma_peers(_desc_ref(desc), loc16.data_ptr(), stepvals.data_ptr(), int(slot), k0, k1,
                                               stream.cuda_stream))


def xchg_cast_push(desc, img: torch.Tensor, txt: torch.Tensor, feat_dtype: torch.dtype, normalize: bool = False):
    """xchg_cast_local + xchg_push on the current stream (no overlap)."""
    I16, T16 = xchg_cast_local(desc, img, txt, feat_dtype, normalize)
    xchg_push(desc, img, txt, feat_dtype, normalize)
    return I16, T16


def fwd_xchg_slots(n_loc: int, world: int, D: int) -> int:
    return int(_lib.load().nans_clip_loss_fwd_xchg_slots(n_loc, world, D))


def fwd_xchg(desc, I16: torch.Tensor, T16: torch.Tensor, s_dev: torch.Tensor, with_acc: bool, ws: torch.Tensor) -> None:
    """Both strips of this rank over the gathered buffers, tiles consumed as their flags go up."""
    _require_cuda(I16, T16, s_dev, ws)
    assert I16.is_contiguous() and T16.is_contiguous() and I16.dtype == T16.dtype
    with _on_device(I16.device) as stream:
        check(_lib.load().nans_clip_loss_fwd_xchg(_desc_ref(desc), I16.data_ptr(), T16.data_ptr(), dtype_code(I16.dtype),
                                                  s_dev.data_ptr(), NANS_LOSS_WITH_ACC if with_acc else 0,
                                                  ws.data_ptr(), ws.numel(), stream))
    _count(1)


def fwd_finalize_push(desc, total_slots: int, s_dev: torch.Tensor, with_acc: bool, ws: torch.Tensor):
    """fwd_finalize for this rank's rows + push of the packed result into every rank's lse table.
    Returns (lse2 [2, n_loc], scalars [8]) — the local copies."""
    n_loc = int(desc.n_loc)
    pad = (n_loc + 3) // 4 * 4
    packed = torch.empty((2 * pad + 8,), dtype=torch.float32, device=ws.device)
    lse = packed[:2 * pad].view(2, pad)[:, :n_loc]
    scalars = packed[2 * pad:]
    with _on_device(ws.device) as stream:
        check(_lib.load().nans_clip_loss_fwd_finalize_push(_desc_ref(desc), total_slots, s_dev.data_ptr(),
                                                           NANS_LOSS_WITH_ACC if with_acc else 0, ws.data_ptr(),
                                                           ws.numel(), lse[0].data_ptr(), lse[1].data_ptr(),
                                                           scalars.data_ptr(), stream))
    _count(1)
    return lse, scalars


def exchange_finish_xchg(desc, device):
    """Waits on the device for every rank's packed row, then as exchange_finish.  Returns (lse_all [2, N],
    out [4], lse_minmax int32 [2], step int32 [1] — the step word the backward takes)."""
    N = int(desc.world) * int(desc.n_loc)
    ld = (N + 3) // 4 * 4
    lse_all = torch.empty((2, ld), dtype=torch.float32, device=device)[:, :N]
    out = torch.empty((4,), dtype=torch.float32, device=device)
    mm = torch.empty((2,), dtype=torch.int32, device=device)
    step = torch.empty((1,), dtype=torch.int32, device=device)
    with _on_device(device) as stream:
        check(_lib.load().nans_clip_loss_exchange_finish_xchg(_desc_ref(desc), lse_all.data_ptr(), ld, out.data_ptr(),
                                                              mm.data_ptr(), step.data_ptr(), stream))
    _count(1)
    return lse_all, out, mm, step


def bwd_xchg(desc, step: torch.Tensor, I16: torch.Tensor, T16: torch.Tensor, *, s_dev: torch.Tensor,
             lse_all: torch.Tensor, lse_minmax: torch.Tensor, grad_out: torch.Tensor, grad_mult: float,
             row_begin: int, row_count: int, out_dtype: torch.dtype):
    """bwd with the column operands taken from the gathered buffers of the forward's step."""
    _require_cuda(I16, T16, s_dev, lse_all, grad_out, step, lse_minmax)
    n_loc, D = I16.shape
    N = int(desc.world) * n_loc
    dev = I16.device
    dI = torch.empty((row_count, D), dtype=out_dtype, device=dev)
    dT = torch.empty((row_count, D), dtype=out_dtype, device=dev)
    lib = _lib.load()
    ws = torch.empty(int(lib.nans_clip_loss_bwd_workspace_bytes(row_count, N, D)), dtype=torch.uint8, device=dev)
    assert lse_all.shape == (2, N) and lse_all.stride(1) == 1 and lse_all[1].data_ptr() % 16 == 0
    assert grad_out.dtype == torch.float32 and grad_out.numel() == 1
    with _on_device(dev) as stream:
        check(lib.nans_clip_loss_bwd_xchg(_desc_ref(desc), step.data_ptr(), I16.data_ptr(), T16.data_ptr(),
                                          dtype_code(I16.dtype), s_dev.data_ptr(), lse_all[0].data_ptr(),
                                          lse_all[1].data_ptr(), lse_minmax.data_ptr(), grad_out.data_ptr(),
                                          float(grad_mult), row_begin, row_count, dI.data_ptr(), dT.data_ptr(),
                                          dtype_code(out_dtype), ws.data_ptr(), ws.numel(), stream))
    _count(1 + (0 if out_dtype == torch.float32 else 2))
    return dI, dT


# --------------------------------------------------------------------------------------------
# (3b) label smoothing (train_lora.py:95-110)
# --------------------------------------------------------------------------------------------
def smooth_stats(I32: torch.Tensor, T32: torch.Tensor) -> torch.Tensor:
    """[2 D + 1] fp32: column sums of I, column sums of T, sum_i I_i.T_i (this rank's rows)."""
    _require_cuda(I32, T32)
    I32, T32 = _rowmajor(I32), _rowmajor(T32)
    assert I32.dtype == T32.dtype == torch.float32 and I32.shape == T32.shape
    rows, D = I32.shape
    if rows > 1 and I32.stride(0) != T32.stride(0):
        I32, T32 = I32.contiguous(), T32.contiguous()
    stats = torch.empty((2 * D + 1,), dtype=torch.float32, device=I32.device)
    with _on_device(I32.device) as stream:
        check(_lib.load().nans_label_smooth_stats(I32.data_ptr(), T32.data_ptr(),
                                                  I32.stride(0) if rows > 1 else D, rows, D,
                                                  stats.data_ptr(), stream))
    _count(1 if rows else 0)
    return stats


def smooth_bwd(dI: torch.Tensor, dT: torch.Tensor, I_rows: torch.Tensor, T_rows: torch.Tensor,
               stats: torch.Tensor, s_dev: torch.Tensor, grad_out: torch.Tensor, coef: float,
               inv_n: float) -> None:
    """In place: dI += a T_rows - a/N Tsum, dT += a I_rows - a/N Isum, a = grad_out * s * coef."""
    _require_cuda(dI, dT, I_rows, T_rows, stats, s_dev, grad_out)
    I_rows, T_rows = _rowmajor(I_rows), _rowmajor(T_rows)
    rows, D = dI.shape
    assert dI.dtype == dT.dtype == I_rows.dtype == T_rows.dtype == torch.float32
    assert dI.is_contiguous() and dT.is_contiguous() and I_rows.shape == T_rows.shape == (rows, D)
    if rows > 1 and I_rows.stride(0) != T_rows.stride(0):
        I_rows, T_rows = I_rows.contiguous(), T_rows.contiguous()
    with _on_device(dI.device) as stream:
        check(_lib.load().nans_label_smooth_bwd(dI.data_ptr(), dT.data_ptr(), I_rows.data_ptr(),
                                                T_rows.data_ptr(), I_rows.stride(0) if rows > 1 else D,
                                                rows, D, stats.data_ptr(), s_dev.data_ptr(),
                                                grad_out.data_ptr(), float(coef), float(inv_n), stream))
    _count(1 if rows else 0)


# --------------------------------------------------------------------------------------------
# (4) retrieval
# --------------------------------------------------------------------------------------------
def topk_ip(Q16, G16, Q32, G32, k: int, k_cand: int, gallery_index_offset: int = 0):
    """Returns (scores [Q, k] fp32, index [Q, k] int64).  Reference: make_topk_predictions.py:71-85."""
    _require_cuda(Q16, G16, Q32, G32)
    Qn, D = Q16.shape
    Gn = G16.shape[0]
    dev = Q16.device
    Q16, G16 = Q16.contiguous(), G16.contiguous()
    if Q32 is not None:
        Q32, G32 = Q32.contiguous(), G32.contiguous()
    scores = torch.empty((Qn, k), dtype=torch.float32, device=dev)
    index = torch.empty((Qn, k), dtype=torch.int64, device=dev)
    lib = _lib.load()
    ws = torch.empty(int(lib.nans_topk_ip_workspace_bytes(Qn, Gn, D, k_cand)), dtype=torch.uint8,
                     device=dev)
    with _on_device(dev) as stream:
        check(lib.nans_topk_ip(Q16.data_ptr(), G16.data_ptr(), dtype_code(Q16.dtype), _ptr(Q32),
                               _ptr(G32), Qn, Gn, D, k, k_cand, gallery_index_offset,
                               scores.data_ptr(), index.data_ptr(), ws.data_ptr(), ws.numel(),
                               stream))
    _count(0 if not Qn else (1 if not Gn else (3 if Gn >= 3 * 16384 else 2)))  # [pre-pass +] sweep + finalize
    return scores, index


def topk_merge(scores: torch.Tensor, index: torch.Tensor):
    """[n_shards, Q, k] x2 -> ([Q, k], [Q, k]) in (score desc, index asc) order."""
    _require_cuda(scores, index)
    n_shards, Qn, k = scores.shape
    scores, index = scores.contiguous(), index.contiguous()
    out_s = torch.empty((Qn, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Qn, k), dtype=torch.int64, device=scores.device)
    with _on_device(scores.device) as stream:
        check(_lib.load().nans_topk_merge(scores.data_ptr(), index.data_ptr(), n_shards, Qn, k,
                                          out_s.data_ptr(), out_i.data_ptr(), stream))
    _count(1)
    return out_s, out_i
