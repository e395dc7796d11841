"""Global-batch contrastive loss on the fused sm_100a kernels (host orchestration).

Mirrors the arithmetic of cn_clip/training/train.py:52-121 (`get_loss` from `logit_scale.mean()`
down to the accuracy dict) without ever forming the N x N logits:

  * features are cast once to the 16-bit tensor-core operand type (kernel 1, cast-only: the
    reference's CLIP.forward has already normalised them, model.py:412-413);
  * with a process group, the 16-bit features are all-gathered asynchronously
    (`all_gather_into_tensor`, replaces train.py:59-60 / 72-73) while the local x local block of
    the forward already runs (phase 0); the remaining columns follow as phases 1 and 2;
  * each rank sweeps only its N_local x N strips; 6 scalars (two loss sums, two d(scale) sums, two
    hit counts) are all-reduced, and the per-row log-sum-exps are all-gathered for the backward;
  * the backward (kernel 3) emits dI_local, dT_local for the rows that carry gradient.

Gradient conventions kept from the reference (verified against it under gloo):
  gather_with_grad=False -> dI_local = dL_global/dI_local exactly ("substitute local slot");
  gather_with_grad=True  -> W x that (the summed all-gather backward of identical loss copies);
  d(logit_scale) is the full global derivative on every rank.
The reference's local-first slot order (train.py:75-84) only permutes rows/columns of the logits;
loss, accuracy and gradients are invariant to it, so columns stay in rank order here.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import exchange
from . import kernels as K


@dataclass(frozen=True)
class LossConfig:
    group: Optional[object] = None      # torch.distributed process group, or None for local-only
    gather_with_grad: bool = False      # args.gather_with_grad (params.py:184)
    report_acc: bool = False            # args.report_training_batch_acc (params.py:52)
    feat_dtype: torch.dtype = torch.float16
    overlap_gather: bool = True
    label_smoothing: float = 0.0        # F.cross_entropy(label_smoothing=...) of train_lora.py:95-110


def _world(group) -> tuple[int, int]:
    if group is None:
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


_COEFS: dict = {}


def _scalar_coefs(N: int, dev) -> torch.Tensor:
    key = (N, str(dev))
    t = _COEFS.get(key)
    if t is None:
        i2n = 1.0 / (2.0 * N)
        t = torch.tensor([i2n, i2n, i2n, i2n, 1.0 / N, 1.0 / N, 0.0, 0.0], dtype=torch.float32, device=dev)
        _COEFS[key] = t
    return t


def _as_operand(x: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """[n, D] contiguous 16-bit operand; a no-op when x already is one."""
    if x.dtype == dt and x.is_contiguous():
        return x
    y16, _, _ = K.l2norm_cast(x, dt, normalize=False)
    return y16


class _ClipLossFn(torch.autograd.Function):
    """forward(chunk_img, chunk_txt, s, full_img, full_txt, row_begin, cfg) -> (loss, i2t, t2i).

    `full_*` are the rank's whole [N_local, D] blocks (no grad); `chunk_*` are the rows
    [row_begin, row_begin + B) of them that receive gradient (the whole block outside the
    gradient-accumulation path, train.py:48-51)."""

    @staticmethod
    def forward(ctx, chunk_img, chunk_txt, s, full_img, full_txt, row_begin: int, cfg: LossConfig):
        W, rank = _world(cfg.group)
        n_loc, D = full_img.shape
        if full_txt.shape != (n_loc, D):
            raise ValueError("image and text feature blocks must have the same shape")
        N = W * n_loc
        dev = full_img.device
        s_dev = s.detach().to(torch.float32).reshape(1).contiguous()
        label_begin = rank * n_loc

        # ---- multi-GPU, NVLink push exchange (exchange.py / csrc/exchange.cu): no collective per step ----
        ex = exchange.for_group(cfg.group) if (W > 1 and full_img.is_cuda and exchange.eligible(n_loc, D, W)) else None
        desc = ex.ensure(n_loc, D) if ex is not None else None
        if desc is not None:
            # main stream: cast (local) -> forward -> finalize + lse push -> wait for the peers' lse;
            # side stream: the rows and their flags into the peers' buffers.  Default: copy-engine copies
            # behind the local cast (no SM involved: nothing to co-schedule with the forward).
            src_i, src_t = full_img.detach(), full_txt.detach()
            main = torch.cuda.current_stream(dev)
            mode = os.environ.get("NANS_PUSH", "dma")
            keep = None
            if mode in ("dma", "hybrid"):
                slot = (ex.forwards + 1) & 1
                # hybrid: the copy engines serve the peers whose rows are needed first (k = 1 .. kd), a small
                # push kernel — forked FIRST, so that its grid is resident before the forward's — the rest
                kd = W if mode == "dma" else min(W, 1 + int(os.environ.get("NANS_PUSH_DMA_PEERS", str((W + 1) // 2))))
                if kd < W:
                    ex.fork.record(main)
                    ex.push_stream_b.wait_event(ex.fork)
                    keep_sm = K.xchg_push(desc, src_i, src_t, cfg.feat_dtype, stream=ex.push_stream_b, peers=(kd, W))
                    ex.join_b.record(ex.push_stream_b)
                loc16, stepvals = K.xchg_cast_local_dma(desc, src_i, src_t, cfg.feat_dtype, slot)
                I16, T16 = loc16[0], loc16[1]
                ex.fork2.record(main)
                ex.push_stream.wait_event(ex.fork2)
                K.xchg_push_dma(desc, loc16, stepvals, slot, ex.push_stream, peers=(1, kd))
                if kd < W:
                    ex.push_stream.wait_event(ex.join_b)
                ex.join.record(ex.push_stream)
                keep = (loc16, stepvals, keep_sm if kd < W else None)
            elif mode == "sm":
                # push kernel on the SMs, forked FIRST so that its grid is resident before the forward's
                ex.fork.record(main)
                ex.push_stream.wait_event(ex.fork)
                keep = K.xchg_push(desc, src_i, src_t, cfg.feat_dtype, stream=ex.push_stream)
                ex.join.record(ex.push_stream)
                I16, T16 = K.xchg_cast_local(desc, src_i, src_t, cfg.feat_dtype)
            else:   # "serial": push kernel, then forward
                I16, T16 = K.xchg_cast_push(desc, src_i, src_t, cfg.feat_dtype)
            nslots = K.fwd_xchg_slots(n_loc, W, D)
            ws = K.fwd_workspace(n_loc, nslots, dev)
            K.fwd_xchg(desc, I16, T16, s_dev, cfg.report_acc, ws)
            K.fwd_finalize_push(desc, nslots, s_dev, cfg.report_acc, ws)
            lse_all, res, lse_minmax, step = K.exchange_finish_xchg(desc, dev)
            if keep is not None:
                main.wait_event(ex.join)   # the push has read its sources: they may change from here on
                del keep
            ex.forwards += 1
            loss, dscale, acc_i2t, acc_t2i = res[0], res[1], res[2], res[3]
            stats = I32 = T32 = None
            eps = float(cfg.label_smoothing)
            if eps != 0.0:
                I32 = full_img.detach().to(torch.float32)
                T32 = full_txt.detach().to(torch.float32)
                stats = K.smooth_stats(I32, T32)
                dist.all_reduce(stats, group=cfg.group)
                corr = (eps / N) * stats[2 * D] - (eps / (float(N) * N)) * torch.dot(stats[:D], stats[D:2 * D])
                loss = loss + s_dev[0] * corr
                dscale = dscale + corr
            ctx.save_for_backward(I16, T16, step, None, s_dev, lse_all, dscale, stats, I32, T32, lse_minmax)
            ctx.cfg = cfg
            ctx.xchg = (ex, desc, ex.forwards)
            ctx.meta = (W, label_begin, int(row_begin), chunk_img.shape[0], chunk_img.dtype, chunk_txt.dtype)
            ctx.mark_non_differentiable(acc_i2t, acc_t2i)
            return loss, acc_i2t, acc_t2i
        ctx.xchg = None

        I16 = _as_operand(full_img.detach(), cfg.feat_dtype)
        T16 = _as_operand(full_txt.detach(), cfg.feat_dtype)

        # ---- phases of the column sweep -----------------------------------------------------
        # a phase = (T columns, I columns, global index of column 0, skipped range (begin, count),
        #            strip (None = both), index of the gather it has to wait for (-1 = none))
        if W == 1:
            I_all, T_all = I16, T16
            phases = [(T16, I16, 0, (0, 0), None, -1)]
            works = []
        else:
            I_all = torch.empty((N, D), dtype=cfg.feat_dtype, device=dev)
            T_all = torch.empty((N, D), dtype=cfg.feat_dtype, device=dev)
            # text first: the image strip only reads T_all and runs while I_all is still in flight
            works = [dist.all_gather_into_tensor(T_all, T16, group=cfg.group, async_op=True),
                     dist.all_gather_into_tensor(I_all, I16, group=cfg.group, async_op=True)]
            lo, hi = rank * n_loc, (rank + 1) * n_loc
            phases = [(T16, I16, lo, (0, 0), None, -1 if cfg.overlap_gather else 1)]  # local block
            if n_loc % 256 == 0:
                # everything else in one launch per strip over the gathered buffers, skipping the
                # local tiles; both launches fill the same slots
                phases.append((T_all, I_all, 0, (lo, n_loc), "img", 0))
                phases.append((T_all, I_all, 0, (lo, n_loc), "txt", 1))
            else:
                if lo > 0:
                    phases.append((T_all[:lo], I_all[:lo], 0, (0, 0), None, 1))
                if hi < N:
                    phases.append((T_all[hi:], I_all[hi:], hi, (0, 0), None, 1))
        slots = [K.fwd_phase_slots(n_loc, tc.shape[0] - sk[1], D, st) for tc, _, _, sk, st, _ in phases]
        total_slots = sum(ns for ns, ph in zip(slots, phases) if ph[4] != "txt")
        ws = K.fwd_workspace(n_loc, total_slots, dev)
        slot = 0
        waited = -1
        for i, (tc, ic, col0, sk, st, need) in enumerate(phases):
            while waited < need:
                waited += 1
                works[waited].wait()
            if st == "txt":
                slot -= slots[i]      # the text strip fills the other half of the image strip's slots
            K.fwd_phase(I16, T16, tc, ic, col_global_begin=col0, label_begin=label_begin,
                        s_dev=s_dev, with_acc=cfg.report_acc, ws=ws, slot_begin=slot,
                        skip_begin=sk[0], skip_count=sk[1], strip=st)
            slot += slots[i]
        for w in works[waited + 1:]:
            w.wait()
        lse, scalars, packed = K.fwd_finalize(n_loc, slot, label_begin, s_dev, cfg.report_acc, ws)

        # ---- one small exchange: per-row lse (for the backward) and the 8 partial scalars -------
        lse_minmax = None
        if W > 1:
            L = packed.numel()
            gathered = torch.empty((W, L), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered.view(-1), packed, group=cfg.group)
            # one launch: rank-major lse table, the scalars summed over ranks and scaled
            # (loss = (sum_i + sum_j) / 2N, train.py:112-115; acc = hits / N), lse min/max for the backward
            lse_all, res, lse_minmax = K.exchange_finish(gathered, n_loc)
            loss, dscale, acc_i2t, acc_t2i = res[0], res[1], res[2], res[3]
        else:
            lse_all = lse
            red = scalars * _scalar_coefs(N, dev)
            loss = red[0] + red[1]
            dscale = red[2] + red[3]
            acc_i2t = red[4]
            acc_t2i = red[5]

        # ---- label smoothing: the plain loss plus O(N D) terms (csrc/smooth.cu) ------------------
        stats = I32 = T32 = None
        eps = float(cfg.label_smoothing)
        if eps != 0.0:
            I32 = full_img.detach().to(torch.float32)
            T32 = full_txt.detach().to(torch.float32)
            stats = K.smooth_stats(I32, T32)
            if W > 1:
                dist.all_reduce(stats, group=cfg.group)
            corr = (eps / N) * stats[2 * D] - (eps / (float(N) * N)) * torch.dot(stats[:D], stats[D:2 * D])
            loss = loss + s_dev[0] * corr
            dscale = dscale + corr

        ctx.save_for_backward(I16, T16, I_all, T_all, s_dev, lse_all, dscale, stats, I32, T32, lse_minmax)
        ctx.cfg = cfg
        ctx.meta = (W, label_begin, int(row_begin), chunk_img.shape[0], chunk_img.dtype,
                    chunk_txt.dtype)
        ctx.mark_non_differentiable(acc_i2t, acc_t2i)
        return loss, acc_i2t, acc_t2i

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2):
        I16, T16, I_all, T_all, s_dev, lse_all, dscale, stats, I32, T32, lse_minmax = ctx.saved_tensors
        W, label_begin, row_begin, rows, dt_i, dt_t = ctx.meta
        cfg = ctx.cfg
        need_feat = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dI = dT = None
        if need_feat:
            g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
            out_dt = dt_i if (dt_i == dt_t and stats is None) else torch.float32
            mult = float(W) if cfg.gather_with_grad else 1.0
            if ctx.xchg is not None:
                ex, desc, fwd_no = ctx.xchg
                if ex.forwards - fwd_no > 1 or ex.shape != (int(desc.n_loc), int(desc.D)):
                    # the gathered buffers are double-buffered by step: one later forward is fine (its
                    # features went to the other slot), two are not
                    raise RuntimeError("push exchange: backward() called after two later forwards of the same "
                                       "process group overwrote this step's gathered features; call backward "
                                       "right after the loss (as train.py does) or set NANS_EXCHANGE=nccl")
                step = I_all   # saved in I_all's place
                dI, dT = K.bwd_xchg(desc, step, I16, T16, s_dev=s_dev, lse_all=lse_all, lse_minmax=lse_minmax,
                                    grad_out=g, grad_mult=mult, row_begin=row_begin, row_count=rows, out_dtype=out_dt)
            else:
                dI, dT = K.bwd(I16, T16, T_all, I_all, label_begin=label_begin, s_dev=s_dev,
                               lse_all=lse_all, grad_out=g, grad_mult=mult,
                               row_begin=row_begin, row_count=rows, out_dtype=out_dt, lse_minmax=lse_minmax)
            if stats is not None:
                N = W * I16.shape[0]
                K.smooth_bwd(dI, dT, I32[row_begin:row_begin + rows], T32[row_begin:row_begin + rows], stats,
                             s_dev, g, mult * float(cfg.label_smoothing) / N, 1.0 / N)
            dI = dI.to(dt_i) if ctx.needs_input_grad[0] else None
            dT = dT.to(dt_t) if ctx.needs_input_grad[1] else None
        ds = g_loss * dscale if ctx.needs_input_grad[2] else None
        return dI, dT, ds, None, None, None, None


def clip_contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor,
                          logit_scale: torch.Tensor, *, group=None, gather_with_grad: bool = False,
                          report_acc: bool = False, feat_dtype: torch.dtype = torch.float16,
                          full_image_features: Optional[torch.Tensor] = None,
                          full_text_features: Optional[torch.Tensor] = None, row_begin: int = 0,
                          overlap_gather: bool = True, label_smoothing: float = 0.0):
    """Contrastive loss of unit-norm features against arange labels.

    Returns (loss, acc) with acc = None or {"i2t": t, "t2i": t} exactly as train.py:109-126.
    `logit_scale` is the already exponentiated scale s (what CLIP.forward returns, model.py:415).
    Pass `full_*` + `row_begin` on the gradient-accumulation path: the block the loss is computed
    on, of which `image_features` / `text_features` are rows [row_begin, row_begin + B).
    `label_smoothing` = the `label_smoothing` of F.cross_entropy in the fork's train_lora.py:105-108.
    """
    if not 0.0 <= float(label_smoothing) < 1.0:
        raise ValueError(f"label_smoothing must be in [0, 1), got {label_smoothing}")
    if full_image_features is None:
        full_image_features, full_text_features, row_begin = image_features, text_features, 0
    cfg = LossConfig(group=group, gather_with_grad=gather_with_grad, report_acc=report_acc,
                     feat_dtype=feat_dtype, overlap_gather=overlap_gather,
                     label_smoothing=float(label_smoothing))
    loss, i2t, t2i = _ClipLossFn.apply(image_features, text_features, logit_scale,
                                       full_image_features, full_text_features, row_begin, cfg)
    acc = {"i2t": i2t, "t2i": t2i} if report_acc else None
    return loss, acc
