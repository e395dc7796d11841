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

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from . import kernels as K


@dataclass(frozen=True)
class LossConfig:
    group: Optional[object] = None      # torch.distributed process group, or None for local-only
    gather_with_grad: bool = False      # args.gather_with_grad (params.py:184)
    report_acc: bool = False            # args.report_training_batch_acc (params.py:52)
    feat_dtype: torch.dtype = torch.float16
    overlap_gather: bool = True


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
        I16 = _as_operand(full_img.detach(), cfg.feat_dtype)
        T16 = _as_operand(full_txt.detach(), cfg.feat_dtype)
        s_dev = s.detach().to(torch.float32).reshape(1).contiguous()
        label_begin = rank * n_loc

        # ---- phases of the column sweep -----------------------------------------------------
        # a phase = (T columns, I columns, global index of column 0, skipped range (begin, count))
        if W == 1:
            I_all, T_all = I16, T16
            phases = [(T16, I16, 0, (0, 0))]
            works = []
        else:
            I_all = torch.empty((N, D), dtype=cfg.feat_dtype, device=dev)
            T_all = torch.empty((N, D), dtype=cfg.feat_dtype, device=dev)
            works = [dist.all_gather_into_tensor(I_all, I16, group=cfg.group, async_op=True),
                     dist.all_gather_into_tensor(T_all, T16, group=cfg.group, async_op=True)]
            lo, hi = rank * n_loc, (rank + 1) * n_loc
            phases = [(T16, I16, lo, (0, 0))]               # local block: needs no remote data
            if n_loc % 256 == 0:
                # everything else in ONE launch over the gathered buffers, skipping the local tiles
                phases.append((T_all, I_all, 0, (lo, n_loc)))
            else:
                if lo > 0:
                    phases.append((T_all[:lo], I_all[:lo], 0, (0, 0)))
                if hi < N:
                    phases.append((T_all[hi:], I_all[hi:], hi, (0, 0)))
        slots = [K.fwd_phase_slots(n_loc, tc.shape[0] - sk[1], D) for tc, _, _, sk in phases]
        ws = K.fwd_workspace(n_loc, sum(slots), dev)
        slot = 0
        for i, (tc, ic, col0, sk) in enumerate(phases):
            if i == 1 or (i == 0 and works and not cfg.overlap_gather):
                for w in works:
                    w.wait()
                works = []
            K.fwd_phase(I16, T16, tc, ic, col_global_begin=col0, label_begin=label_begin,
                        s_dev=s_dev, with_acc=cfg.report_acc, ws=ws, slot_begin=slot,
                        skip_begin=sk[0], skip_count=sk[1])
            slot += slots[i]
        for w in works:
            w.wait()
        lse, scalars, packed = K.fwd_finalize(n_loc, slot, label_begin, s_dev, cfg.report_acc, ws)

        # ---- one small exchange: per-row lse (for the backward) and the 8 partial scalars -------
        if W > 1:
            L = packed.numel()
            pad = (L - 8) // 2
            gathered = torch.empty((W, L), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered.view(-1), packed, group=cfg.group)
            lse_all = gathered[:, :2 * pad].view(W, 2, pad)[:, :, :n_loc].permute(1, 0, 2).reshape(2, N)
            scalars = gathered[:, 2 * pad:].sum(dim=0)   # == all_reduce(SUM) of the partial sums
        else:
            lse_all = lse

        # loss = (sum_i + sum_j) / 2N (train.py:112-115); d loss / d s likewise; acc = hits / N
        red = scalars * _scalar_coefs(N, dev)
        loss = red[0] + red[1]
        dscale = red[2] + red[3]
        acc_i2t = red[4]
        acc_t2i = red[5]

        ctx.save_for_backward(I16, T16, I_all, T_all, s_dev, lse_all, dscale)
        ctx.cfg = cfg
        ctx.meta = (W, label_begin, int(row_begin), chunk_img.shape[0], chunk_img.dtype,
                    chunk_txt.dtype)
        ctx.mark_non_differentiable(acc_i2t, acc_t2i)
        return loss, acc_i2t, acc_t2i

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2):
        I16, T16, I_all, T_all, s_dev, lse_all, dscale = ctx.saved_tensors
        W, label_begin, row_begin, rows, dt_i, dt_t = ctx.meta
        cfg = ctx.cfg
        need_feat = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        dI = dT = None
        if need_feat:
            g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
            out_dt = dt_i if dt_i == dt_t else torch.float32
            dI, dT = K.bwd(I16, T16, T_all, I_all, label_begin=label_begin, s_dev=s_dev,
                           lse_all=lse_all, grad_out=g,
                           grad_mult=float(W) if cfg.gather_with_grad else 1.0,
                           row_begin=row_begin, row_count=rows, out_dtype=out_dt)
            dI = dI.to(dt_i) if ctx.needs_input_grad[0] else None
            dT = dT.to(dt_t) if ctx.needs_input_grad[1] else None
        ds = g_loss * dscale if ctx.needs_input_grad[2] else None
        return dI, dT, ds, None, None, None, None


def clip_contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor,
                          logit_scale: torch.Tensor, *, group=None, gather_with_grad: bool = False,
                          report_acc: bool = False, feat_dtype: torch.dtype = torch.float16,
                          full_image_features: Optional[torch.Tensor] = None,
                          full_text_features: Optional[torch.Tensor] = None, row_begin: int = 0,
                          overlap_gather: bool = True):
    """Contrastive loss of unit-norm features against arange labels.

    Returns (loss, acc) with acc = None or {"i2t": t, "t2i": t} exactly as train.py:109-126.
    `logit_scale` is the already exponentiated scale s (what CLIP.forward returns, model.py:415).
    Pass `full_*` + `row_begin` on the gradient-accumulation path: the block the loss is computed
    on, of which `image_features` / `text_features` are rows [row_begin, row_begin + B).
    """
    if full_image_features is None:
        full_image_features, full_text_features, row_begin = image_features, text_features, 0
    cfg = LossConfig(group=group, gather_with_grad=gather_with_grad, report_acc=report_acc,
                     feat_dtype=feat_dtype, overlap_gather=overlap_gather)
    loss, i2t, t2i = _ClipLossFn.apply(image_features, text_features, logit_scale,
                                       full_image_features, full_text_features, row_begin, cfg)
    acc = {"i2t": i2t, "t2i": t2i} if report_acc else None
    return loss, acc
