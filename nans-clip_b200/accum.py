"""Incremental forward for the gradient-accumulation path.

Reference behaviour (cn_clip/training/train.py:205-247 + get_loss :34-51): one optimizer step caches
the features of all A chunks without gradient, then calls get_loss A times; call j re-forwards chunk j
with gradient, splices it between the cached chunks and recomputes the WHOLE N x N loss, although only
chunk j's B rows per rank (B W rows and columns of the global matrix) differ from the cached features.

Here the per-row log-sum-exp state is kept PER COLUMN CHUNK, so call j only computes
    (a) all local rows        x the B W new columns      (replaces column chunk j's partial state)
    (b) the B re-forwarded rows x all N columns           (their state is rebuilt from scratch)
i.e. 4 N_l N D / A flops instead of 2 N_l N D (x 2 strips), plus ONE full forward over the cached
features per optimizer step (the first call builds it).  The result is the same sum of the same
terms as the full recomputation (partial log-sum-exps merge exactly), so the reference's value and
gradients are reproduced to rounding; tests/ hold both paths to the same fixtures.

Layout: the gathered features are kept CHUNK-major ([A][W][B] rows instead of rank-major
[W][A][B]) so that "column chunk j" is one contiguous block; the loss is invariant to a column
permutation as long as labels follow, and they do: local row (c, i) of rank r has its label at
column c W B + r B + i.

The state lives on the cached feature tensors themselves (an attribute of accum_image_features[0]),
so it dies with them at the end of the optimizer step and can never be matched to a later step's
tensors.  `logit_scale` must not change between the calls of one step (it cannot: parameters are
updated after the step, train.py:249-262).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import kernels as K
from .loss import LossConfig, _as_operand, _scalar_coefs, _world

_ATTR = "_nans_accum_state"
MIN_CHUNKS = 4   # one extra full forward per step + 2/A per call: pays from A = 4 on


def eligible(cache_img, cache_txt, chunk_rows: int, W: int, label_smoothing: float) -> bool:
    """The incremental path needs equal chunks, tile-aligned column chunks (the chunk being replaced
    is skipped as whole 256-column tiles) and the plain cross-entropy."""
    if os.environ.get("NANS_ACCUM_INCREMENTAL", "1") == "0":
        return False
    A = len(cache_img)
    if A < MIN_CHUNKS or len(cache_txt) != A or label_smoothing != 0.0:
        return False
    if any(t.shape[0] != chunk_rows for t in cache_img) or any(t.shape[0] != chunk_rows for t in cache_txt):
        return False
    return (W * chunk_rows) % 256 == 0


class _State:
    """Everything that depends only on the cached (no-grad) features of one optimizer step."""

    def __init__(self, cache_img, cache_txt, s_dev: torch.Tensor, cfg: LossConfig):
        W, rank = _world(cfg.group)
        self.A, self.B = len(cache_img), cache_img[0].shape[0]
        A, B = self.A, self.B
        self.D = cache_img[0].shape[1]
        self.W, self.rank = W, rank
        self.WB = W * B
        self.n_loc, self.N = A * B, W * A * B
        self.fingerprint = tuple((id(t), t._version) for t in list(cache_img) + list(cache_txt))
        dev = cache_img[0].device
        dt = cfg.feat_dtype
        self.loc_I = _as_operand(torch.cat([t.detach() for t in cache_img]), dt)
        self.loc_T = _as_operand(torch.cat([t.detach() for t in cache_txt]), dt)
        if W == 1:
            self.C_I, self.C_T = self.loc_I, self.loc_T
        else:
            gi = torch.empty((W, A, B, self.D), dtype=dt, device=dev)
            gt = torch.empty((W, A, B, self.D), dtype=dt, device=dev)
            dist.all_gather_into_tensor(gi.view(-1), self.loc_I.view(-1), group=cfg.group)
            dist.all_gather_into_tensor(gt.view(-1), self.loc_T.view(-1), group=cfg.group)
            self.C_I = gi.permute(1, 0, 2, 3).reshape(self.N, self.D).contiguous()   # chunk-major
            self.C_T = gt.permute(1, 0, 2, 3).reshape(self.N, self.D).contiguous()
        # label column (chunk-major) of every local row, for the accuracy of rows outside chunk j
        c = torch.arange(A, device=dev).repeat_interleave(B)
        i = torch.arange(B, device=dev).repeat(A)
        self.label_cols = (c * self.WB + rank * B + i).to(torch.int32)
        # per-column-chunk partial state of ALL local rows over the cached features
        self.ns = K.fwd_phase_slots(self.n_loc, self.WB, self.D)
        self.ws0 = K.fwd_workspace(self.n_loc, A * self.ns, dev)
        for ch in range(A):
            blk = slice(ch * self.WB, (ch + 1) * self.WB)
            K.fwd_phase(self.loc_I, self.loc_T, self.C_T[blk], self.C_I[blk], col_global_begin=ch * self.WB,
                        label_begin=self.label_begin(ch), s_dev=s_dev, with_acc=cfg.report_acc, ws=self.ws0,
                        slot_begin=ch * self.ns, label_rows=(ch * B, B))
        # working copies that carry the re-forwarded chunk of the current call
        self.Wk_I, self.Wk_T = self.C_I.clone(), self.C_T.clone()
        self.L_I, self.L_T = self.loc_I.clone(), self.loc_T.clone()
        self.dirty: Optional[int] = None
        self.version = 0

    def label_begin(self, ch: int) -> int:
        """`label_begin` such that local row r of chunk `ch` has its label at column label_begin + r."""
        return ch * self.WB + self.rank * self.B - ch * self.B

    def matches(self, cache_img, cache_txt) -> bool:
        return self.fingerprint == tuple((id(t), t._version) for t in list(cache_img) + list(cache_txt))

    def install(self, j: int, nI16, nT16, NI_all, NT_all) -> None:
        """Put the re-forwarded chunk j into the working buffers (and the previous one back)."""
        B, WB = self.B, self.WB
        d = self.dirty
        if d is not None and d != j:
            self.Wk_I[d * WB:(d + 1) * WB].copy_(self.C_I[d * WB:(d + 1) * WB])
            self.Wk_T[d * WB:(d + 1) * WB].copy_(self.C_T[d * WB:(d + 1) * WB])
            self.L_I[d * B:(d + 1) * B].copy_(self.loc_I[d * B:(d + 1) * B])
            self.L_T[d * B:(d + 1) * B].copy_(self.loc_T[d * B:(d + 1) * B])
        self.Wk_I[j * WB:(j + 1) * WB].copy_(NI_all)
        self.Wk_T[j * WB:(j + 1) * WB].copy_(NT_all)
        self.L_I[j * B:(j + 1) * B].copy_(nI16)
        self.L_T[j * B:(j + 1) * B].copy_(nT16)
        self.dirty = j
        self.version += 1


def _get_state(cache_img, cache_txt, s_dev, cfg) -> _State:
    st = getattr(cache_img[0], _ATTR, None)
    if st is None or not st.matches(cache_img, cache_txt):
        st = _State(cache_img, cache_txt, s_dev, cfg)
        setattr(cache_img[0], _ATTR, st)
    return st


class _AccumLossFn(torch.autograd.Function):
    """forward(chunk_img, chunk_txt, s, cache_img, cache_txt, j, cfg) -> (loss, i2t, t2i)."""

    @staticmethod
    def forward(ctx, chunk_img, chunk_txt, s, cache_img, cache_txt, j: int, cfg: LossConfig):
        s_dev = s.detach().to(torch.float32).reshape(1).contiguous()
        st = _get_state(cache_img, cache_txt, s_dev, cfg)
        W, rank, A, B, WB, D = st.W, st.rank, st.A, st.B, st.WB, st.D
        n_loc, N = st.n_loc, st.N
        dev = chunk_img.device
        nI16 = _as_operand(chunk_img.detach(), cfg.feat_dtype)
        nT16 = _as_operand(chunk_txt.detach(), cfg.feat_dtype)
        if W == 1:
            NI_all, NT_all = nI16, nT16
        else:
            NI_all = torch.empty((WB, D), dtype=cfg.feat_dtype, device=dev)
            NT_all = torch.empty((WB, D), dtype=cfg.feat_dtype, device=dev)
            dist.all_gather_into_tensor(NT_all, nT16, group=cfg.group)
            dist.all_gather_into_tensor(NI_all, nI16, group=cfg.group)
        st.install(j, nI16, nT16, NI_all, NT_all)
        rows_j = slice(j * B, (j + 1) * B)

        # (a) all local rows x the new columns: replaces column chunk j's slots
        ws_a = K.clone_workspace(st.ws0)
        K.fwd_phase(st.L_I, st.L_T, NT_all, NI_all, col_global_begin=j * WB, label_begin=st.label_begin(j),
                    s_dev=s_dev, with_acc=cfg.report_acc, ws=ws_a, slot_begin=j * st.ns, label_rows=(j * B, B))
        lse_a, _, _, rows_a = K.fwd_finalize(n_loc, A * st.ns, 0, s_dev, cfg.report_acc, ws_a, want_row_stats=True)
        # (b) the re-forwarded rows x all columns (working buffers = cached features with chunk j replaced)
        ns_b = K.fwd_phase_slots(B, N, D)
        ws_b = K.fwd_workspace(B, ns_b, dev)
        lb = j * WB + rank * B
        K.fwd_phase(nI16, nT16, st.Wk_T, st.Wk_I, col_global_begin=0, label_begin=lb, s_dev=s_dev,
                    with_acc=cfg.report_acc, ws=ws_b, slot_begin=0)
        lse_b, sc_b, _ = K.fwd_finalize(B, ns_b, lb, s_dev, cfg.report_acc, ws_b)

        # ---- merge: rows outside chunk j from (a), chunk j's rows from (b) ------------------------
        pad = (n_loc + 3) // 4 * 4
        packed = torch.empty((2 * pad + 8,), dtype=torch.float32, device=dev)
        lse_loc = packed[:2 * pad].view(2, pad)[:, :n_loc]
        lse_loc.copy_(lse_a)
        lse_loc[:, rows_j] = lse_b
        terms = rows_a[:2].sum(dim=2) - rows_a[:2, :, rows_j].sum(dim=2)        # [2 kinds, 2 strips]
        scal = packed[2 * pad:]
        scal.zero_()
        scal[0:2] = terms[0] + sc_b[0:2]
        scal[2:4] = terms[1] + sc_b[2:4]
        if cfg.report_acc:
            hit = (rows_a[2].view(torch.int32) == st.label_cols[None, :]).to(torch.float32)
            scal[4:6] = hit.sum(dim=1) - hit[:, rows_j].sum(dim=1) + sc_b[4:6]

        if W > 1:
            L = packed.numel()
            gathered = torch.empty((W, L), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gathered.view(-1), packed, group=cfg.group)
            # rank-major [W][2][A][B] -> chunk-major [2][A][W][B]
            lse_all = gathered[:, :2 * pad].view(W, 2, pad)[:, :, :n_loc].reshape(W, 2, A, B)
            lse_all = lse_all.permute(1, 2, 0, 3).reshape(2, N)
            scalars = gathered[:, 2 * pad:].sum(dim=0)
        else:
            lse_all = lse_loc
            scalars = scal
        red = scalars * _scalar_coefs(N, dev)
        loss = red[0] + red[1]
        dscale = red[2] + red[3]
        acc_i2t, acc_t2i = red[4], red[5]

        ctx.save_for_backward(s_dev, lse_all.contiguous(), dscale)
        ctx.state, ctx.state_version, ctx.cfg, ctx.j = st, st.version, cfg, j
        ctx.dtypes = (chunk_img.dtype, chunk_txt.dtype)
        ctx.mark_non_differentiable(acc_i2t, acc_t2i)
        return loss, acc_i2t, acc_t2i

    @staticmethod
    def backward(ctx, g_loss, _g1, _g2):
        s_dev, lse_all, dscale = ctx.saved_tensors
        st, cfg, j = ctx.state, ctx.cfg, ctx.j
        if st.version != ctx.state_version:
            raise RuntimeError("incremental accumulate path: backward() of call j must run before get_loss of the "
                               "next chunk (train.py:241-247 does); set NANS_ACCUM_INCREMENTAL=0 otherwise")
        dt_i, dt_t = ctx.dtypes
        dI = dT = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            g = g_loss.detach().to(torch.float32).reshape(1).contiguous()
            out_dt = dt_i if dt_i == dt_t else torch.float32
            dI, dT = K.bwd(st.L_I, st.L_T, st.Wk_T, st.Wk_I, label_begin=st.label_begin(j), s_dev=s_dev,
                           lse_all=lse_all, grad_out=g, grad_mult=float(st.W) if cfg.gather_with_grad else 1.0,
                           row_begin=j * st.B, row_count=st.B, out_dtype=out_dt)
            dI = dI.to(dt_i) if ctx.needs_input_grad[0] else None
            dT = dT.to(dt_t) if ctx.needs_input_grad[1] else None
        ds = g_loss * dscale if ctx.needs_input_grad[2] else None
        return dI, dT, ds, None, None, None, None


def incremental_accum_loss(image_features, text_features, logit_scale, cache_img, cache_txt, accum_idx: int, *,
                           group=None, gather_with_grad: bool = False, report_acc: bool = False,
                           feat_dtype: torch.dtype = torch.float16):
    """Loss of call `accum_idx` of the accumulate path: (loss, acc) exactly as clip_contrastive_loss
    on the spliced block.  `cache_*` are the step's lists of cached (no-grad) chunk features,
    `image_features` / `text_features` the re-forwarded chunk `accum_idx` (with grad)."""
    cfg = LossConfig(group=group, gather_with_grad=gather_with_grad, report_acc=report_acc, feat_dtype=feat_dtype)
    loss, i2t, t2i = _AccumLossFn.apply(image_features, text_features, logit_scale, cache_img, cache_txt,
                                        int(accum_idx), cfg)
    return loss, ({"i2t": i2t, "t2i": t2i} if report_acc else None)
