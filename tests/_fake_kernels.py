"""CPU stand-ins for nans_clip_b200.kernels, used ONLY by the gloo host-logic tests.

They follow the kernels' contract (phases, slots, partial statistics, finalize scalars, strip
backward) in plain fp32 torch so that the multi-rank orchestration in loss.py / retrieval.py —
label offsets, column phases, the scalar all-reduce, the lse all-gather, the xW convention — can
be exercised under torch.distributed gloo without a GPU.  Test infrastructure, not a fallback.
"""
import math

import torch

STATE = {}


def install(K):
    for name in ("l2norm_cast", "fwd_phase_slots", "fwd_workspace", "fwd_phase", "fwd_finalize", "bwd",
                 "smooth_stats", "smooth_bwd", "clone_workspace", "exchange_finish", "topk_ip", "topk_merge"):
        setattr(K, name, globals()[name])


def l2norm_cast(x, out_dtype=torch.float16, *, normalize=True, want_fp32=False, want_inv_norm=False):
    xf = x.float()
    nrm = xf.norm(dim=-1, keepdim=True)
    y = xf / nrm if normalize else xf
    return (y.to(out_dtype) if out_dtype is not None else None, y if want_fp32 else None,
            (1 / nrm).squeeze(-1) if want_inv_norm else None)


def fwd_phase_slots(n_loc, ncols, D, strip=None):
    base = 2 if ncols > 30 else 1  # exercise multi-slot merging
    return base if strip is None else base + 1  # single-strip launches split differently


def fwd_workspace(n_loc, total_slots, device):
    t = torch.zeros(16, dtype=torch.uint8)
    STATE[t.data_ptr()] = {"slots": {}, "diag": torch.zeros(2, n_loc), "n": total_slots}
    return t


def clone_workspace(ws):
    import copy
    t = torch.zeros(16, dtype=torch.uint8)
    src = STATE[ws.data_ptr()]
    STATE[t.data_ptr()] = {"slots": {k: dict(v) for k, v in src["slots"].items()}, "diag": src["diag"].clone(),
                           "n": src["n"]}
    return t


def fwd_phase(I_loc, T_loc, T_cols, I_cols, *, col_global_begin, label_begin, s_dev, with_acc, ws, slot_begin,
              skip_begin=0, skip_count=0, strip=None, label_rows=None):
    st = STATE[ws.data_ptr()]
    s = float(s_dev)
    n_loc, ncols = I_loc.shape[0], T_cols.shape[0]
    assert skip_begin % 256 == 0 and skip_count % 256 == 0
    keep = torch.cat([torch.arange(0, skip_begin), torch.arange(skip_begin + skip_count, ncols)])
    ns = fwd_phase_slots(n_loc, keep.numel(), I_loc.shape[1], strip)
    bounds = [keep.numel() * i // ns for i in range(ns + 1)]
    strips = {None: (0, 1), "img": (0,), "txt": (1,)}[strip]
    for sl in range(ns):
        cols = keep[bounds[sl]:bounds[sl + 1]]
        entry = st["slots"].setdefault(slot_begin + sl, {})
        for k in strips:
            A, B = ((I_loc, T_cols), (T_loc, I_cols))[k]
            # (a slot may be overwritten on purpose: the incremental accumulate path replaces one
            #  column chunk's slots in a cloned workspace)
            cos = A.float() @ B[cols].float().t()
            t = cos * (s * math.log2(math.e))
            m = t.max(dim=1).values
            e = torch.exp2(t - m[:, None])
            bv, bj = cos.max(dim=1)
            entry[k] = (m, e.sum(1), (e * cos).sum(1), bv, cols[bj] + col_global_begin)
            lab = torch.arange(n_loc) + label_begin - col_global_begin
            if label_rows is not None:
                in_rows = (torch.arange(n_loc) >= label_rows[0]) & (torch.arange(n_loc) < label_rows[0] + label_rows[1])
                lab = torch.where(in_rows, lab, torch.full_like(lab, -1))
            pos = torch.full((ncols + 1,), -1, dtype=torch.long)
            pos[cols] = torch.arange(cols.numel())
            where = pos[lab.clamp(0, ncols)]
            where[(lab < 0) | (lab >= ncols)] = -1
            rows = torch.nonzero(where >= 0).squeeze(1)
            st["diag"][k, rows] = cos[rows, where[rows]]


def smooth_stats(I32, T32):
    return torch.cat([I32.sum(0), T32.sum(0), (I32 * T32).sum().reshape(1)])


def smooth_bwd(dI, dT, I_rows, T_rows, stats, s_dev, grad_out, coef, inv_n):
    D = dI.shape[1]
    a = float(grad_out) * float(s_dev) * coef
    dI += a * T_rows - a * inv_n * stats[D:2 * D]
    dT += a * I_rows - a * inv_n * stats[:D]


def fwd_finalize(n_loc, total_slots, label_begin, s_dev, with_acc, ws, want_row_stats=False):
    st = STATE.pop(ws.data_ptr())
    rows = torch.zeros(3, 2, n_loc)
    assert sorted(st["slots"]) == list(range(total_slots)), "a slot was skipped"
    assert all(sorted(e) == [0, 1] for e in st["slots"].values()), "a slot half was left unwritten"
    s = float(s_dev)
    lse = torch.empty(2, n_loc)
    sc = torch.zeros(8)
    for strip in range(2):
        parts = [st["slots"][k][strip] for k in range(total_slots)]
        M = torch.stack([p[0] for p in parts]).max(0).values
        L = sum(p[1] * torch.exp2(p[0] - M) for p in parts)
        Wt = sum(p[2] * torch.exp2(p[0] - M) for p in parts)
        bv = torch.stack([p[3] for p in parts])
        bi = torch.stack([p[4] for p in parts])
        best = bv.max(0).values
        cand = torch.where(bv == best[None], bi, torch.full_like(bi, 1 << 60))
        arg = cand.min(0).values
        lse[strip] = M + torch.log2(L)  # base-2 domain, as the kernel returns it
        d = st["diag"][strip]
        sc[strip] = (lse[strip].double() * math.log(2) - s * d.double()).sum()
        sc[2 + strip] = ((Wt / L).double() - d.double()).sum()
        if with_acc:
            sc[4 + strip] = (arg == torch.arange(n_loc) + label_begin).sum()
        rows[0, strip] = (lse[strip].double() * math.log(2) - s * d.double()).float()
        rows[1, strip] = ((Wt / L).double() - d.double()).float()
        rows[2, strip] = (arg if with_acc else torch.full_like(arg, -1)).to(torch.int32).view(torch.float32)
    packed = torch.cat([lse.reshape(-1), sc])
    return (lse, sc, packed, rows) if want_row_stats else (lse, sc, packed)


def exchange_finish(gathered, n_loc):
    W, L = gathered.shape
    pad = (L - 8) // 2
    N = W * n_loc
    lse_all = gathered[:, :2 * pad].view(W, 2, pad)[:, :, :n_loc].permute(1, 0, 2).reshape(2, N).clone()
    sc = gathered[:, 2 * pad:].sum(0)
    out = torch.stack([(sc[0] + sc[1]) / (2 * N), (sc[2] + sc[3]) / (2 * N), sc[4] / N, sc[5] / N])
    return lse_all, out, torch.zeros(2, dtype=torch.int32)


def bwd(I_loc, T_loc, T_all, I_all, *, label_begin, s_dev, lse_all, grad_out, grad_mult, row_begin,
        row_count, out_dtype, lse_minmax=None):
    s = float(s_dev)
    N = T_all.shape[0]
    rows = slice(row_begin, row_begin + row_count)
    outs = []
    for strip, (A, B) in enumerate(((I_loc, T_all), (T_loc, I_all))):
        lr = lse_all[strip][label_begin + row_begin: label_begin + row_begin + row_count] * math.log(2)
        lc = lse_all[1 - strip] * math.log(2)
        S = s * A[rows].float() @ B.float().t()
        G = torch.exp(S - lr[:, None]) + torch.exp(S - lc[None, :])
        idx = torch.arange(row_count)
        G[idx, idx + label_begin + row_begin] -= 2
        outs.append((float(grad_out) * s * grad_mult / (2 * N) * (G @ B.float())).to(out_dtype))
    return outs[0], outs[1]


def topk_ip(Q16, G16, Q32, G32, k, k_cand, gallery_index_offset=0):
    sc = Q32.float() @ G32.float().t()
    kk = min(k, G32.shape[0])
    s, i = torch.sort(sc, dim=1, descending=True, stable=True)
    out_s = torch.full((Q32.shape[0], k), -math.inf)
    out_i = torch.full((Q32.shape[0], k), -1, dtype=torch.int64)
    out_s[:, :kk] = s[:, :kk]
    out_i[:, :kk] = i[:, :kk] + gallery_index_offset
    return out_s, out_i


def topk_merge(scores, index):
    W, Q, k = scores.shape
    s = scores.permute(1, 0, 2).reshape(Q, W * k)
    i = index.permute(1, 0, 2).reshape(Q, W * k)
    key_i = torch.where(i < 0, torch.full_like(i, 1 << 60), i)
    order = torch.argsort(key_i, dim=1, stable=True)
    s, i = torch.gather(s, 1, order), torch.gather(i, 1, order)
    order = torch.argsort(s, dim=1, descending=True, stable=True)
    return torch.gather(s, 1, order)[:, :k], torch.gather(i, 1, order)[:, :k]
