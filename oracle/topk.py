"""Oracle for the retrieval path (numpy / torch on CPU).  TEST INFRASTRUCTURE — see __init__.py."""
from __future__ import annotations

import numpy as np
import torch


def topk_literal(gallery_ids, gallery: np.ndarray, query: np.ndarray, k: int,
                 eval_batch_size: int = 32768):
    """cn_clip/eval/make_topk_predictions.py:74-85 for ONE query, statement by statement: chunked
    fp32 products (:78-80), python floats (:81-82), stable descending sort (:84), ids (:85)."""
    score_tuples = []
    q = torch.tensor([list(map(float, query))], dtype=torch.float)
    idx = 0
    while idx < len(gallery_ids):
        chunk = torch.from_numpy(gallery[idx: min(idx + eval_batch_size, len(gallery_ids))])
        batch_scores = q @ chunk.t()
        for gid, score in zip(gallery_ids[idx: min(idx + eval_batch_size, len(gallery_ids))],
                              batch_scores.squeeze(0).tolist()):
            score_tuples.append((gid, score))
        idx += eval_batch_size
    top = sorted(score_tuples, key=lambda x: x[1], reverse=True)[:k]
    return [e[0] for e in top], [e[1] for e in top]


def topk_vectorised(gallery: torch.Tensor, queries: torch.Tensor, k: int, block: int = 1024):
    """Same result as topk_literal for all queries at once: fp32 Q @ G^T, stable descending sort
    (ties keep ascending gallery position — what sorted(..., reverse=True) does, :84).
    Returns (scores [Q, k'], positions [Q, k']) with k' = min(k, G)."""
    G = gallery.shape[0]
    kk = min(k, G)
    out_s, out_i = [], []
    for b in range(0, queries.shape[0], block):
        sc = queries[b:b + block].float() @ gallery.float().t()
        s, i = torch.sort(sc, dim=1, descending=True, stable=True)
        out_s.append(s[:, :kk].clone())
        out_i.append(i[:, :kk].clone())
    if not out_s:
        return torch.empty(0, kk), torch.empty(0, kk, dtype=torch.int64)
    return torch.cat(out_s), torch.cat(out_i)


def excusable(ref_scores: torch.Tensor, gap: float = 1e-4) -> torch.Tensor:
    """Positions of a reference top-k list whose order is NOT pinned by the tolerance
    (BASELINE.json: indices identical wherever the score gap exceeds 1e-4): a position is pinned
    only if both neighbouring gaps exceed `gap`.  ref_scores must include one extra column (k+1)
    so that the last position's lower gap is known."""
    d = (ref_scores[:, :-1] - ref_scores[:, 1:]).abs() <= gap   # [Q, k]: gap below position j
    k = ref_scores.shape[1] - 1
    loose = torch.zeros(ref_scores.shape[0], k, dtype=torch.bool)
    loose |= d[:, :k]
    loose[:, 1:] |= d[:, :k - 1]
    return loose
