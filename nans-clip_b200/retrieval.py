"""Top-k inner-product retrieval on the sm_100a kernels (host orchestration).

Functional core of cn_clip/eval/make_topk_predictions.py:57-85 (and _tr.py): for every query the k
gallery rows with the largest fp32 inner product — no renormalisation (the reference does none,
SURVEY.md §3.4) — ordered by (score descending, gallery position ascending), which is what the
reference's stable `sorted(..., reverse=True)[:k]` yields.

With a process group the GALLERY is sharded by contiguous row ranges (rank r holds rows
[offset_r, offset_r + G_r)), the queries are replicated, every rank runs the two-pass kernel on its
shard and the per-shard lists are all-gathered and merged with the same ordering rule
(nans_topk_merge); contiguous sharding makes the tie rule a plain index compare.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import kernels as K

MAX_K = 32          # what ONE kernel pass returns (its per-thread candidate list holds at most 32 entries)
SAFE_PER_PART = 24  # k > 32: entries a gallery partition may contribute before its 32-entry list counts as truncated


def _merge_candidates(scores: torch.Tensor, index: torch.Tensor, k: int):
    """[Q, C] candidate (score, global index) pairs in any order (index -1 = padding) -> the first k in
    the reference's order: score descending, ties by ascending gallery position (the stable sorted(...,
    reverse=True) of make_topk_predictions.py:84).  Small host-orchestrated merge (two stable sorts)."""
    big = torch.iinfo(torch.int64).max
    key_i = torch.where(index < 0, torch.full_like(index, big), index)
    order = torch.sort(key_i, dim=1, stable=True).indices                       # index ascending, padding last
    sc = torch.gather(torch.where(index < 0, torch.full_like(scores, -float("inf")), scores), 1, order)
    ix = torch.gather(index, 1, order)
    order = torch.sort(sc, dim=1, descending=True, stable=True).indices         # score descending, stable
    sc, ix = torch.gather(sc, 1, order), torch.gather(ix, 1, order)
    if sc.shape[1] < k:
        pad = k - sc.shape[1]
        sc = torch.cat([sc, sc.new_full((sc.shape[0], pad), -float("inf"))], 1)
        ix = torch.cat([ix, ix.new_full((ix.shape[0], pad), -1)], 1)
    return sc[:, :k].contiguous(), ix[:, :k].contiguous()


def _pick_k_cand(k: int, k_cand: Optional[int]) -> int:
    if k < 1 or k > MAX_K:
        raise ValueError(f"one kernel pass returns at most {MAX_K} neighbours (got k={k})")
    if k_cand is None:
        k_cand = 16 if k <= 10 else 32
    if k_cand not in (16, 32) or k_cand < k:
        raise ValueError("k_cand must be 16 or 32 and >= k")
    return k_cand


class GalleryShard:
    """A gallery shard resident in HBM: the fp32 rows (exact rescoring) and their 16-bit copy
    (tensor-core candidate pass).  Build once, query many times."""

    def __init__(self, gallery: torch.Tensor, device=None, feat_dtype: torch.dtype = torch.float16,
                 index_offset: int = 0, gallery16: Optional[torch.Tensor] = None):
        """`gallery16`: an existing 16-bit copy of `gallery` in `feat_dtype` (e.g. from a binary
        feature shard, eval/feature_io.py); made by kernel (1) when absent."""
        device = torch.device(device) if device is not None else (
            gallery.device if gallery.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        self.g32 = gallery.to(device=device, dtype=torch.float32, non_blocking=True).contiguous()
        self.feat_dtype = feat_dtype
        self.index_offset = int(index_offset)
        if gallery16 is not None:
            if gallery16.dtype != feat_dtype or gallery16.shape != gallery.shape:
                raise ValueError("gallery16 must have the gallery's shape and dtype feat_dtype")
            self.g16 = gallery16.to(device=device, non_blocking=True).contiguous()
        elif self.g32.shape[0] > 0:
            self.g16, _, _ = K.l2norm_cast(self.g32, feat_dtype, normalize=False)
        else:
            self.g16 = torch.empty_like(self.g32, dtype=feat_dtype)

    @property
    def rows(self) -> int:
        return self.g32.shape[0]

    def _search_large_k(self, q16: torch.Tensor, q32: torch.Tensor, k: int, parts: int):
        """k > 32 (the reference accepts any --top-k, make_topk_predictions.py:29-33): the shard is cut into
        `parts` contiguous partitions, each gives its exact top-32 with the ordinary kernel, and the lists are
        merged.  A partition whose list supplies >= SAFE_PER_PART of a query's k results may have been
        truncated (its 33rd row could belong too; beyond that margin the 16-bit candidate order is not
        pinned either): those queries are searched again with 4x the partitions, until no list is near
        full or a partition holds <= 32 rows (then nothing can be missing).  Same sweep flops, more launches."""
        G = self.rows
        parts = max(1, min(parts, (G + MAX_K - 1) // MAX_K))
        bounds = [G * p // parts for p in range(parts + 1)]
        ss, ii = [], []
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            kk = min(MAX_K, hi - lo)
            s, i = K.topk_ip(q16, self.g16[lo:hi], q32, self.g32[lo:hi], kk, 32 if kk > 16 else 16, self.index_offset + lo)
            ss.append(s)
            ii.append(i)
        cand_s, cand_i = torch.cat(ss, 1), torch.cat(ii, 1)
        out_s, out_i = _merge_candidates(cand_s, cand_i, k)
        if parts * MAX_K >= G:          # every partition returned all of its rows
            return out_s, out_i
        # how many of each query's results came from each partition
        edges = torch.tensor(bounds[1:-1], device=out_i.device, dtype=torch.int64) + self.index_offset
        part_of = torch.bucketize(out_i.clamp_min(self.index_offset), edges, right=True)
        counts = torch.zeros((out_i.shape[0], parts), dtype=torch.int32, device=out_i.device)
        counts.scatter_add_(1, part_of, (out_i >= 0).to(torch.int32))
        redo = (counts >= SAFE_PER_PART).any(dim=1).nonzero().flatten()
        if redo.numel():
            rs, ri = self._search_large_k(q16[redo].contiguous(), q32[redo].contiguous(), k, parts * 4)
            out_s[redo], out_i[redo] = rs, ri
        return out_s, out_i

    def search(self, queries: torch.Tensor, k: int = 10, k_cand: Optional[int] = None,
               query_block: int = 32768):
        """queries [Q, D] (any float dtype, host or device) -> (scores [Q, k], index [Q, k])."""
        if k > MAX_K:
            dev = self.g32.device
            out_s, out_i = [], []
            for b in range(0, max(queries.shape[0], 1), query_block):
                q32 = queries[b:b + query_block].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
                if q32.shape[0] == 0:
                    break
                q16, _, _ = K.l2norm_cast(q32, self.feat_dtype, normalize=False)
                s, i = self._search_large_k(q16, q32, k, max(2, (k + 7) // 8))
                out_s.append(s)
                out_i.append(i)
            if not out_s:
                return (torch.empty((0, k), dtype=torch.float32, device=dev),
                        torch.empty((0, k), dtype=torch.int64, device=dev))
            return torch.cat(out_s), torch.cat(out_i)
        k_cand = _pick_k_cand(k, k_cand)
        dev = self.g32.device
        out_s, out_i = [], []
        for b in range(0, max(queries.shape[0], 1), query_block):
            q32 = queries[b:b + query_block].to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            if q32.shape[0] == 0:
                break
            q16, _, _ = K.l2norm_cast(q32, self.feat_dtype, normalize=False)
            s, i = K.topk_ip(q16, self.g16, q32, self.g32, k, k_cand, self.index_offset)
            out_s.append(s)
            out_i.append(i)
        if not out_s:
            return (torch.empty((0, k), dtype=torch.float32, device=dev),
                    torch.empty((0, k), dtype=torch.int64, device=dev))
        return torch.cat(out_s), torch.cat(out_i)


def topk_retrieve(queries: torch.Tensor, gallery_shard: torch.Tensor, k: int = 10, *,
                  k_cand: Optional[int] = None, group=None, index_offset: Optional[int] = None,
                  feat_dtype: torch.dtype = torch.float16, device=None,
                  gallery16: Optional[torch.Tensor] = None):
    """Global top-k of `queries` against the gallery whose local shard is `gallery_shard`.

    Without `group`: single GPU, `index_offset` (default 0) is added to the returned positions.
    With `group`: rank r passes its contiguous shard; offsets default to the exclusive prefix sum
    of the shard sizes in rank order.  Every rank returns the merged global result."""
    if group is None:
        shard = GalleryShard(gallery_shard, device, feat_dtype, index_offset or 0, gallery16=gallery16)
        return shard.search(queries, k, k_cand)
    W, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if index_offset is None:
        sizes = torch.zeros(W, dtype=torch.int64, device=dev)
        sizes[rank] = gallery_shard.shape[0]
        dist.all_reduce(sizes, group=group)
        index_offset = int(sizes[:rank].sum().item())
    shard = GalleryShard(gallery_shard, dev, feat_dtype, index_offset, gallery16=gallery16)
    s, i = shard.search(queries, k, k_cand)
    Qn = s.shape[0]
    all_s = torch.empty((W * Qn, k), dtype=s.dtype, device=dev)
    all_i = torch.empty((W * Qn, k), dtype=i.dtype, device=dev)
    dist.all_gather_into_tensor(all_s, s.contiguous(), group=group)
    dist.all_gather_into_tensor(all_i, i.contiguous(), group=group)
    if W * k > 512:   # beyond the merge kernel's candidate buffer (large k): merge with two stable sorts
        return _merge_candidates(all_s.view(W, Qn, k).permute(1, 0, 2).reshape(Qn, W * k),
                                 all_i.view(W, Qn, k).permute(1, 0, 2).reshape(Qn, W * k), k)
    return K.topk_merge(all_s.view(W, Qn, k), all_i.view(W, Qn, k))
