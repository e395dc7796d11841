# -*- coding: utf-8 -*-
"""Drop-in for cn_clip/eval/make_topk_predictions.py: kNN search of text features against image
features, writing the text-to-image prediction file for evaluation.py.

Same command line (`--image-feats --text-feats --top-k --eval-batch-size --output`), same JSONL
input (`{"image_id": int, "feature": [...]}` / `{"text_id": int, "feature": [...]}`) and output
(`{"text_id": int, "image_ids": [k ints]}`, make_topk_predictions.py:85), same ordering of ties.
The per-query GEMV + Python sort of the reference (:71-85) is replaced by the fused tensor-core
top-k kernel; `--eval-batch-size` is kept for compatibility and bounds the QUERY block here (the
gallery is resident in HBM once instead of being re-uploaded per query).

Either feature file may also be a binary shard written by `feature_io`: no JSON
parsing, memory-mapped, and a stored 16-bit copy of the matching type is used as the operand as is.

Launched under torchrun (WORLD_SIZE > 1) the gallery is sharded over the ranks and rank 0 writes
the output.  Extra, optional flags: --feat-dtype {fp16,bf16}, --k-cand {16,32}.
"""
from __future__ import annotations

import argparse
import json
import os

import numpy as np
import torch

QUERY_KEY, QUERY_FEATS_ARG = "text_id", "text_feats"
GALLERY_KEY, GALLERY_FEATS_ARG = "image_id", "image_feats"
OUT_QUERY_KEY, OUT_LIST_KEY = "text_id", "image_ids"
QUERY_NAME, GALLERY_NAME = "texts", "image"


def parse_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--image-feats", type=str, required=True, help="Specify the path of image features.")
    parser.add_argument("--text-feats", type=str, required=True, help="Specify the path of text features.")
    parser.add_argument("--top-k", type=int, default=10, help="Specify the k value of top-k predictions.")
    parser.add_argument("--eval-batch-size", type=int, default=32768,
                        help="Query-side block size of the fused kernel (the reference used it as the "
                             "gallery-side batch of its per-query products).")
    parser.add_argument("--output", type=str, required=True, help="Specify the output jsonl prediction filepath.")
    parser.add_argument("--feat-dtype", choices=["fp16", "bf16"], default="fp16",
                        help="16-bit operand type of the candidate pass (scores are re-computed in fp32).")
    parser.add_argument("--k-cand", type=int, default=None, choices=[16, 32],
                        help="Candidates kept per query before exact rescoring.")
    return parser.parse_args(argv)


from .feature_io import DT16_BF16, DT16_F16, load_features, load_jsonl_features  # noqa: F401  (re-export)


def run(args, query_key=QUERY_KEY, query_path=None, gallery_key=GALLERY_KEY, gallery_path=None,
        out_query_key=OUT_QUERY_KEY, out_list_key=OUT_LIST_KEY, log=print):
    from ..retrieval import topk_retrieve  # imports the CUDA library: fails loudly without it
    import torch.distributed as dist

    query_path = query_path or args.text_feats
    gallery_path = gallery_path or args.image_feats
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("make_topk_predictions: a CUDA device (sm_100) is required; there is no CPU path")
    torch.cuda.set_device(local_rank)
    group = None
    if world > 1:
        if not dist.is_initialized():
            dist.init_process_group("nccl")
        group = dist.group.WORLD

    log(f"Begin to load {GALLERY_NAME if gallery_key == GALLERY_KEY else 'gallery'} features...")
    # every rank reads the ids of the whole gallery but only ITS contiguous share of the feature rows
    gallery_ids, gallery, gallery16, g16_code = load_features(gallery_path, gallery_key, shard=(rank, world))
    log("Finished loading features.")
    G = len(gallery_ids)
    lo, hi = (G * rank) // world, (G * (rank + 1)) // world
    query_ids, queries, _, _ = load_features(query_path, query_key)
    queries = np.array(queries, dtype=np.float32)
    k = args.top_k   # any k, like the reference (above 32: partitioned search + merge, retrieval.py)
    feat_dtype = torch.float16 if args.feat_dtype == "fp16" else torch.bfloat16

    log(f"Begin to compute top-{args.top_k} predictions...")
    positions = []
    qb = max(int(args.eval_batch_size), 1)
    shard = torch.from_numpy(np.array(gallery, dtype=np.float32))  # a writable copy of the (mmap) slice
    assert shard.shape[0] == hi - lo
    shard16 = None
    if gallery16 is not None and g16_code == (DT16_F16 if feat_dtype == torch.float16 else DT16_BF16):
        shard16 = torch.from_numpy(np.array(gallery16).view(np.int16)).view(feat_dtype)
    if queries.shape[0] > 0 and G > 0:
        if group is None:
            from ..retrieval import GalleryShard
            gs = GalleryShard(shard, None, feat_dtype, lo, gallery16=shard16)
            _, idx = gs.search(torch.from_numpy(queries), k, args.k_cand, query_block=qb)
        else:
            _, idx = topk_retrieve(torch.from_numpy(queries), shard, k, k_cand=args.k_cand, group=group,
                                   index_offset=lo, feat_dtype=feat_dtype, gallery16=shard16)
        positions = idx.cpu().numpy()
    if rank == 0:
        with open(args.output, "w") as fout:
            for qi, qid in enumerate(query_ids):
                row = positions[qi] if len(positions) else []
                ids = [int(gallery_ids[int(p)]) for p in row if int(p) >= 0]
                qid = int(qid) if isinstance(qid, np.integer) else qid
                fout.write("{}\n".format(json.dumps({out_query_key: qid, out_list_key: ids})))
        log("Top-{} predictions are saved in {}".format(args.top_k, args.output))
    if group is not None:
        dist.barrier()
    log("Done!")


def main(argv=None):
    args = parse_args(argv)
    print("Params:")
    for name in sorted(vars(args)):
        print(f"  {name}: {getattr(args, name)}")
    run(args)


if __name__ == "__main__":
    main()
