# -*- coding: utf-8 -*-
"""Drop-in for cn_clip/eval/make_topk_predictions_tr.py: the image-to-text twin — image features
are the queries, text features the gallery; output `{"image_id": int, "text_ids": [k ints]}`
(make_topk_predictions_tr.py:85).  Same kernel, roles swapped."""
from __future__ import annotations

from .make_topk_predictions import parse_args, run


def main(argv=None):
    args = parse_args(argv)
    print("Params:")
    for name in sorted(vars(args)):
        print(f"  {name}: {getattr(args, name)}")
    run(args, query_key="image_id", query_path=args.image_feats, gallery_key="text_id",
        gallery_path=args.text_feats, out_query_key="image_id", out_list_key="text_ids")


if __name__ == "__main__":
    main()
