"""nans_clip_b200 — the B200 (sm_100a) global-batch contrastive loss and top-k retrieval path of
n571e/NanS-CLIP, behind the reference's own Python call sites.

  nans_clip_b200.training.train.get_loss              <- cn_clip/training/train.py:21-126
  nans_clip_b200.clip.model.forward / get_similarity  <- cn_clip/clip/model.py:402-431
  nans_clip_b200.eval.make_topk_predictions           <- cn_clip/eval/make_topk_predictions.py
  nans_clip_b200.loss.clip_contrastive_loss           (the functional core of get_loss)
  nans_clip_b200.kernels                              (torch wrappers of include/nans_clip.h)

The package directory is `nans-clip_b200/`; `nans_clip_b200/` at the repo root is the import alias.
"""
__version__ = "0.1.0"

__all__ = ["__version__"]
