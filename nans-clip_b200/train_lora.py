"""Drop-in for the loss of the fork's LoRA fine-tuning script (train_lora.py:95-110).

    from nans_clip_b200.train_lora import contrastive_loss      # same signature, same value

The reference normalises the features (F.normalize), forms `logit_scale * I @ T.T` and averages
F.cross_entropy(..., label_smoothing=eps) over both directions.  Here the normalisation is kernel (1)
with its backward, the loss is the fused forward/backward plus the O(N D) smoothing terms
(csrc/smooth.cu); the N x N logits never exist.  Only the loss function is mirrored: the script's
LMDB dataset, LoRA wrapping and training loop are outside the hot path.
"""
from __future__ import annotations

import torch

from .clip.model import l2_normalize
from .loss import clip_contrastive_loss
from .training.train import FEAT_DTYPE


def contrastive_loss(image_features: torch.Tensor, text_features: torch.Tensor, logit_scale,
                     label_smoothing: float = 0.05) -> torch.Tensor:
    """InfoNCE with label smoothing; `logit_scale` is the multiplier itself (train_lora.py:100),
    a tensor (with or without grad) or a Python number."""
    if not torch.is_tensor(logit_scale):
        logit_scale = torch.tensor(float(logit_scale), device=image_features.device)
    loss, _ = clip_contrastive_loss(l2_normalize(image_features), l2_normalize(text_features), logit_scale,
                                    feat_dtype=FEAT_DTYPE, label_smoothing=label_smoothing)
    return loss
