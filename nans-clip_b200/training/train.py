"""Drop-in for the loss call site of the reference: cn_clip/training/train.py:21-126.

`get_loss` keeps the reference's signature, the `args` fields it reads and its return value
`(total_loss, acc)`; `train()` (train.py:192-203, 241-247) can call it unchanged and then
`total_loss.backward()`.  What changed is below the call: the gather, the logits, both
cross-entropies, the accuracy and the whole backward run on the fused sm_100a kernels
(nans_clip_b200.loss) and the N x N logits are never materialised.

Kept quirks: `logit_scale.mean()` on the 0-d scale (train.py:52); the accumulate
path splices chunk j into the cached blocks (train.py:48-51) and only that chunk gets gradient;
`aggregate=False` means a local-batch loss with no collective (train.py:103-104).  The knowledge-
distillation branch (train.py:25-32, 37-46, 62-63, 90-100, 106-107, 123-124) is outside the fused
path: with `args.distillation` the cosine KD term is added with ordinary PyTorch ops on top of the
fused contrastive loss.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .. import accum
from ..loss import clip_contrastive_loss

# 16-bit operand type handed to the tensor cores.  fp16 keeps unit-norm features 8x more precisely
# than bf16 and is what meets the 1e-3 gradient tolerance (DESIGN.md "Precision").
FEAT_DTYPE = torch.float16


def is_master(args):
    return args.rank == 0


def _check_criterion(crit, name):
    """The fused kernels implement nn.CrossEntropyLoss() with its defaults (train.py:145-146), plus
    its `label_smoothing` (the variant of the fork's train_lora.py:105-108).  Returns the smoothing."""
    if crit is None:
        return 0.0
    if not isinstance(crit, nn.CrossEntropyLoss):
        raise NotImplementedError(f"{name}: the fused loss implements nn.CrossEntropyLoss only")
    if crit.weight is not None or crit.reduction != "mean" or crit.ignore_index != -100:
        raise NotImplementedError(f"{name}: only nn.CrossEntropyLoss with mean reduction and no class "
                                  "weights is supported")
    return float(getattr(crit, "label_smoothing", 0.0))


def cosineSimilarityLoss(feature1, feature2):
    """KD term of the reference (train.py:406-419): the student features are bilinearly resized
    to the teacher's [rows, width] and the loss is 1 - mean row-wise cosine similarity."""
    resized = F.interpolate(feature2[None, None], size=tuple(feature1.shape[:2]), mode="bilinear",
                            align_corners=False)[0, 0]
    return 1 - F.cosine_similarity(feature1, resized, dim=1).mean()


def _teacher_features(teacher_model, images):
    with torch.no_grad():
        output = teacher_model.module.get_feature(images)
        return output[0] if isinstance(output, tuple) else output


def get_loss(model, images, texts, loss_img, loss_txt, args, accum_image_features=None,
             accum_text_features=None, accum_idx=-1, teacher_model=None,
             teacher_accum_image_features=None):
    smoothing = _check_criterion(loss_img, "loss_img")
    if _check_criterion(loss_txt, "loss_txt") != smoothing:
        raise NotImplementedError("loss_img and loss_txt must use the same label_smoothing")
    teacher_image_features = None
    if args.accum_freq == 1:
        image_features, text_features, logit_scale = model(images, texts, args.mask_ratio)
        full_image_features, full_text_features, row_begin = image_features, text_features, 0
        if args.distillation:
            teacher_image_features = _teacher_features(teacher_model, images)
    else:
        assert accum_image_features and accum_text_features and accum_idx != -1
        image_features, text_features, logit_scale = model(images, texts, args.mask_ratio)
        if args.distillation:
            teacher_chunk = _teacher_features(teacher_model, images)
            teacher_image_features = torch.cat(teacher_accum_image_features[:accum_idx] + [teacher_chunk]
                                               + teacher_accum_image_features[accum_idx + 1:])
        row_begin = sum(int(f.shape[0]) for f in accum_image_features[:accum_idx])
    logit_scale = logit_scale.mean()

    group = dist.group.WORLD if (args.aggregate and dist.is_available() and dist.is_initialized()) else None
    if args.aggregate and group is None:
        raise RuntimeError("args.aggregate is set but torch.distributed is not initialised "
                           "(the reference calls dist.get_world_size() here, train.py:54)")
    world = dist.get_world_size(group) if group is not None else 1
    incremental = (args.accum_freq > 1 and image_features.is_cuda
                   and accum.eligible(accum_image_features, accum_text_features, int(image_features.shape[0]),
                                      world, smoothing))
    if args.accum_freq > 1 and (not incremental or args.distillation):
        # the block the loss sees: chunk j spliced between the cached (no-grad) chunks (train.py:48-51)
        full_image_features = torch.cat(accum_image_features[:accum_idx] + [image_features.detach()]
                                        + accum_image_features[accum_idx + 1:])
        full_text_features = torch.cat(accum_text_features[:accum_idx] + [text_features.detach()]
                                       + accum_text_features[accum_idx + 1:])
    if incremental:
        # only chunk j's rows and columns differ from the cached features: incremental forward (accum.py)
        total_loss, acc = accum.incremental_accum_loss(
            image_features, text_features, logit_scale, accum_image_features, accum_text_features, accum_idx,
            group=group, gather_with_grad=bool(args.aggregate and args.gather_with_grad),
            report_acc=bool(args.report_training_batch_acc), feat_dtype=FEAT_DTYPE)
    else:
        total_loss, acc = clip_contrastive_loss(
            image_features, text_features, logit_scale, group=group,
            gather_with_grad=bool(args.aggregate and args.gather_with_grad),
            report_acc=bool(args.report_training_batch_acc), feat_dtype=FEAT_DTYPE,
            full_image_features=full_image_features, full_text_features=full_text_features,
            row_begin=row_begin, label_smoothing=smoothing)

    if args.distillation:
        # outside the fused path: plain PyTorch, same gather order as train.py:90-100
        if group is not None:
            W, rank = dist.get_world_size(), dist.get_rank()
            gathered = [torch.zeros_like(teacher_image_features) for _ in range(W)]
            dist.all_gather(gathered, teacher_image_features)
            all_teacher = torch.cat([teacher_image_features] + gathered[:rank] + gathered[rank + 1:])
            spliced = full_image_features.clone()
            spliced[row_begin:row_begin + image_features.shape[0]] = image_features
            mine = [torch.zeros_like(spliced) for _ in range(W)]
            dist.all_gather(mine, spliced.detach())
            all_image = torch.cat([spliced] + mine[:rank] + mine[rank + 1:])
            kd_loss = cosineSimilarityLoss(all_teacher, all_image)
        else:
            spliced = full_image_features.clone()
            spliced[row_begin:row_begin + image_features.shape[0]] = image_features
            kd_loss = cosineSimilarityLoss(teacher_image_features, spliced)
        total_loss = total_loss + kd_loss * args.kd_loss_weight
    return total_loss, acc
