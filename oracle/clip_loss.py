"""Oracle for the contrastive-loss path (torch on CPU, fp32 like the reference; fp64 on request).

TEST INFRASTRUCTURE — see oracle/__init__.py.  Every function cites the reference lines it restates
(paths relative to /root/reference).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def normalize(x: torch.Tensor) -> torch.Tensor:
    """cn_clip/clip/model.py:412-413 — `x / x.norm(dim=-1, keepdim=True)`."""
    return x / x.norm(dim=-1, keepdim=True)


def forward_tail(image_features, text_features, logit_scale_param):
    """cn_clip/clip/model.py:412-415 — normalise both towers' outputs, return exp(logit_scale)."""
    return normalize(image_features), normalize(text_features), logit_scale_param.exp()


def logits_pair(all_image_features, all_text_features, logit_scale):
    """cn_clip/training/train.py:87-88 — note Python precedence: (s * I) @ T^T; the text-side
    logits are the transposed VIEW of the image-side ones."""
    logits_per_image = logit_scale * all_image_features @ all_text_features.t()
    return logits_per_image, logits_per_image.t()


def loss_from_logits(logits_per_image, logits_per_text, report_acc: bool = False):
    """cn_clip/training/train.py:109-121 with the nn.CrossEntropyLoss() defaults of :145-146
    (mean reduction, no smoothing)."""
    gt = torch.arange(len(logits_per_image)).long()
    total = (F.cross_entropy(logits_per_image, gt) + F.cross_entropy(logits_per_text, gt)) / 2
    acc = None
    if report_acc:
        i2t = (logits_per_image.argmax(-1) == gt).sum() / len(logits_per_image)
        t2i = (logits_per_text.argmax(-1) == gt).sum() / len(logits_per_text)
        acc = {"i2t": i2t, "t2i": t2i}
    return total, acc


def local_loss(image_features, text_features, logit_scale, report_acc: bool = False):
    """cn_clip/training/train.py:103-104 + 109-121 — the `aggregate == False` branch: two
    independent products on the local batch."""
    lpi = logit_scale * image_features @ text_features.t()
    lpt = logit_scale * text_features @ image_features.t()
    return loss_from_logits(lpi, lpt, report_acc)


def gather_local_first(blocks: list[torch.Tensor], rank: int) -> torch.Tensor:
    """cn_clip/training/train.py:75-84 — the no-grad gather puts the local block (the only one
    that carries gradient) in slot 0 and the others after it in rank order."""
    return torch.cat([blocks[rank]] + blocks[:rank] + blocks[rank + 1:])


def rank_loss(image_blocks: list[torch.Tensor], text_blocks: list[torch.Tensor], logit_scale,
              rank: int, gather_with_grad: bool, report_acc: bool = False):
    """What ONE rank of a W-rank job computes (cn_clip/training/train.py:53-121), restated in a
    single process: `image_blocks[r]` is rank r's local features.

    gather_with_grad=False (train.py:65-84): the other ranks' blocks are constants, the local block
      sits in slot 0.
    gather_with_grad=True (train.py:59-60): torch.distributed.nn.all_gather — every block is
      differentiable on every rank and the backward SUMS the W ranks' (identical) gradients, so
      each local block receives W x d(loss)/d(block).

    Returns (loss, acc, dI_local, dT_local, d_logit_scale) with the gradients this rank's autograd
    would hold after loss.backward() (before DDP averages parameter gradients)."""
    W = len(image_blocks)
    s = logit_scale.detach().clone().requires_grad_(True)
    Il = image_blocks[rank].detach().clone().requires_grad_(True)
    Tl = text_blocks[rank].detach().clone().requires_grad_(True)
    if gather_with_grad:
        ib = [b.detach() for b in image_blocks]
        tb = [b.detach() for b in text_blocks]
        ib[rank], tb[rank] = Il, Tl
        all_i, all_t = torch.cat(ib), torch.cat(tb)
        mult = float(W)
    else:
        ib = [b.detach() for b in image_blocks]
        tb = [b.detach() for b in text_blocks]
        ib[rank], tb[rank] = Il, Tl
        all_i, all_t = gather_local_first(ib, rank), gather_local_first(tb, rank)
        mult = 1.0
    lpi, lpt = logits_pair(all_i, all_t, s)
    loss, acc = loss_from_logits(lpi, lpt, report_acc)
    loss.backward()
    return loss.detach(), acc, Il.grad * mult, Tl.grad * mult, s.grad


def accum_splice(cache: list[torch.Tensor], chunk: torch.Tensor, accum_idx: int) -> torch.Tensor:
    """cn_clip/training/train.py:48-51 — re-forwarded chunk j replaces its cached copy."""
    return torch.cat(cache[:accum_idx] + [chunk] + cache[accum_idx + 1:])


def global_loss_and_grads(I: torch.Tensor, T: torch.Tensor, s: float, dtype=torch.float32):
    """Single-process global loss on [N, D] features and its gradients (the W = 1 case of
    rank_loss, also the quantity every rank's loss equals)."""
    Ii = I.to(dtype).clone().requires_grad_(True)
    Ti = T.to(dtype).clone().requires_grad_(True)
    si = torch.tensor(float(s), dtype=dtype, requires_grad=True)
    lpi, lpt = logits_pair(Ii, Ti, si)
    loss, acc = loss_from_logits(lpi, lpt, True)
    loss.backward()
    return {"loss": loss.detach(), "i2t": acc["i2t"], "t2i": acc["t2i"], "dI": Ii.grad, "dT": Ti.grad,
            "ds": si.grad, "lse_img": torch.logsumexp(lpi.detach(), dim=1),
            "lse_txt": torch.logsumexp(lpt.detach(), dim=1)}


def lora_contrastive_loss(image_features, text_features, logit_scale, label_smoothing: float = 0.05):
    """train_lora.py:95-110 — the fork's LoRA script: F.normalize both feature sets, logits =
    logit_scale * I @ T^T, mean of F.cross_entropy(label_smoothing=eps) over both directions."""
    I = F.normalize(image_features, dim=-1)
    T = F.normalize(text_features, dim=-1)
    logits = logit_scale * I @ T.T
    labels = torch.arange(logits.shape[0])
    return (F.cross_entropy(logits, labels, label_smoothing=label_smoothing)
            + F.cross_entropy(logits.T, labels, label_smoothing=label_smoothing)) / 2


def global_smoothed_loss_and_grads(I: torch.Tensor, T: torch.Tensor, s: float, eps: float, dtype=torch.float32):
    """Label-smoothed loss of already-normalised [N, D] features and its gradients: the CE of
    cn_clip/training/train.py:109-115 with the label_smoothing of train_lora.py:105-108."""
    Ii = I.to(dtype).clone().requires_grad_(True)
    Ti = T.to(dtype).clone().requires_grad_(True)
    si = torch.tensor(float(s), dtype=dtype, requires_grad=True)
    lpi, lpt = logits_pair(Ii, Ti, si)
    gt = torch.arange(len(lpi))
    loss = (F.cross_entropy(lpi, gt, label_smoothing=eps) + F.cross_entropy(lpt, gt, label_smoothing=eps)) / 2
    loss.backward()
    return {"loss": loss.detach(), "dI": Ii.grad, "dT": Ti.grad, "ds": si.grad}
