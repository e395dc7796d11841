"""Drop-in for the tail of the reference's CLIP.forward and for get_similarity
(cn_clip/clip/model.py:402-431).

The encoders are the reference's own (out of scope); only the lines after them are replaced:

    image_features = image_features / image_features.norm(dim=-1, keepdim=True)     # :412
    text_features  = text_features  / text_features.norm(dim=-1, keepdim=True)      # :413
    return image_features, text_features, self.logit_scale.exp()                    # :415

`forward` / `get_similarity` below have the reference's signatures and can be bound onto a
reference CLIP instance with `patch_clip(model)`; `l2_normalize` is the differentiable kernel
call (one HBM pass forward, one backward).
"""
from __future__ import annotations

import types

import torch

from .. import kernels as K


class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x2 = x.reshape(-1, x.shape[-1])
        _, y32, inv = K.l2norm_cast(x2, None, normalize=True, want_fp32=True, want_inv_norm=True)
        ctx.save_for_backward(x2, inv)
        ctx.shape, ctx.dtype = x.shape, x.dtype
        # model.py:412 divides the tower output by an fp32-or-wider norm: the result has the
        # tower's dtype for fp32 towers and is promoted for 16-bit ones; fp32 covers both.
        return y32.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, inv = ctx.saved_tensors
        dx = K.l2norm_bwd(x2, inv, dy.reshape(x2.shape))
        return dx.reshape(ctx.shape).to(ctx.dtype)


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """x / x.norm(dim=-1, keepdim=True) on the sm_100a kernel (fp32 result)."""
    return _L2Normalize.apply(x)


def forward(self, image, text, mask_ratio=0):
    """cn_clip/clip/model.py:402-415 with the normalisation on the fused kernel."""
    assert image is not None or text is not None, "text and image cannot both be None!"
    if image is None:
        return self.encode_text(text)
    elif text is None:
        return self.encode_image(image)
    image_features = self.encode_image(image, mask_ratio)
    text_features = self.encode_text(text)
    return l2_normalize(image_features), l2_normalize(text_features), self.logit_scale.exp()


def get_similarity(self, image, text):
    """cn_clip/clip/model.py:417-431.  The API returns the materialised logits, so at this (small,
    latency-bound) call site the product itself stays a plain torch matmul; only the
    normalisation runs on the kernel.  Association as in the reference: (s * I) @ T^T."""
    image_features = l2_normalize(self.encode_image(image))
    text_features = l2_normalize(self.encode_text(text))
    logit_scale = self.logit_scale.exp()
    logits_per_image = logit_scale * image_features @ text_features.t()
    return logits_per_image, logits_per_image.t()


def patch_clip(model):
    """Bind the fused tail onto an existing reference `CLIP` instance (or DDP-wrapped one)."""
    target = model.module if hasattr(model, "module") else model
    target.forward = types.MethodType(forward, target)
    target.get_similarity = types.MethodType(get_similarity, target)
    return model
