"""Prompt front-end on the GPU: the left-padding of the T5 caption embeddings that `autoregressive/sample/sample_t2i.py:146-156` does
with a Python loop and one `.item()` per prompt — here one kernel, no host synchronisation."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, cur_stream, _ptr


def left_pad_captions(caption_embs: torch.Tensor, emb_masks: torch.Tensor):
    """caption_embs [B, L, D], emb_masks [B, L] (1 on the valid PREFIX, as the T5 tokenizer pads on the right) ->
    (new_caption_embs, new_emb_masks): valid tokens rotated to the end of each sequence, masks flipped."""
    if caption_embs.device.type != "cuda":
        raise RuntimeError("left_pad_captions needs CUDA tensors (controlar_b200 has no CPU path)")
    x = caption_embs.contiguous()
    m = emb_masks.to(torch.int64).contiguous()
    B, L, D = x.shape
    out = torch.empty_like(x)
    mo = torch.empty_like(m)
    with torch.cuda.device(x.device):
        check(_lib.lib().car_left_pad_captions(_ptr(x), _ptr(m), B, L, D * x.element_size(), _ptr(out), _ptr(mo), cur_stream()),
              "car_left_pad_captions")
    return out, mo.to(emb_masks.dtype)
