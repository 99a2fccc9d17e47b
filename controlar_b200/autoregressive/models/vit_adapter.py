"""Drop-in for the reference ``autoregressive/models/vit_adapter.py`` — the control encoder of the legacy c2i class
(``gpt.py``): HF ``ViTModel`` ViT-S/16 called with ``interpolate_pos_encoding=True``, CLS token dropped
(vit_adapter.py:13-15; third party: transformers, 5.5.0 restated in oracle/vision_oracle.py:vit_adapter_oracle).

Parameters live under the same state-dict keys as HF ViTModel (``model.embeddings.*``, ``model.encoder.layer.{i}.*``,
``model.layernorm.*``, ``model.pooler.dense.*``); the forward runs in the library's encoder kernels (car_dino_* with patch 16,
unit LayerScale, eps 1e-12, no input resize).  No HF model is instantiated.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn


class _Holder(nn.Module):
    pass


class ViTBackbone(nn.Module):
    """Parameter container with HF ViTModel's key names (modeling_vit.py, transformers 5.5.0)."""

    def __init__(self, hidden: int = 384, heads: int = 6, layers: int = 12, intermediate: int = 1536, patch: int = 16,
                 image: int = 224, eps: float = 1e-12):
        super().__init__()
        assert intermediate == 4 * hidden, "the encoder kernels assume an MLP ratio of 4"
        self.hidden, self.heads, self.n_layers, self.patch, self.eps = hidden, heads, layers, patch, eps
        self.pos_grid = image // patch
        emb = _Holder()
        emb.cls_token = nn.Parameter(torch.randn(1, 1, hidden))
        emb.position_embeddings = nn.Parameter(torch.randn(1, self.pos_grid ** 2 + 1, hidden))
        pe = _Holder()
        pe.projection = nn.Conv2d(3, hidden, kernel_size=patch, stride=patch)
        emb.patch_embeddings = pe
        self.embeddings = emb
        enc = _Holder()
        blocks = []
        for _ in range(layers):
            b = _Holder()
            att, inner, outp = _Holder(), _Holder(), _Holder()
            inner.query, inner.key, inner.value = nn.Linear(hidden, hidden), nn.Linear(hidden, hidden), nn.Linear(hidden, hidden)
            att.attention = inner
            outp.dense = nn.Linear(hidden, hidden)
            att.output = outp
            b.attention = att
            inter, out2 = _Holder(), _Holder()
            inter.dense = nn.Linear(hidden, intermediate)
            out2.dense = nn.Linear(intermediate, hidden)
            b.intermediate, b.output = inter, out2
            b.layernorm_before = nn.LayerNorm(hidden, eps=eps)
            b.layernorm_after = nn.LayerNorm(hidden, eps=eps)
            blocks.append(b)
        enc.layer = nn.ModuleList(blocks)
        self.encoder = enc
        self.layernorm = nn.LayerNorm(hidden, eps=eps)
        pool = _Holder()
        pool.dense = nn.Linear(hidden, hidden)          # present in the checkpoint (AutoModel adds the pooler); unused by forward
        self.pooler = pool


class ViT_Adapter(nn.Module):
    condition_type = "canny"          # (no resize branch in this adapter; the attribute only feeds the shared handle)

    def __init__(self, input_dim=3, output_dim=768, attention=False, pool=False, nheads=8, dropout=0.1, layers: int = 12):
        super().__init__()
        self.model = ViTBackbone(layers=layers)
        f = os.path.join("autoregressive", "models", "vit-small", "model.safetensors")     # vit_adapter.py:11, relative to CWD
        if os.path.isfile(f):
            from safetensors.torch import load_file
            self.model.load_state_dict(load_file(f), strict=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ... import vision as _vision
        return _vision.dinov2_forward(self, x)
