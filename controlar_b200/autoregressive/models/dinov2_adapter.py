"""Drop-in for the reference ``autoregressive/models/dinov2_adapter.py`` (control encoder).

The reference wraps HF ``transformers.AutoModel.from_pretrained('autoregressive/models/dinov2-{size}')``
(dinov2_adapter.py:13) — a third-party dependency (unpinned, requirements.txt:19; 5.5.0 installed here).  This
module owns parameters under the *same state-dict keys* (``model.embeddings.*``, ``model.encoder.layer.{i}.*``,
``model.layernorm.*``) and runs the forward (resize -> patch embed -> 12 pre-LN blocks -> LN -> drop CLS,
dinov2_adapter.py:16-29 + modeling_dinov2.py) in the library's kernels.  No HF model is instantiated.
"""
from __future__ import annotations

import json
import os

import torch
import torch.nn as nn

_SIZES = {"small": dict(hidden=384, heads=6), "base": dict(hidden=768, heads=12)}


class _Holder(nn.Module):
    pass


def _linear(i, o):
    return nn.Linear(i, o, bias=True)


class Dinov2Backbone(nn.Module):
    """Parameter container with HF Dinov2Model's key names (modeling_dinov2.py, transformers 5.5.0)."""

    def __init__(self, hidden: int, heads: int, layers: int = 12, mlp_ratio: int = 4, patch: int = 14, image: int = 518,
                 eps: float = 1e-6):
        super().__init__()
        self.hidden, self.heads, self.n_layers, self.patch, self.eps = hidden, heads, layers, patch, eps
        self.pos_grid = image // patch
        emb = _Holder()
        emb.cls_token = nn.Parameter(torch.randn(1, 1, hidden))
        emb.mask_token = nn.Parameter(torch.zeros(1, hidden))
        emb.position_embeddings = nn.Parameter(torch.randn(1, self.pos_grid ** 2 + 1, hidden))
        pe = _Holder()
        pe.projection = nn.Conv2d(3, hidden, kernel_size=patch, stride=patch)
        emb.patch_embeddings = pe
        self.embeddings = emb
        enc = _Holder()
        blocks = []
        for _ in range(layers):
            b = _Holder()
            b.norm1 = nn.LayerNorm(hidden, eps=eps)
            att = _Holder()
            inner = _Holder()
            inner.query, inner.key, inner.value = _linear(hidden, hidden), _linear(hidden, hidden), _linear(hidden, hidden)
            att.attention = inner
            outp = _Holder()
            outp.dense = _linear(hidden, hidden)
            att.output = outp
            b.attention = att
            ls1 = _Holder(); ls1.lambda1 = nn.Parameter(torch.ones(hidden)); b.layer_scale1 = ls1
            b.norm2 = nn.LayerNorm(hidden, eps=eps)
            mlp = _Holder()
            mlp.fc1, mlp.fc2 = _linear(hidden, hidden * mlp_ratio), _linear(hidden * mlp_ratio, hidden)
            b.mlp = mlp
            ls2 = _Holder(); ls2.lambda1 = nn.Parameter(torch.ones(hidden)); b.layer_scale2 = ls2
            blocks.append(b)
        enc.layer = nn.ModuleList(blocks)
        self.encoder = enc
        self.layernorm = nn.LayerNorm(hidden, eps=eps)


class Dinov2_Adapter(nn.Module):
    def __init__(self, input_dim=1, output_dim=768, attention=False, pool=False, nheads=8, dropout=0.1,
                 adapter_size="small", condition_type="canny"):
        super().__init__()
        sz = _SIZES[adapter_size]
        self.model = Dinov2Backbone(sz["hidden"], sz["heads"])
        self.condition_type = condition_type
        self.adapter_size = adapter_size
        self._load_pretrained_if_present(adapter_size)

    def _load_pretrained_if_present(self, adapter_size: str) -> None:
        """The reference loads 'autoregressive/models/dinov2-{size}' relative to CWD (dinov2_adapter.py:13).  If that
        directory holds a safetensors checkpoint we load it the same way; otherwise parameters stay random until
        the ControlAR checkpoint (which contains adapter.model.*) is loaded."""
        d = os.path.join("autoregressive", "models", f"dinov2-{adapter_size}")
        f = os.path.join(d, "model.safetensors")
        if os.path.isfile(f):
            from safetensors.torch import load_file
            self.model.load_state_dict(load_file(f), strict=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ... import vision as _vision
        return _vision.dinov2_forward(self, x)
