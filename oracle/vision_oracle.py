"""TEST INFRASTRUCTURE — CPU restatements of the control encoder and the image tokenizer.

* ``dinov2_adapter_oracle`` restates Dinov2_Adapter.forward (autoregressive/models/dinov2_adapter.py:16-29) on top of
  HF ``Dinov2Model`` (third-party: transformers, unpinned in requirements.txt:19; restated from the installed
  5.5.0 ``modeling_dinov2.py``: Dinov2Embeddings.forward/interpolate_pos_encoding, Dinov2Layer.forward,
  Dinov2SelfAttention (sdpa), Dinov2MLP, Dinov2LayerScale).
* ``vq_decode_oracle`` / ``vq_encode_oracle`` restate VQModel.decode_code / encode
  (tokenizer/tokenizer_image/vq_model.py:41-56,129-195,65-125,198-277,280-397).
* ``vit_adapter_oracle`` restates ViT_Adapter.forward (autoregressive/models/vit_adapter.py:13-15, HF ViTModel) — the
  control encoder of the legacy c2i class ``gpt.py`` (oracle only, no CUDA path yet).
Pinned by tests/golden/dinov2.pt, tests/golden/vit.pt and tests/golden/vq16.pt, which were produced by the reference modules.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def dinov2_adapter_oracle(sd: Dict[str, torch.Tensor], x: torch.Tensor, condition_type: str, dtype, heads: int,
                          layers: int = 12, eps: float = 1e-6, prefix: str = "model.") -> torch.Tensor:
    w = {k[len(prefix):]: v.to(dtype) for k, v in sd.items() if k.startswith(prefix)}
    x = x.to(dtype)
    B, _, H, W = x.shape
    nh, nw = (H // 16) * 14, (W // 16) * 14
    if condition_type in ("canny", "seg"):
        x = F.interpolate(x, size=(nh, nw), mode="nearest")
    else:
        x = F.interpolate(x, size=(nh, nw), mode="bicubic", align_corners=True)
    e = F.conv2d(x, w["embeddings.patch_embeddings.projection.weight"], w["embeddings.patch_embeddings.projection.bias"],
                 stride=14).flatten(2).transpose(1, 2)
    C = e.shape[-1]
    e = torch.cat([w["embeddings.cls_token"].expand(B, -1, -1), e], dim=1)
    pos = w["embeddings.position_embeddings"]
    G = int(round((pos.shape[1] - 1) ** 0.5))
    h, wd = nh // 14, nw // 14
    if not (h * wd == G * G and h == wd):
        pp = pos[:, 1:].reshape(1, G, G, C).permute(0, 3, 1, 2)
        pp = F.interpolate(pp.float(), size=(h, wd), mode="bicubic", align_corners=False).to(dtype)
        pos = torch.cat([pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, C)], dim=1)
    hs = e + pos
    for l in range(layers):
        p = f"encoder.layer.{l}."
        y = F.layer_norm(hs, (C,), w[p + "norm1.weight"], w[p + "norm1.bias"], eps)
        q = F.linear(y, w[p + "attention.attention.query.weight"], w[p + "attention.attention.query.bias"])
        k = F.linear(y, w[p + "attention.attention.key.weight"], w[p + "attention.attention.key.bias"])
        v = F.linear(y, w[p + "attention.attention.value.weight"], w[p + "attention.attention.value.bias"])
        sh = lambda t: t.view(B, -1, heads, C // heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sh(q), sh(k), sh(v), scale=(C // heads) ** -0.5)
        a = a.transpose(1, 2).reshape(B, -1, C)
        a = F.linear(a, w[p + "attention.output.dense.weight"], w[p + "attention.output.dense.bias"])
        hs = a * w[p + "layer_scale1.lambda1"] + hs
        y = F.layer_norm(hs, (C,), w[p + "norm2.weight"], w[p + "norm2.bias"], eps)
        y = F.linear(F.gelu(F.linear(y, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])), w[p + "mlp.fc2.weight"],
                     w[p + "mlp.fc2.bias"])
        hs = y * w[p + "layer_scale2.lambda1"] + hs
    hs = F.layer_norm(hs, (C,), w["layernorm.weight"], w["layernorm.bias"], eps)
    return hs[:, 1:]


def vit_adapter_oracle(sd: Dict[str, torch.Tensor], x: torch.Tensor, dtype, heads: int = 6, layers: int = 12,
                       eps: float = 1e-12, prefix: str = "model.") -> torch.Tensor:
    """ViT_Adapter.forward (autoregressive/models/vit_adapter.py:13-15): HF ViTModel(x, interpolate_pos_encoding=True)
    .last_hidden_state[:, 1:] — restated from the installed transformers 5.5.0 modeling_vit.py (ViTEmbeddings.forward /
    interpolate_pos_encoding :60-128, ViTLayer :315-346, ViTSelfAttention :199-252 (sdpa), final layernorm).  No input
    resize, patch 16, no LayerScale, position table interpolated in the MODEL dtype (Dinov2 interpolates in fp32)."""
    w = {k[len(prefix):]: v.to(dtype) for k, v in sd.items() if k.startswith(prefix)}
    x = x.to(dtype)
    B, _, H, W = x.shape
    P = w["embeddings.patch_embeddings.projection.weight"].shape[-1]
    e = F.conv2d(x, w["embeddings.patch_embeddings.projection.weight"], w["embeddings.patch_embeddings.projection.bias"],
                 stride=P).flatten(2).transpose(1, 2)
    C = e.shape[-1]
    e = torch.cat([w["embeddings.cls_token"].expand(B, -1, -1), e], dim=1)
    pos = w["embeddings.position_embeddings"]
    G = int(round((pos.shape[1] - 1) ** 0.5))
    h, wd = H // P, W // P
    if not (h * wd == G * G and H == W):
        pp = pos[:, 1:].reshape(1, G, G, C).permute(0, 3, 1, 2)
        pp = F.interpolate(pp, size=(h, wd), mode="bicubic", align_corners=False)
        pos = torch.cat([pos[:, :1], pp.permute(0, 2, 3, 1).reshape(1, -1, C)], dim=1)
    hs = e + pos
    for l in range(layers):
        p = f"encoder.layer.{l}."
        y = F.layer_norm(hs, (C,), w[p + "layernorm_before.weight"], w[p + "layernorm_before.bias"], eps)
        q = F.linear(y, w[p + "attention.attention.query.weight"], w[p + "attention.attention.query.bias"])
        k = F.linear(y, w[p + "attention.attention.key.weight"], w[p + "attention.attention.key.bias"])
        v = F.linear(y, w[p + "attention.attention.value.weight"], w[p + "attention.attention.value.bias"])
        sh = lambda t: t.view(B, -1, heads, C // heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(sh(q), sh(k), sh(v), scale=(C // heads) ** -0.5)
        a = a.transpose(1, 2).reshape(B, -1, C)
        hs = F.linear(a, w[p + "attention.output.dense.weight"], w[p + "attention.output.dense.bias"]) + hs
        y = F.layer_norm(hs, (C,), w[p + "layernorm_after.weight"], w[p + "layernorm_after.bias"], eps)
        y = F.gelu(F.linear(y, w[p + "intermediate.dense.weight"], w[p + "intermediate.dense.bias"]))
        hs = F.linear(y, w[p + "output.dense.weight"], w[p + "output.dense.bias"]) + hs
    hs = F.layer_norm(hs, (C,), w["layernorm.weight"], w["layernorm.bias"], eps)
    return hs[:, 1:]


# ---------------------------------------------------------------------------------------------------------------
def _gn(x, w, p):
    return F.group_norm(x, 32, w[p + ".weight"], w[p + ".bias"], 1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _res(x, w, p):
    h = F.conv2d(_swish(_gn(x, w, p + "norm1")), w[p + "conv1.weight"], w[p + "conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(h, w, p + "norm2")), w[p + "conv2.weight"], w[p + "conv2.bias"], padding=1)
    if (p + "nin_shortcut.weight") in w:
        x = F.conv2d(x, w[p + "nin_shortcut.weight"], w[p + "nin_shortcut.bias"])
    return x + h


def _attn(x, w, p):
    h = _gn(x, w, p + "norm")
    q = F.conv2d(h, w[p + "q.weight"], w[p + "q.bias"])
    k = F.conv2d(h, w[p + "k.weight"], w[p + "k.bias"])
    v = F.conv2d(h, w[p + "v.weight"], w[p + "v.bias"])
    b, c, hh, ww = q.shape
    a = torch.bmm(q.reshape(b, c, -1).permute(0, 2, 1), k.reshape(b, c, -1)) * (int(c) ** -0.5)
    a = F.softmax(a, dim=2)
    o = torch.bmm(v.reshape(b, c, -1), a.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(o, w[p + "proj_out.weight"], w[p + "proj_out.bias"])


def vq_decode_oracle(sd: Dict[str, torch.Tensor], codes: torch.Tensor, shape, n_levels: int = 5) -> torch.Tensor:
    w = {k: v.float() for k, v in sd.items()}
    B, e, h, wd = shape
    emb = F.normalize(w["quantize.embedding.weight"], p=2, dim=-1)
    z = emb[codes.reshape(-1).long()].reshape(B, h, wd, e).permute(0, 3, 1, 2).contiguous()
    x = F.conv2d(z, w["post_quant_conv.weight"], w["post_quant_conv.bias"])
    x = F.conv2d(x, w["decoder.conv_in.weight"], w["decoder.conv_in.bias"], padding=1)
    x = _res(x, w, "decoder.mid.0."); x = _attn(x, w, "decoder.mid.1."); x = _res(x, w, "decoder.mid.2.")
    for idx in range(n_levels):
        for b in range(3):
            x = _res(x, w, f"decoder.conv_blocks.{idx}.res.{b}.")
            if f"decoder.conv_blocks.{idx}.attn.{b}.norm.weight" in w:
                x = _attn(x, w, f"decoder.conv_blocks.{idx}.attn.{b}.")
        if f"decoder.conv_blocks.{idx}.upsample.conv.weight" in w:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, w[f"decoder.conv_blocks.{idx}.upsample.conv.weight"], w[f"decoder.conv_blocks.{idx}.upsample.conv.bias"], padding=1)
    x = _swish(_gn(x, w, "decoder.norm_out"))
    return F.conv2d(x, w["decoder.conv_out.weight"], w["decoder.conv_out.bias"], padding=1)


def vq_encode_oracle(sd: Dict[str, torch.Tensor], img: torch.Tensor, n_levels: int = 5):
    w = {k: v.float() for k, v in sd.items()}
    x = F.conv2d(img.float(), w["encoder.conv_in.weight"], w["encoder.conv_in.bias"], padding=1)
    for lvl in range(n_levels):
        for b in range(2):
            x = _res(x, w, f"encoder.conv_blocks.{lvl}.res.{b}.")
            if f"encoder.conv_blocks.{lvl}.attn.{b}.norm.weight" in w:
                x = _attn(x, w, f"encoder.conv_blocks.{lvl}.attn.{b}.")
        if f"encoder.conv_blocks.{lvl}.downsample.conv.weight" in w:
            x = F.conv2d(F.pad(x, (0, 1, 0, 1)), w[f"encoder.conv_blocks.{lvl}.downsample.conv.weight"],
                         w[f"encoder.conv_blocks.{lvl}.downsample.conv.bias"], stride=2)
    x = _res(x, w, "encoder.mid.0."); x = _attn(x, w, "encoder.mid.1."); x = _res(x, w, "encoder.mid.2.")
    x = F.conv2d(_swish(_gn(x, w, "encoder.norm_out")), w["encoder.conv_out.weight"], w["encoder.conv_out.bias"], padding=1)
    z = F.conv2d(x, w["quant_conv.weight"], w["quant_conv.bias"])
    zf = F.normalize(z.permute(0, 2, 3, 1).reshape(-1, z.shape[1]), p=2, dim=-1)
    emb = F.normalize(w["quantize.embedding.weight"], p=2, dim=-1)
    d = (zf ** 2).sum(1, keepdim=True) + (emb ** 2).sum(1) - 2 * zf @ emb.t()       # vq_model.py:228-230
    idx = torch.argmin(d, dim=1)
    return idx, z, d
