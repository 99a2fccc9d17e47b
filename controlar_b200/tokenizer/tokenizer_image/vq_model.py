"""Drop-in for the reference ``tokenizer/tokenizer_image/vq_model.py`` (VQGAN image tokenizer): same
``ModelArgs`` / ``VQ_models`` / module and parameter names (identical state-dict keys, vq_model.py:28-61), with
``encode`` / ``decode`` / ``decode_code`` / ``forward`` running in the library's kernels (implicit-GEMM
convolutions on tensor cores, GroupNorm/swish, single-head attention, fused quantiser).  Inference only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class ModelArgs:                       # reference vq_model.py:12-24
    codebook_size: int = 16384
    codebook_embed_dim: int = 8
    codebook_l2_norm: bool = True
    codebook_show_usage: bool = True
    commit_loss_beta: float = 0.25
    entropy_loss_ratio: float = 0.0
    encoder_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    decoder_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    z_channels: int = 256
    dropout_p: float = 0.0


def _gn(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):          # parameter container (reference vq_model.py:280-315)
    def __init__(self, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, norm_type="group"):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _gn(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _gn(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                raise NotImplementedError("conv_shortcut=True is never used by the reference configs")
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):            # reference vq_model.py:318-352
    def __init__(self, in_channels, norm_type="group"):
        super().__init__()
        self.norm = _gn(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)


class Encoder(nn.Module):              # reference vq_model.py:65-125
    def __init__(self, in_channels=3, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, norm_type="group", dropout=0.0,
                 resamp_with_conv=True, z_channels=256):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        in_mult = (1,) + tuple(ch_mult)
        self.conv_blocks = nn.ModuleList()
        block_in = ch
        for lvl in range(self.num_resolutions):
            blk = nn.Module()
            res, attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_mult[lvl], ch * ch_mult[lvl]
            for _ in range(num_res_blocks):
                res.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if lvl == self.num_resolutions - 1:
                    attn.append(AttnBlock(block_in))
            blk.res, blk.attn = res, attn
            if lvl != self.num_resolutions - 1:
                blk.downsample = Downsample(block_in)
            self.conv_blocks.append(blk)
        self.mid = nn.ModuleList([ResnetBlock(block_in, block_in), AttnBlock(block_in), ResnetBlock(block_in, block_in)])
        self.norm_out = _gn(block_in)
        self.conv_out = nn.Conv2d(block_in, z_channels, 3, 1, 1)


class Decoder(nn.Module):              # reference vq_model.py:129-195
    def __init__(self, z_channels=256, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, norm_type="group", dropout=0.0,
                 resamp_with_conv=True, out_channels=3):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.ModuleList([ResnetBlock(block_in, block_in), AttnBlock(block_in), ResnetBlock(block_in, block_in)])
        self.conv_blocks = nn.ModuleList()
        for lvl in reversed(range(self.num_resolutions)):
            blk = nn.Module()
            res, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                res.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if lvl == self.num_resolutions - 1:
                    attn.append(AttnBlock(block_in))
            blk.res, blk.attn = res, attn
            if lvl != 0:
                blk.upsample = Upsample(block_in)
            self.conv_blocks.append(blk)
        self.norm_out = _gn(block_in)
        self.conv_out = nn.Conv2d(block_in, out_channels, 3, 1, 1)

    @property
    def last_layer(self):
        return self.conv_out.weight


class VectorQuantizer(nn.Module):      # reference vq_model.py:198-277 (parameters + codebook init)
    def __init__(self, n_e, e_dim, beta, entropy_loss_ratio, l2_norm, show_usage):
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.entropy_loss_ratio, self.l2_norm, self.show_usage = entropy_loss_ratio, l2_norm, show_usage
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)
        if l2_norm:
            self.embedding.weight.data = F.normalize(self.embedding.weight.data, p=2, dim=-1)
        if show_usage:
            self.register_buffer("codebook_used", torch.zeros(65536))


class VQModel(nn.Module):
    def __init__(self, config: ModelArgs):
        super().__init__()
        if not config.codebook_l2_norm:
            raise NotImplementedError("controlar_b200: only the l2-normalised codebook used by ControlAR is implemented")
        self.config = config
        self.encoder = Encoder(ch_mult=config.encoder_ch_mult, z_channels=config.z_channels, dropout=config.dropout_p)
        self.decoder = Decoder(ch_mult=config.decoder_ch_mult, z_channels=config.z_channels, dropout=config.dropout_p)
        self.quantize = VectorQuantizer(config.codebook_size, config.codebook_embed_dim, config.commit_loss_beta,
                                        config.entropy_loss_ratio, config.codebook_l2_norm, config.codebook_show_usage)
        self.quant_conv = nn.Conv2d(config.z_channels, config.codebook_embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(config.codebook_embed_dim, config.z_channels, 1)
        self._car_vq = None

    def _handle(self):
        from ... import vision as _vision
        if self._car_vq is None:
            object.__setattr__(self, "_car_vq", _vision.VQHandle(self))
        return self._car_vq

    @torch.no_grad()
    def encode(self, x):
        """-> (quant [B,e,h,w], (vq_loss, commit_loss, entropy_loss, usage), (None, None, indices int64 [B*h*w]))
        as reference vq_model.py:41-46 in eval mode (losses None, usage 0)."""
        if self.training:
            raise NotImplementedError("controlar_b200: tokenizer training is out of scope (SURVEY.md §2a row 16)")
        quant, idx = self._handle().encode(x)
        return quant, (None, None, None, 0), (None, None, idx.to(torch.int64))

    @torch.no_grad()
    def decode_code(self, code_b, shape=None, channel_first=True):
        """codes -> image [B,3,H,W] fp32 (reference vq_model.py:53-56).  shape = [B, e_dim, h, w]."""
        if shape is None or not channel_first:
            raise NotImplementedError("decode_code needs shape=[B, C, h, w] with channel_first=True (the only form ControlAR uses)")
        B, _, h, w = [int(v) for v in shape]
        return self._handle().decode_code(code_b, B, h, w)

    @torch.no_grad()
    def decode(self, quant):
        """quant [B,e,h,w] -> image (reference vq_model.py:48-51)."""
        return self._handle().decode(quant)

    def forward(self, input):
        quant, diff, _ = self.encode(input)
        return self.decode(quant), diff


def VQ_8(**kwargs):
    return VQModel(ModelArgs(encoder_ch_mult=[1, 2, 2, 4], decoder_ch_mult=[1, 2, 2, 4], **kwargs))


def VQ_16(**kwargs):
    return VQModel(ModelArgs(encoder_ch_mult=[1, 1, 2, 2, 4], decoder_ch_mult=[1, 1, 2, 2, 4], **kwargs))


VQ_models = {"VQ-16": VQ_16, "VQ-8": VQ_8}
