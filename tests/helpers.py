"""Shared test helpers: build the product modules / oracle from (spec, seed) and load golden fixtures."""
from __future__ import annotations

import os

import torch

from oracle.weights import GPTSpec, make_gpt_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name: str):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def dtype_of(g) -> torch.dtype:
    return {"torch.bfloat16": torch.bfloat16, "torch.float32": torch.float32}[g["dtype"]]


def build_product_gpt(spec: GPTSpec, seed: int, dtype, device="cuda"):
    from controlar_b200.autoregressive.models.gpt_t2i import Transformer, ModelArgs
    m = Transformer(ModelArgs(dim=spec.dim, n_layer=spec.n_layer, n_head=spec.n_head, multiple_of=spec.multiple_of,
                              vocab_size=spec.vocab_size, cls_token_num=spec.cls_token_num, block_size=spec.block_size,
                              caption_dim=spec.caption_dim, num_classes=spec.num_classes, model_type=spec.model_type,
                              adapter_size=spec.adapter_size, condition_type=spec.condition_type))
    sd = make_gpt_state_dict(spec, seed)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return m.to(device=device, dtype=dtype).eval(), sd


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def near_tie_bound(raw_absmax: float, cfg_scale: float) -> float:
    """Largest logit margin that two bf16 implementations can legitimately disagree on: each raw logit may land on
    either neighbouring bf16 value (1 ulp at the magnitude of the largest logits) and CFG combines
    u + (c - u) * s, i.e. amplifies c by s and u by (s - 1); both candidates can move, hence the factor 2."""
    import math
    ulp = 2.0 ** (math.floor(math.log2(max(raw_absmax, 1e-30))) - 7)
    amp = (2.0 * cfg_scale - 1.0) if cfg_scale > 1.0 else 1.0
    return 2.0 * amp * ulp


def assert_mismatches_are_near_ties(z_ref_combined, raw_ref, ref_tok, mine_tok, cfg_scale, what=""):
    """Every position where `mine_tok` differs from the reference's greedy token must be a near-tie in the
    REFERENCE's own (CFG-combined) logits."""
    mism = (mine_tok != ref_tok)
    for b, i in mism.nonzero().tolist():
        margin = float(z_ref_combined[b, i, ref_tok[b, i]] - z_ref_combined[b, i, mine_tok[b, i]])
        bound = near_tie_bound(float(raw_ref[:, i].abs().max()), cfg_scale)
        assert margin <= bound, f"{what}: token ({b},{i}) differs with margin {margin:.4f} > near-tie bound {bound:.4f}"
    return float(mism.float().mean())
